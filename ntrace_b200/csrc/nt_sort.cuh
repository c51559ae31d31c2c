// ntrace_b200 — hand-written scan and radix-sort primitives shared by the BVH builder and the ray sorter
// (no CUB / thrust; replaces thrust::sort_by_key, src/rt/bvh/HLBVH/radixSort.cu:22-46).
#pragma once
#include "nt_common.cuh"

namespace nt {
namespace {

typedef unsigned int uint;
typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------
// Exclusive scan (reduce / scan block sums / apply), T = uint or u64.  Hand-written, no CUB/thrust.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <class T>
__device__ __forceinline__ T block_exclusive(T v, T* s_warp, T& total)
{
    // exclusive scan of one value per thread across a 256-thread block
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = (lane < kScanThreads / 32) ? s_warp[lane] : T(0);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { T y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane < kScanThreads / 32) s_warp[lane] = x;          // inclusive over warps
    }
    __syncthreads();
    const T warpBase = (w == 0) ? T(0) : s_warp[w - 1];
    total = s_warp[kScanThreads / 32 - 1];
    __syncthreads();
    return warpBase + inc - v;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const T* __restrict__ in, long long n, T* __restrict__ blockSums)
{
    __shared__ T s_warp[kScanThreads / 32];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    T sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) if (base + i < n) sum += in[base + i];
    T total;
    block_exclusive<T>(sum, s_warp, total);
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(T* __restrict__ blockSums, int numBlocks, T* __restrict__ grandTotal)
{
    __shared__ T s_warp[kScanThreads / 32];
    T carry = 0;
    for (int base = 0; base < numBlocks; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const T v = (i < numBlocks) ? blockSums[i] : T(0);
        T total;
        const T ex = block_exclusive<T>(v, s_warp, total);
        if (i < numBlocks) blockSums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && grandTotal) *grandTotal = carry;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const T* __restrict__ in, T* __restrict__ out, long long n, const T* __restrict__ blockSums)
{
    __shared__ T s_warp[kScanThreads / 32];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    T v[kScanItems];
    T sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) { v[i] = (base + i < n) ? in[base + i] : T(0); sum += v[i]; }
    T total;
    T run = block_exclusive<T>(sum, s_warp, total) + blockSums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; i++) { if (base + i < n) out[base + i] = run; run += v[i]; }
}

template <class T>
cudaError_t exclusive_scan(const T* in, T* out, long long n, T* blockSums /* >= ceil(n/tile) */, T* grandTotal, cudaStream_t s, int* launches)
{
    const int nb = (int)((n + kScanTile - 1) / kScanTile);
    scan_reduce_kernel<T><<<nb, kScanThreads, 0, s>>>(in, n, blockSums);
    scan_sums_kernel<T><<<1, kScanThreads, 0, s>>>(blockSums, nb, grandTotal);
    scan_apply_kernel<T><<<nb, kScanThreads, 0, s>>>(in, out, n, blockSums);
    *launches += 3;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort of (key, index) pairs, 8-bit digits.
// ------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;                       // rounds of 32 consecutive keys per warp
constexpr int kSortTile = kSortThreads * kSortItems;

template <class KeyT>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const KeyT* __restrict__ keys, int n, int shift, uint* __restrict__ hist, int numBlocks)
{
    __shared__ uint s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortItems; r++) {
        const int i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&s_hist[(uint)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * numBlocks + blockIdx.x] = s_hist[threadIdx.x];     // digit-major: one scan gives global offsets
}

template <class KeyT>
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const KeyT* __restrict__ keysIn, const int* __restrict__ idxIn,
                                                                     KeyT* __restrict__ keysOut, int* __restrict__ idxOut,
                                                                     int n, int shift, const uint* __restrict__ histScan, int numBlocks)
{
    __shared__ uint s_cnt[kSortThreads / 32][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();

    const int segBase = blockIdx.x * kSortTile + w * (kSortItems * 32);
    KeyT key[kSortItems];
    uint rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; r++) {
        const int i = segBase + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keysIn[i] : KeyT(0);
        const uint digit = valid ? ((uint)(key[r] >> shift) & 255u) : 256u;
        const uint peers = __match_any_sync(0xffffffffu, digit);
        uint pre = 0;
        if (valid) pre = s_cnt[w][digit];
        __syncwarp();
        if (valid && lane == (31 - __clz(peers))) s_cnt[w][digit] = pre + __popc(peers);
        __syncwarp();
        rank[r] = pre + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    {
        // digit = threadIdx.x: exclusive prefix over the warps of this tile + global base of (digit, tile)
        uint run = histScan[threadIdx.x * numBlocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < kSortThreads / 32; ww++) { const uint c = s_cnt[ww][threadIdx.x]; s_cnt[ww][threadIdx.x] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; r++) {
        const int i = segBase + r * 32 + lane;
        if (i < n) {
            const uint pos = s_cnt[w][(uint)(key[r] >> shift) & 255u] + rank[r];
            keysOut[pos] = key[r];
            idxOut[pos] = idxIn[i];
        }
    }
}

// Stable LSD radix sort of (key, index) pairs: `passes` 8-bit digits starting at bit 0.  Ping-pongs between the A and B
// buffers; with an even number of passes the sorted data ends in A.  hist needs 256 * tiles uints, blockSums as below.
template <class KeyT>
cudaError_t radix_sort_pairs(KeyT* keysA, int* idxA, KeyT* keysB, int* idxB, int n, int passes,
                             uint* hist, uint* blockSums, cudaStream_t stream, int* launches)
{
    const int nb = (n + kSortTile - 1) / kSortTile;
    const long long histLen = (long long)nb * 256;
    KeyT* kin = keysA; int* iin = idxA; KeyT* kout = keysB; int* iout = idxB;
    for (int pass = 0; pass < passes; pass++) {
        const int shift = pass * 8;
        radix_hist_kernel<KeyT><<<nb, kSortThreads, 0, stream>>>(kin, n, shift, hist, nb);
        *launches += 1;
        cudaError_t e = exclusive_scan<uint>(hist, hist, histLen, blockSums, nullptr, stream, launches);
        if (e != cudaSuccess) return e;
        radix_scatter_kernel<KeyT><<<nb, kSortThreads, 0, stream>>>(kin, iin, kout, iout, n, shift, hist, nb);
        *launches += 1;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        KeyT* tk = kin; kin = kout; kout = tk;
        int* ti = iin; iin = iout; iout = ti;
    }
    return cudaSuccess;
}
inline size_t radix_hist_bytes(int n) { return (size_t)((n + kSortTile - 1) / kSortTile) * 256 * 4; }
inline size_t scan_block_sums_bytes(long long len) { return ((size_t)((len + kScanTile - 1) / kScanTile) + 16) * 8; }

} // namespace
} // namespace nt
