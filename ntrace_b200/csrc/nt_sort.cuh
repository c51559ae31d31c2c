// ntrace_b200 — hand-written scan and radix-sort primitives shared by the BVH builder and the ray sorter
// (no CUB / thrust; replaces thrust::sort_by_key, src/rt/bvh/HLBVH/radixSort.cu:22-46).
#pragma once
#include "nt_common.cuh"
#include <cstdlib>

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

// Single-pass form (chained scan with decoupled look-back): one kernel instead of three.  `desc` holds one zeroed word per tile
// (flag in the two top bits: 1 = tile total, 2 = inclusive prefix; the values scanned here stay below 2^30 per 32-bit half),
// `ticket` is a zeroed counter that hands out tile indices in launch order.  Warp 0 looks back 32 tiles at a time.
template <class T> struct ScanFlag;
template <> struct ScanFlag<uint> { static constexpr int shift = 30; };
template <> struct ScanFlag<u64>  { static constexpr int shift = 62; };

template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_chained_kernel(const T* __restrict__ in, T* __restrict__ out, long long n,
                                                                    T* __restrict__ desc, uint* __restrict__ ticket, T* __restrict__ grandTotal)
{
    constexpr int SH = ScanFlag<T>::shift;
    constexpr T VMASK = (T(1) << SH) - T(1);
    __shared__ T s_warp[kScanThreads / 32];
    __shared__ T s_prefix;
    __shared__ uint s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const long long tile = s_tile;
    const long long base = tile * kScanTile + (long long)threadIdx.x * kScanItems;
    T v[kScanItems];
    T sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) { v[i] = (base + i < n) ? in[base + i] : T(0); sum += v[i]; }
    T total;
    const T ex = block_exclusive<T>(sum, s_warp, total);
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        volatile T* d = desc;
        if (lane == 0) d[tile] = (T(tile == 0 ? 2 : 1) << SH) | total;
        T excl = 0;
        for (long long t0 = tile - 1; t0 >= 0; t0 -= 32) {
            const long long t = t0 - lane;
            T w = T(2) << SH;                                   // before the first tile: an inclusive prefix of 0
            if (t >= 0) { do { w = d[t]; } while ((w >> SH) == T(0)); }
            const uint done = __ballot_sync(0xffffffffu, (w >> SH) == T(2));
            const int first = done ? (__ffs(done) - 1) : 31;    // nearest tile that already knows its inclusive prefix
            T c = (lane <= first) ? (w & VMASK) : T(0);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            excl += c;
            if (done) break;
        }
        if (lane == 0) {
            if (tile > 0) d[tile] = (T(2) << SH) | (excl + total);
            s_prefix = excl;
            if (grandTotal && (tile + 1) * kScanTile >= n) *grandTotal = excl + total;
        }
    }
    __syncthreads();
    T run = ex + s_prefix;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) { if (base + i < n) out[base + i] = run; run += v[i]; }
}
inline size_t scan_desc_bytes(long long n) { return ((size_t)((n + kScanTile - 1) / kScanTile) + 2) * 8; }

template <class T>
cudaError_t exclusive_scan_chained(const T* in, T* out, long long n, T* zeroedDesc, uint* zeroedTicket, T* grandTotal, cudaStream_t s, int* launches)
{
    const int nb = (int)((n + kScanTile - 1) / kScanTile);
    scan_chained_kernel<T><<<nb, kScanThreads, 0, s>>>(in, out, n, zeroedDesc, zeroedTicket, grandTotal);
    *launches += 1;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort of (key, index) pairs, 8-bit digits.
// ------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
// Keys per thread (rounds of 32 consecutive keys per warp).  Measured on the 10 M-triangle builds: 8 (63 registers, 28 KB of
// shared memory) beats 12 and 16 -- the longer runs per digit of a bigger tile do not pay for the lost occupancy.
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;

// lanes of the warp that hold the same digit (0..255; 256 = no key).  NT_SORT_MATCH_BALLOT: nine ballots instead of MATCH.ANY,
// whose cost grows with the number of distinct values in the warp (measured: profiles/r2_summary.md, builder section)
#ifndef NT_SORT_MATCH_BALLOT
#define NT_SORT_MATCH_BALLOT 1
#endif
__device__ __forceinline__ uint match_digit(uint digit)
{
#if NT_SORT_MATCH_BALLOT
    uint peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 9; b++) {
        const bool bit = (digit >> b) & 1u;
        const uint m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
#else
    return __match_any_sync(0xffffffffu, digit);
#endif
}

template <class KeyT>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const KeyT* __restrict__ keys, int n, int shift, uint* __restrict__ hist, int numBlocks)
{
    constexpr int ITEMS = kSortItems;
    // one sub-histogram per warp: the keys of a tile often share a few digits (Morton codes of neighbouring triangles), and
    // eight warps hammering the same shared-memory counters serialise
    __shared__ uint s_hist[kSortThreads / 32][256];
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kSortThreads / 32; i++) s_hist[i][threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * (kSortThreads * ITEMS);
    KeyT k[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const int i = base + r * kSortThreads + threadIdx.x;
        k[r] = (i < n) ? keys[i] : KeyT(0);
    }
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const int i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&s_hist[w][(uint)(k[r] >> shift) & 255u], 1u);
    }
    __syncthreads();
    uint sum = 0;
#pragma unroll
    for (int i = 0; i < kSortThreads / 32; i++) sum += s_hist[i][threadIdx.x];
    hist[threadIdx.x * numBlocks + blockIdx.x] = sum;                     // digit-major: one scan gives global offsets
}

// Scatter of one pass.  Ranks: warp w owns ITEMS rounds of 32 consecutive keys; within a round MATCH.ANY groups equal digits,
// the per-warp counters carry the running count, so the order inside a digit is the input order (stable).  The tile is then
// put in digit order in shared memory and written out by consecutive threads, so each digit's run leaves as contiguous
// segments instead of 32 scattered 4-byte stores per warp.
// TMA experiment (north_star: "TMA/shared-memory staging where it measurably helps"): the tile's keys arrive in shared memory through ONE
// bulk asynchronous copy (cp.async.bulk global -> shared, completion on an mbarrier) issued by one thread, instead of eight coalesced
// 128-byte-per-warp loads per thread.  BULK = true selects it for full tiles (the copy needs 16-byte granularity); NT_SORT_BULK=1 at run
// time.  Measured on the B200 (scripts/tma_experiments.sh, profiles/r2i_tma_experiments.txt): 10 M-triangle LBVH build 1.731 vs 1.736 ms,
// 10.5 M scene 1.625 vs 1.626 ms, sorted output identical -- the pass is bound by the ranking arithmetic and the scattered stores, not by
// the latency of its (already coalesced) tile load -- so the default stays the plain loads.
__device__ __forceinline__ uint smem_u32(const void* p) { return (uint)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load_tile(void* smemDst, const void* gmemSrc, uint bytes, unsigned long long* bar)
{
    // one thread: arm the barrier with the byte count, start the copy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smemDst)), "l"(gmemSrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_wait(unsigned long long* bar, uint phase)
{
    uint ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!ok);
}

template <class KeyT, bool BULK>
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const KeyT* __restrict__ keysIn, const int* __restrict__ idxIn,
                                                                     KeyT* __restrict__ keysOut, int* __restrict__ idxOut,
                                                                     int n, int shift, const uint* __restrict__ histScan, int numBlocks)
{
    constexpr int ITEMS = kSortItems;
    constexpr int TILE = kSortTile;
    __shared__ uint s_cnt[kSortThreads / 32][256];
    __shared__ uint s_digitBase[256];          // first local position of each digit
    __shared__ uint s_outOfs[256];             // global position = s_outOfs[digit] + local position (mod 2^32)
    __shared__ uint s_warp[kSortThreads / 32];
    __shared__ __align__(16) KeyT s_key[TILE];
    __shared__ int  s_idx[TILE];
    __shared__ __align__(8) unsigned long long s_bar;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tileBase = blockIdx.x * TILE;
    const bool bulk = BULK && (tileBase + TILE <= n);
    if (BULK) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar)) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    if (bulk) {
        if (threadIdx.x == 0) bulk_load_tile(s_key, keysIn + tileBase, (uint)(TILE * sizeof(KeyT)), &s_bar);
        bulk_wait(&s_bar, 0);
    }

    const int segBase = tileBase + w * (ITEMS * 32);
    KeyT key[ITEMS];
    uint rank[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const int i = segBase + r * 32 + lane;
        const bool valid = i < n;
        key[r] = bulk ? s_key[i - tileBase] : (valid ? keysIn[i] : KeyT(0));
        const uint digit = valid ? ((uint)(key[r] >> shift) & 255u) : 256u;
        const uint peers = match_digit(digit);
        uint pre = 0;
        if (valid) pre = s_cnt[w][digit];
        __syncwarp();
        if (valid && lane == (31 - __clz(peers))) s_cnt[w][digit] = pre + __popc(peers);
        __syncwarp();
        rank[r] = pre + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    {
        // digit = threadIdx.x: exclusive prefix over the warps of this tile, then over the digits
        const uint d = threadIdx.x;
        uint run = 0;
#pragma unroll
        for (int ww = 0; ww < kSortThreads / 32; ww++) { const uint c = s_cnt[ww][d]; s_cnt[ww][d] = run; run += c; }
        uint total;
        const uint base = block_exclusive<uint>(run, s_warp, total);
        s_digitBase[d] = base;
        s_outOfs[d] = histScan[d * numBlocks + blockIdx.x] - base;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const int i = segBase + r * 32 + lane;
        if (i < n) {
            const uint d = (uint)(key[r] >> shift) & 255u;
            const uint lp = s_digitBase[d] + s_cnt[w][d] + rank[r];
            s_key[lp] = key[r];
            s_idx[lp] = idxIn[i];
        }
    }
    __syncthreads();
    const int count = min(TILE, n - tileBase);
    for (int j = threadIdx.x; j < count; j += kSortThreads) {
        const KeyT k = s_key[j];
        const uint pos = s_outOfs[(uint)(k >> shift) & 255u] + (uint)j;
        keysOut[pos] = k;
        idxOut[pos] = s_idx[j];
    }
}

// ------------------------------------------------------------------------------------------------
// One-sweep form of the same sort (decoupled look-back): per pass ONE kernel reads a tile, ranks it exactly like
// radix_scatter_kernel (so the pass is stable) and gets the tile's global digit offsets by looking back over the descriptors
// of the tiles before it instead of from a scanned per-tile histogram.  Per sort that is 1 memset + 1 histogram kernel (all
// digits at once) + `passes` sweeps, instead of `passes` x (histogram, 3 scan kernels, scatter): at 283 K keys the sort was
// 20 launches of ~5 us each, at 10 M keys the per-pass histogram + scans re-read every key.
// Descriptor word: flag (2 bits: 0 not ready, 1 tile count, 2 inclusive prefix) | value (30 bits; n < 2^30).
// Tiles take their index from a ticket counter, so a tile only ever waits for tiles that were handed out before it.
// The look-back runs AFTER the tile has been staged in digit order in shared memory (which needs local offsets only), so its
// latency overlaps the staging instead of sitting between ranking and staging.
// ------------------------------------------------------------------------------------------------
constexpr uint kOsFlagShift = 30, kOsValueMask = 0x3fffffffu;
constexpr int kOsMaxPasses = 8;
// zone (uints, zeroed by ONE memset per sort): [0, 8*256) global digit histograms, 8 tickets, 8 pad, then passes * tiles * 256 descriptors
constexpr int kOsHeadWords = kOsMaxPasses * 256 + 16;
inline size_t onesweep_zone_bytes(int n, int passes) { return ((size_t)kOsHeadWords + (size_t)passes * ((n + kSortTile - 1) / kSortTile) * 256) * 4; }

template <class KeyT>
__global__ void __launch_bounds__(kSortThreads) radix_hist_all_kernel(const KeyT* __restrict__ keys, int n, int passes, uint* __restrict__ globalHist)
{
    constexpr int ITEMS = kSortItems;
    __shared__ uint s_hist[kSortThreads / 32][256];     // one sub-histogram per warp (see radix_hist_kernel)
    const int w = threadIdx.x >> 5;
    uint acc[kOsMaxPasses];
#pragma unroll
    for (int p = 0; p < kOsMaxPasses; p++) acc[p] = 0;
    const int numTiles = (n + kSortTile - 1) / kSortTile;
    for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        const int base = tile * kSortTile;
        KeyT k[ITEMS];
#pragma unroll
        for (int r = 0; r < ITEMS; r++) {
            const int i = base + r * kSortThreads + threadIdx.x;
            k[r] = (i < n) ? keys[i] : KeyT(0);
        }
#pragma unroll
        for (int p = 0; p < kOsMaxPasses; p++) {
            if (p < passes) {
#pragma unroll
                for (int i = 0; i < kSortThreads / 32; i++) s_hist[i][threadIdx.x] = 0;
                __syncthreads();
#pragma unroll
                for (int r = 0; r < ITEMS; r++) {
                    const int i = base + r * kSortThreads + threadIdx.x;
                    if (i < n) atomicAdd(&s_hist[w][(uint)(k[r] >> (8 * p)) & 255u], 1u);
                }
                __syncthreads();
                uint sum = 0;
#pragma unroll
                for (int i = 0; i < kSortThreads / 32; i++) sum += s_hist[i][threadIdx.x];
                acc[p] += sum;
                __syncthreads();
            }
        }
    }
#pragma unroll
    for (int p = 0; p < kOsMaxPasses; p++)
        if (p < passes && acc[p]) atomicAdd(globalHist + p * 256 + threadIdx.x, acc[p]);
}

template <class KeyT>
__global__ void __launch_bounds__(kSortThreads) radix_onesweep_kernel(const KeyT* __restrict__ keysIn, const int* __restrict__ idxIn,
                                                                      KeyT* __restrict__ keysOut, int* __restrict__ idxOut,
                                                                      int n, int shift, const uint* __restrict__ digitTotals,
                                                                      uint* __restrict__ desc, uint* __restrict__ ticket)
{
    constexpr int ITEMS = kSortItems;
    constexpr int TILE = kSortTile;
    __shared__ uint s_cnt[kSortThreads / 32][256];
    __shared__ uint s_digitBase[256];
    __shared__ uint s_outOfs[256];
    __shared__ u64  s_warp[kSortThreads / 32];
    __shared__ KeyT s_key[TILE];
    __shared__ int  s_idx[TILE];
    __shared__ uint s_tile;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int tile = (int)s_tile;

    const int tileBase = tile * TILE;
    const int segBase = tileBase + w * (ITEMS * 32);
    KeyT key[ITEMS];
    int  val[ITEMS];
    uint rank[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const int i = segBase + r * 32 + lane;
        key[r] = (i < n) ? keysIn[i] : KeyT(0);
        val[r] = (i < n) ? idxIn[i] : 0;
    }
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const int i = segBase + r * 32 + lane;
        const bool valid = i < n;
        const uint digit = valid ? ((uint)(key[r] >> shift) & 255u) : 256u;
        const uint peers = match_digit(digit);
        uint pre = 0;
        if (valid) pre = s_cnt[w][digit];
        __syncwarp();
        if (valid && lane == (31 - __clz(peers))) s_cnt[w][digit] = pre + __popc(peers);
        __syncwarp();
        rank[r] = pre + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    // digit = threadIdx.x: exclusive prefix over the warps of this tile, the tile's count of the digit (published at once for the
    // tiles after this one), and two prefixes over the digits in one scan: local positions (low word) and global digit bases (high)
    const uint d = threadIdx.x;
    uint run = 0;
#pragma unroll
    for (int ww = 0; ww < kSortThreads / 32; ww++) { const uint c = s_cnt[ww][d]; s_cnt[ww][d] = run; run += c; }
    volatile uint* my = desc + (size_t)tile * 256 + d;
    *my = ((tile == 0 ? 2u : 1u) << kOsFlagShift) | run;
    u64 total;
    const u64 both = block_exclusive<u64>((u64)run | ((u64)__ldg(digitTotals + d) << 32), s_warp, total);
    const uint base = (uint)both, gbase = (uint)(both >> 32);
    s_digitBase[d] = base;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const int i = segBase + r * 32 + lane;
        if (i < n) {
            const uint dg = (uint)(key[r] >> shift) & 255u;
            const uint lp = s_digitBase[dg] + s_cnt[w][dg] + rank[r];
            s_key[lp] = key[r];
            s_idx[lp] = val[r];
        }
    }
    {
        uint excl = 0;
        for (int t = tile - 1; t >= 0; t--) {
            const volatile uint* pd = desc + (size_t)t * 256 + d;
            uint v;
            do { v = *pd; } while ((v >> kOsFlagShift) == 0u);
            excl += v & kOsValueMask;
            if ((v >> kOsFlagShift) == 2u) break;
        }
        if (tile > 0) *my = (2u << kOsFlagShift) | (excl + run);              // inclusive prefix: later tiles stop here
        s_outOfs[d] = gbase + excl - base;
    }
    __syncthreads();
    const int count = min(TILE, n - tileBase);
    for (int j = threadIdx.x; j < count; j += kSortThreads) {
        const KeyT k = s_key[j];
        const uint pos = s_outOfs[(uint)(k >> shift) & 255u] + (uint)j;
        keysOut[pos] = k;
        idxOut[pos] = s_idx[j];
    }
}

// Stable LSD radix sort of (key, index) pairs: `passes` 8-bit digits starting at bit 0.  Ping-pongs between the A and B
// buffers; with an even number of passes the sorted data ends in A.
// One-sweep form (zone != null: onesweep_zone_bytes(n, passes) of scratch; zeroed here unless the caller says it already is):
// 1 memset + 1 histogram kernel + `passes` sweeps.  Classic form (zone == null): hist needs 256 * tiles uints, blockSums as below.
template <class KeyT>
cudaError_t radix_sort_pairs(KeyT* keysA, int* idxA, KeyT* keysB, int* idxB, int n, int passes,
                             uint* hist, uint* blockSums, cudaStream_t stream, int* launches, uint* zone = nullptr, bool zoneIsZero = false)
{
    constexpr int TILE = kSortTile;
    const int nb = (n + TILE - 1) / TILE;
    const long long histLen = (long long)nb * 256;
    KeyT* kin = keysA; int* iin = idxA; KeyT* kout = keysB; int* iout = idxB;
    if (zone && passes <= kOsMaxPasses && n < (1 << 30)) {
        cudaError_t e;
        if (!zoneIsZero) {
            e = cudaMemsetAsync(zone, 0, onesweep_zone_bytes(n, passes), stream);
            if (e != cudaSuccess) return e;
        }
        int hgrid = nb < 592 ? nb : 592;
        radix_hist_all_kernel<KeyT><<<hgrid, kSortThreads, 0, stream>>>(keysA, n, passes, zone);
        *launches += 1;
        for (int pass = 0; pass < passes; pass++) {
            radix_onesweep_kernel<KeyT><<<nb, kSortThreads, 0, stream>>>(kin, iin, kout, iout, n, pass * 8, zone + pass * 256,
                                                                         zone + kOsHeadWords + (size_t)pass * histLen, zone + kOsMaxPasses * 256 + pass);
            *launches += 1;
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            KeyT* tk = kin; kin = kout; kout = tk;
            int* ti = iin; iin = iout; iout = ti;
        }
        return cudaSuccess;
    }
    for (int pass = 0; pass < passes; pass++) {
        const int shift = pass * 8;
        radix_hist_kernel<KeyT><<<nb, kSortThreads, 0, stream>>>(kin, n, shift, hist, nb);
        *launches += 1;
        cudaError_t e = exclusive_scan<uint>(hist, hist, histLen, blockSums, nullptr, stream, launches);
        if (e != cudaSuccess) return e;
        static const bool bulkTile = [] { const char* e = getenv("NT_SORT_BULK"); return e && atoi(e) != 0; }();
        if (bulkTile) radix_scatter_kernel<KeyT, true><<<nb, kSortThreads, 0, stream>>>(kin, iin, kout, iout, n, shift, hist, nb);
        else radix_scatter_kernel<KeyT, false><<<nb, kSortThreads, 0, stream>>>(kin, iin, kout, iout, n, shift, hist, nb);
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
