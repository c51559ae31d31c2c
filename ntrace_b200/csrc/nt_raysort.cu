// ntrace_b200 — GPU Morton sort of a ray batch (RayBuffer::mortonSort).
//
// Replaces src/rt/ray/RayBuffer.cpp:103-163 + src/rt/ray/RayBufferKernels.cu:70-196: the reference finds the AABB of the
// rays on the GPU, builds a 192-bit key per ray (origin and direction, 32 bits per component, bit i of component k at
// key bit k + 6 i), **sorts the keys on the CPU** (multicore quicksort, RayBuffer.cpp:149) and gathers the rays on the GPU.
// Here everything stays on the device: the sort is the same hand-written stable LSD radix sort the BVH builder uses, over the
// FULL key in the comparator's order (compareMortonKey, RayBuffer.cpp:88-99: hash[5] most significant): the 192 bits are held
// as three 64-bit words and sorted least-significant word first — 8 passes on bits 0..63, 8 on bits 64..127 and 4 on bits
// 128..159 (bits 150..191 are always zero: the origin components have 25 significant bits, the direction components 22).
// Rays with identical keys keep their original relative order (the reference's quicksort leaves that case unspecified).
#include "nt_common.cuh"
#include "nt_sort.cuh"
#include <cstdlib>

namespace nt {
namespace {

constexpr float kF32Max = 3.402823466e+38f;
__device__ __forceinline__ int ord(float f) { const int i = __float_as_int(f); return (i >= 0) ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float unord(int i) { return __int_as_float((i >= 0) ? i : i ^ 0x7FFFFFFF); }

__global__ void __launch_bounds__(256) ray_aabb_init_kernel(int* __restrict__ box)
{
    if (threadIdx.x < 6) box[threadIdx.x] = (threadIdx.x < 3) ? ord(kF32Max) : ord(-kF32Max);
}

// findAABBKernel: box over origins and end points origin + direction * tmax
__global__ void __launch_bounds__(256) ray_aabb_kernel(const float4* __restrict__ rays, int n, int* __restrict__ box)
{
    float lo[3] = {kF32Max, kF32Max, kF32Max}, hi[3] = {-kF32Max, -kF32Max, -kF32Max};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
        const float p[3] = {o.x, o.y, o.z};
        const float q[3] = {__fadd_rn(o.x, __fmul_rn(d.x, d.w)), __fadd_rn(o.y, __fmul_rn(d.y, d.w)), __fadd_rn(o.z, __fmul_rn(d.z, d.w))};
#pragma unroll
        for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], fminf(p[k], q[k])); hi[k] = fmaxf(hi[k], fmaxf(p[k], q[k])); }
    }
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(box + k, ord(lo[k])); atomicMax(box + 3 + k, ord(hi[k])); }
}

// (U32)float as the reference's device code converts: saturating, NaN and negatives -> 0
__device__ __forceinline__ unsigned f2u_sat(float f) { return __float2uint_rz(f); }

// genMortonKeysKernel: key bit (k + 6 i) = bit i of component k (collectBits, RayBufferKernels.cu:120-136), as three 64-bit words
__global__ void __launch_bounds__(256) ray_keys_kernel(const float4* __restrict__ rays, int n, const int* __restrict__ box,
                                                        u64* __restrict__ w0, u64* __restrict__ w1, u64* __restrict__ w2, int* __restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
    const float lo[3] = {unord(box[0]), unord(box[1]), unord(box[2])}, hi[3] = {unord(box[3]), unord(box[4]), unord(box[5])};
    const float org[3] = {o.x, o.y, o.z}, dir[3] = {d.x, d.y, d.z};
    float len2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) len2 = __fadd_rn(len2, __fmul_rn(dir[k], dir[k]));
    const float len = __fsqrt_rn(len2);
    const float inv = (len != 0.0f) ? __fdiv_rn(1.0f, len) : 0.0f;
    unsigned comp[6];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float a = __fdiv_rn(__fsub_rn(org[k], lo[k]), __fsub_rn(hi[k], lo[k]));
        const float b = __fmul_rn(__fadd_rn(__fmul_rn(dir[k], inv), 1.0f), 0.5f);
        comp[k] = f2u_sat(__fmul_rn(__fmul_rn(a, 256.0f), 65536.0f));
        comp[3 + k] = f2u_sat(__fmul_rn(__fmul_rn(b, 32.0f), 65536.0f));
    }
    u64 w[3] = {0, 0, 0};
#pragma unroll
    for (int bit = 0; bit < 32; bit++)
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int pos = k + 6 * bit;
            w[pos >> 6] |= (u64)((comp[k] >> bit) & 1u) << (pos & 63);
        }
    w0[i] = w[0]; w1[i] = w[1]; w2[i] = w[2];
    idx[i] = i;
}

// next round of the word-wise LSD sort: the key word of the rays in their current order
__global__ void __launch_bounds__(256) ray_key_gather_kernel(int n, const int* __restrict__ order, const u64* __restrict__ word, u64* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __ldg(word + __ldg(order + i));
}

// reorderRaysKernel
__global__ void __launch_bounds__(256) ray_reorder_kernel(int n, const int* __restrict__ order, const float4* __restrict__ inRays,
                                                           const int* __restrict__ inSlotToID, float4* __restrict__ outRays,
                                                           int* __restrict__ outIDToSlot, int* __restrict__ outSlotToID)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int old = __ldg(order + i);
    const int id = __ldg(inSlotToID + old);
    outRays[2 * i] = __ldg(inRays + 2 * old);
    outRays[2 * i + 1] = __ldg(inRays + 2 * old + 1);
    outIDToSlot[id] = i;
    outSlotToID[i] = id;
}

struct SortScratch { DevBuf keysA, keysB, idxA, idxB, hist, blockSums, box, oldRays, oldS2I, word1, word2, sortAux; };
SortScratch g_ss;
static_assert(sizeof(SortScratch) % sizeof(DevBuf) == 0, "SortScratch holds DevBuf members only");

} // namespace

void release_sort_scratch()
{
    DevBuf* b = reinterpret_cast<DevBuf*>(&g_ss);
    for (size_t i = 0; i < sizeof(SortScratch) / sizeof(DevBuf); i++) b[i].release();
}

cudaError_t ray_sort_device(float4* rays, int* idToSlot, int* slotToID, int n, cudaStream_t stream, int numSMs, int* outLaunches)
{
    int launches = 0;
    cudaError_t e;
#define NT_TRY(call) do { e = (call); if (e != cudaSuccess) { *outLaunches = launches; return e; } } while (0)
    SortScratch& s = g_ss;
    NT_TRY(s.keysA.reserve((size_t)n * 8)); NT_TRY(s.keysB.reserve((size_t)n * 8));
    NT_TRY(s.idxA.reserve((size_t)n * 4)); NT_TRY(s.idxB.reserve((size_t)n * 4));
    NT_TRY(s.hist.reserve(radix_hist_bytes(n)));
    NT_TRY(s.blockSums.reserve(scan_block_sums_bytes((long long)radix_hist_bytes(n) / 4)));
    NT_TRY(s.box.reserve(64)); NT_TRY(s.oldRays.reserve((size_t)n * 32)); NT_TRY(s.oldS2I.reserve((size_t)n * 4));
    NT_TRY(s.word1.reserve((size_t)n * 8)); NT_TRY(s.word2.reserve((size_t)n * 8));
    static const int oneSweepMode = [] { const char* e = getenv("NT_SORT_ONESWEEP"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    const bool oneSweep = oneSweepMode < 0 ? (n <= 2400000) : (oneSweepMode != 0);      // as the builder: while a pass's tiles are all resident
    if (oneSweep) NT_TRY(s.sortAux.reserve(onesweep_zone_bytes(n, 8)));
    uint* aux = oneSweep ? s.sortAux.as<uint>() : nullptr;

    ray_aabb_init_kernel<<<1, 32, 0, stream>>>(s.box.as<int>());
    int grid = (n + 255) / 256;
    if (grid > numSMs * 8) grid = numSMs * 8;
    ray_aabb_kernel<<<grid, 256, 0, stream>>>(rays, n, s.box.as<int>());
    ray_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(rays, n, s.box.as<int>(), s.keysA.as<u64>(), s.word1.as<u64>(), s.word2.as<u64>(), s.idxA.as<int>());
    launches += 3;
    NT_TRY(cudaGetLastError());
    // least significant word first; every round is stable, so after the last one the order is that of the whole 192-bit key
    NT_TRY(radix_sort_pairs<u64>(s.keysA.as<u64>(), s.idxA.as<int>(), s.keysB.as<u64>(), s.idxB.as<int>(), n, 8,
                                 s.hist.as<uint>(), s.blockSums.as<uint>(), stream, &launches, aux));
    const u64* words[2] = {s.word1.as<u64>(), s.word2.as<u64>()};
    const int passes[2] = {8, 4};
    for (int r = 0; r < 2; r++) {
        ray_key_gather_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, s.idxA.as<int>(), words[r], s.keysA.as<u64>());
        launches++;
        NT_TRY(cudaGetLastError());
        NT_TRY(radix_sort_pairs<u64>(s.keysA.as<u64>(), s.idxA.as<int>(), s.keysB.as<u64>(), s.idxB.as<int>(), n, passes[r],
                                     s.hist.as<uint>(), s.blockSums.as<uint>(), stream, &launches, aux));
    }
    NT_TRY(cudaMemcpyAsync(s.oldRays.p, rays, (size_t)n * 32, cudaMemcpyDeviceToDevice, stream));
    NT_TRY(cudaMemcpyAsync(s.oldS2I.p, slotToID, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
    ray_reorder_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, s.idxA.as<int>(), s.oldRays.as<float4>(), s.oldS2I.as<int>(), rays, idToSlot, slotToID);
    launches++;
    NT_TRY(cudaGetLastError());
    *outLaunches = launches;
    return cudaSuccess;
#undef NT_TRY
}

} // namespace nt
