// ntrace_b200 — shared declarations for the sm_100a CUDA sources behind the C ABI (include/ntrace_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace nt {

// Reference: src/rt/kernels/CudaTracerKernels.hpp:36-39, 52-63 (kept bit-identical: these values
// are part of the CudaBVH data format the path exchanges with the host).
enum : int { kEntrypointSentinel = 0x76543210 };
enum BVHLayout : int {
    Layout_AOS_AOS = 0, Layout_AOS_SOA, Layout_SOA_AOS, Layout_SOA_SOA,
    Layout_Compact, Layout_Compact2, Layout_CPU, Layout_Max
};

struct KernelConfig { int bvhLayout, blockWidth, blockHeight, usePersistentThreads; };

// thread-local sticky error string (reference: setError/getError, base/Defs.hpp:143-147)
void set_error(const std::string& msg);
bool check_cuda(cudaError_t e, const char* what, const char* file, int line);

#define NT_CUDA(call)                                                          \
    do {                                                                       \
        if (!::nt::check_cuda((call), #call, __FILE__, __LINE__)) return 1;   \
    } while (0)

// ---- trace (nt_trace.cu) --------------------------------------------------------------------
enum TraceKernelId : int {
    Kernel_PersistentSpeculative = 0,   // persistent while-while, warp-level dynamic ray fetch, speculative leaf
    Kernel_PlainSpeculative = 1,        // one thread per ray, speculative while-while (fermi-style launch shape)
    Kernel_Wide4Persistent = 2,         // persistent speculative while-while over the derived 4-wide quantised node array (nt_wide.cu)
    Kernel_BinaryMr = 3,                // two rays per lane, phase-scheduled (nt_wide.cu "mr"), binary Compact / Compact2 nodes
    Kernel_Wide4Mr = 4,                 // the same over the Wide4 node array
    Kernel_Auto = 5,                    // per batch: any-hit -> Kernel_PersistentSpeculative, closest-hit -> Kernel_Wide4Persistent (API level only)
    Kernel_BinarySw = 6,                // one active + one parked ray per lane, exchanged at the phase boundaries (nt_wide.cu "sw"), binary nodes
    Kernel_Wide4Sw = 7,                 // the same over the Wide4 node array
    Kernel_Count
};

// Several ray batches traced by ONE persistent launch (nt_trace_batches): ray index i of the launch belongs to batch b with
// start[b] <= i < start[b + 1] and is ray i - start[b] of it.  Lives in device memory; read at ray fetch and result store only.
constexpr int kMaxBatches = 64;
struct BatchTable {
    int count; int pad;
    int start[kMaxBatches + 2];
    const float4* rays[kMaxBatches];
    int4* results[kMaxBatches];
};
#ifdef __CUDACC__
__device__ __forceinline__ int batch_of(const BatchTable* __restrict__ bt, int idx)
{
    int lo = 0, hi = __ldg(&bt->count);
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (idx >= __ldg(&bt->start[mid])) lo = mid; else hi = mid; }
    return lo;
}
#endif

struct TraceLaunch {
    int kernel;               // TraceKernelId
    int layout;               // Layout_Compact or Layout_Compact2
    int numRays;
    int anyHit;
    int fast;                 // 1: triangle test with the reference GPU kernels' -use_fast_math arithmetic instead of IEEE
    const float4* rays;       // device, 2 x float4 per ray
    int4* results;            // device, 1 x int4 per ray
    const float4* nodes;      // device, 4 x float4 per node
    const float4* wideNodes = nullptr;   // device, Wide4 node array (Kernel_Wide4Persistent only)
    const float4* woop;       // device
    const int* triIndices;    // device
    int* warpCounter;         // device, zeroed before launch (persistent kernels only)
    int* errorFlag = nullptr; // device-visible (mapped host) word; a kernel ORs 1 into it when a ray's traversal stack would overflow
    const BatchTable* batches = nullptr;   // device; non-null: numRays = rays of all batches, `rays` / `results` unused (one-ray persistent kernels only)
    int numSMs;
    cudaStream_t stream;
};
cudaError_t launch_trace(const TraceLaunch& a, int* outNumLaunches);
// occupancy / carve-out decisions are cached per kernel variant and per launch epoch; nt_shutdown starts a new epoch
int launch_epoch();
void reset_launch_caches();
// nt_wide.cu: 4-wide quantised form of a Compact / Compact2 node buffer (host side; 16 words per node) and its traversal kernel
constexpr int kWideMaxDepth = 42;       // three pushes per level at most: 1 + 3 * depth entries fit the kernel's 128-entry stack
int convert_compact_to_wide4_host(const int32_t* nodes, size_t nodeBytes, int layout, size_t woopRows,
                                  std::vector<uint32_t>& out, int* outMaxDepth, std::string* err);
struct DevBuf;
cudaError_t convert_compact_to_wide4_device(const void* dNodes, size_t nodeBytes, int layout, size_t woopRows, DevBuf& wide, DevBuf& scratch,
                                            size_t* outWideBytes, int* outMaxDepth, int numSMs, cudaStream_t stream, int* launches, std::string* err);
cudaError_t launch_trace_wide4(const TraceLaunch& a, int* outNumLaunches);
cudaError_t launch_trace_mr(const TraceLaunch& a, int* outNumLaunches);
cudaError_t launch_trace_sw(const TraceLaunch& a, int* outNumLaunches);
KernelConfig trace_kernel_config(int kernel, int layout);

// ---- ray generation (nt_raygen.cu) ------------------------------------------------------------
struct PrimaryArgs {
    float4* rays; int* idToSlot; int* slotToID;
    float origin[3]; float n2w[16]; int w, h; float maxDist; unsigned seed;
};
cudaError_t launch_raygen_primary(const PrimaryArgs& a, const int* indexToPixel, cudaStream_t s);
struct AOArgs {
    float4* outRays; int* outIDToSlot; int* outSlotToID;
    const float4* inRays; const int4* inResults; const float* normals;
    int firstInputSlot, numInputRays, numSamples; float maxDist; unsigned seed;
    int order = 0;            // 0: the reference's slot order (slot = id); 1: direction-coherent order inside tiles of <= 1024 rays (nt_raygen.cu)
};
cudaError_t launch_raygen_ao(const AOArgs& a, cudaStream_t s);
struct ShadowArgs {
    float4* outRays; int* outIDToSlot; int* outSlotToID;
    const float4* inRays; const int4* inResults;
    int firstInputSlot, numInputRays, numSamples; float lightPos[3]; float lightRadius; unsigned seed;
};
cudaError_t launch_raygen_shadow(const ShadowArgs& a, cudaStream_t s);
cudaError_t launch_count_hits(const int4* results, int numRays, int* dCounter, cudaStream_t s);
cudaError_t launch_tri_normals(const float* verts, const int* tris, int numTris, float* out, cudaStream_t s);

// ---- ray Morton sort (nt_raysort.cu): in place on device buffers
cudaError_t ray_sort_device(float4* rays, int* idToSlot, int* slotToID, int numRays, cudaStream_t stream, int numSMs, int* outLaunches);

// ---- GPU LBVH / HLBVH builder (nt_build.cu) ---------------------------------------------------
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes);      // grow-only (reference: RayBuffer::resize never shrinks, RayBuffer.cpp:41-45)
    void release();
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct BuildParams {
    int builder;              // 0 = LBVH, 1 = HLBVH
    int hlbvhBits, leafSize; float epsilon;
    float lo[3], hi[3];
    int layout = Layout_Compact;   // Layout_Compact (byte offsets, < 1.98 GB of nodes) or Layout_Compact2 (offsets / 16)
    int collapse = 0;         // 0 = reference leaf rule (count <= leafSize), 1 = SAH-guided collapse
    int collapseMaxLeaf = 0;  // largest leaf the collapse may create (0 = leafSize)
    float collapseTriCost = 1.0f;   // SAH cost of a triangle test relative to a child-box test in the collapse (Platform default 1)
};
struct BuildOutput {          // device buffers owned by the context
    DevBuf* nodes; DevBuf* woop; DevBuf* triIndex;
    size_t nodeBytes, woopBytes, idxBytes;
    DevBuf* sortedKeys; DevBuf* sortedIdx;   // kept for nt_bvh_build_debug
};
// nt_shutdown: free the grow-only scratch of the builder / the ray sorter (it belongs to the device it was allocated on)
void release_build_scratch();
void release_sort_scratch();
// verts/tris are device pointers. Returns cudaSuccess and fills sizes; launches counted into *outLaunches.  `doneEvent` (optional) is
// recorded behind the last kernel, before the copy that brings the sizes back to the host.
cudaError_t build_bvh_device(const float* dVerts, int numVerts, const int* dTris, int numTris,
                             const BuildParams& p, BuildOutput& out, cudaStream_t stream,
                             int numSMs, int* outLaunches, std::string* err, cudaEvent_t doneEvent = nullptr);

// ---- basic CudaBVH layouts (nt_layout.cu): AOS/SOA buffers on the device -> Compact / Compact2 form in `out`
cudaError_t rescale_compact_links(int4* dNodes, size_t numNodes, int mulNum, int mulDen, cudaStream_t stream);
cudaError_t launch_sah(const float4* dNodes, size_t numNodes, const float4* dWoop, size_t woopRows, void* out4, cudaStream_t stream);
cudaError_t convert_basic_layout(int layout, const void* dNodes, size_t nodeBytes, const void* dWoop, size_t woopBytes,
                                 const int* dTriIndex, size_t idxBytes, int targetLayout, BuildOutput& out, DevBuf& scratch,
                                 cudaStream_t stream, int* outLaunches, std::string* err);

// ---- multi-GPU (nt_comm.cu): NCCL bound at run time
int comm_unique_id(void* out128);
int comm_init(int numRanks, int rank, const void* uniqueId128);
int comm_destroy();
bool comm_ready();
int comm_rank();
int comm_size();
int comm_broadcast_bytes(void* devPtr, size_t bytes, int root, cudaStream_t stream);
int comm_allreduce_f64(double* hostValues, int count, int op, cudaStream_t stream);

} // namespace nt
