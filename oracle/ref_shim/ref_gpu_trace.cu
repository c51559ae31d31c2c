// ORACLE — TEST INFRASTRUCTURE ONLY.  Runs one of the reference's OWN traversal kernels (src/rt/kernels/*.cu), compiled
// unmodified from /root/reference for sm_100a, so that the B200 kernel can be checked against — and timed beside — the
// recompiled reference on the same GPU, BVH and rays.  Built once per kernel file by `make -C oracle ref_gpu`:
//     nvcc ... -DREF_KERNEL_FILE='"kernels/fermi_speculative_while_while.cu"' [-DREF_PERSISTENT=1] -> oracle/_ref/libref_<name>.so
// Only shims live here, for language features removed from CUDA since the reference was written (CUDA 4.2):
//   * texture references (`texture<float4,1> t_x;` + tex1Dfetch) -> a __device__ struct holding the pointer, read with __ldg
//   * __any / __all / __ballot without a mask                      -> the _sync forms over __activemask()
// and the host side of CudaBVHTracer::traceBatch (CudaBVHTracer.cpp:88-168): bind the arrays, zero g_warpCounter for
// persistent kernels, size the grid from queryConfig()'s block shape, time the launch with events.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>

template <class T, int D> struct ref_texture { const T* ptr; };
template <class T> __device__ __forceinline__ T tex1Dfetch(const ref_texture<T, 1>& t, int i) { return __ldg(t.ptr + i); }
#define texture __device__ ref_texture
#define __any(p) __any_sync(__activemask(), (p))
#define __all(p) __all_sync(__activemask(), (p))
#define __ballot(p) __ballot_sync(__activemask(), (p))

#include REF_KERNEL_FILE

#undef texture

static char s_err[512] = "";
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(s_err, sizeof(s_err), "%s: %s", #call, cudaGetErrorString(e_)); return 1; } } while (0)

template <class T, int D> static cudaError_t bind(ref_texture<T, D>& sym, const void* p)
{
    ref_texture<T, D> h; h.ptr = (const T*)p;
    return cudaMemcpyToSymbol(sym, &h, sizeof(h));
}

extern "C" const char* ref_gpu_error(void) { return s_err; }

// cfg4 out: {bvhLayout, blockWidth, blockHeight, usePersistentThreads}.  desiredWarps <= 0: the reference's own choice
// (one warp per 32 rays, or the hard-coded 720 warps for persistent kernels, CudaBVHTracer.cpp:151-156).
extern "C" int ref_gpu_trace(const void* dRays, void* dResults, int numRays, int anyHit, const void* dNodes, const void* dWoop,
                             const void* dTriIndex, int desiredWarps, int repeats, float* outBestMs, int* cfg4)
{
    KernelConfig zero; memset(&zero, 0, sizeof(zero));
    CK(cudaMemcpyToSymbol(g_config, &zero, sizeof(zero)));
    queryConfig<<<1, 1>>>();
    KernelConfig cfg;
    CK(cudaMemcpyFromSymbol(&cfg, g_config, sizeof(cfg)));
    if (cfg4) { cfg4[0] = cfg.bvhLayout; cfg4[1] = cfg.blockWidth; cfg4[2] = cfg.blockHeight; cfg4[3] = cfg.usePersistentThreads; }
    CK(bind(t_rays, dRays)); CK(bind(t_nodesA, dNodes)); CK(bind(t_nodesB, dNodes)); CK(bind(t_nodesC, dNodes)); CK(bind(t_nodesD, dNodes));
    CK(bind(t_trisA, dWoop)); CK(bind(t_trisB, dWoop)); CK(bind(t_trisC, dWoop)); CK(bind(t_triIndices, dTriIndex));
    int warps = (numRays + 31) / 32;
    if (cfg.usePersistentThreads != 0) warps = 720;
    if (desiredWarps > 0) warps = desiredWarps;
    const int blockWarps = (cfg.blockWidth * cfg.blockHeight + 31) / 32;
    const int numBlocks = (warps + blockWarps - 1) / blockWarps;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1.0e30f;
    for (int r = 0; r < (repeats > 0 ? repeats : 1); r++) {
#ifdef REF_PERSISTENT
        int z = 0;
        CK(cudaMemcpyToSymbol(g_warpCounter, &z, sizeof(int)));
#endif
        CK(cudaEventRecord(a));
        trace_bvh<<<numBlocks, dim3(cfg.blockWidth, cfg.blockHeight)>>>(numRays, anyHit != 0, (float4*)dRays, (int4*)dResults, (float4*)dNodes, (float4*)dNodes,
                                                                         (float4*)dNodes, (float4*)dNodes, (float4*)dWoop, (float4*)dWoop, (float4*)dWoop, (int*)dTriIndex);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    if (outBestMs) *outBestMs = best;
    return 0;
}
