// ORACLE — TEST INFRASTRUCTURE ONLY.  Runs the reference's OWN GPU builder kernels (src/rt/bvh/HLBVH/emitTreeKernel.cu:
// calcMorton, calcWoopKernel, emitTreeKernel + createLeaf, calcAABB; radixSort.cu: thrust sort_by_key), compiled unmodified
// from /root/reference for sm_100a.  The host side is the launch sequence of HLBVHBuilder::buildLBVH
// (HLBVHBuilder.cpp:451-593): calcMortonAndSort (:67-96), the Woop pass, initMemory (:772-784), buildBottomLevel (:319-376: one
// emitTreeKernel launch per level, queue head read back in between) and calcAABB (:408-447: one launch per level, bottom-up).
// The kernels use warp-synchronous shared-memory scans without __syncwarp (CUDA 4.2 era); they are run as they are.
// Built twice by `make -C oracle ref_gpu`: with the reference's -use_fast_math (HLBVHBuilder.cpp:462) and with IEEE flags.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <thrust/functional.h>
#include <thrust/sort.h>
#include <thrust/unique.h>
#include <thrust/iterator/discard_iterator.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/device_ptr.h>
#include <thrust/host_vector.h>
#if !defined(REF_HAVE_THRUST_BINARY_FUNCTION)
namespace thrust { template <class A, class B, class R> struct ref_binary_function { typedef A first_argument_type; typedef B second_argument_type; typedef R result_type; }; }
#define binary_function ref_binary_function        /* removed from Thrust 2.x; radixSort.cu only inherits the typedefs */
#endif
#include "bvh/HLBVH/radixSort.cu"
#undef binary_function
#include "bvh/HLBVH/emitTreeKernel.cu"

static char s_err[512] = "";
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(s_err, sizeof(s_err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return 1; } } while (0)
extern "C" const char* ref_gpu_error(void) { return s_err; }

template <class T> static cudaError_t setSym(const T& sym, const void* value) { return cudaMemcpyToSymbol(sym, value, sizeof(T)); }
static unsigned gridFor(long long n) { return (unsigned)((n + BLOCK_SIZE - 1) / BLOCK_SIZE); }

struct RefBuild { void *nodes, *woop, *idx, *morton, *order; size_t nodeBytes, woopBytes, idxBytes; int numTris; };
static RefBuild g_last = {};

static void freeLast()
{
    cudaFree(g_last.nodes); cudaFree(g_last.woop); cudaFree(g_last.idx); cudaFree(g_last.morton); cudaFree(g_last.order);
    g_last = RefBuild();
}

// HLBVHBuilder::buildLBVH.  dVerts / dTris: device pointers (float3 per vertex, int3 per triangle).  lo / hi: Scene::getBBox.
extern "C" int ref_lbvh_build(const void* dVerts, const void* dTris, int triCnt, const float* lo, const float* hi, int leafSize, float epsilon, int* outNodesLeaves)
{
    freeLast();
    const int n_bits = 30;
    CK(setSym(c_leafSize, &leafSize)); CK(setSym(c_epsilon, &epsilon));
    CUdeviceptr tris = (CUdeviceptr)dTris;
    CK(setSym(g_tris, &tris)); CK(setSym(g_verts, &dVerts));
    void *morton, *order, *inWoop, *outWoop, *outIdx, *nodes, *q[2];
    CK(cudaMalloc(&morton, (size_t)triCnt * 4)); CK(cudaMalloc(&order, (size_t)triCnt * 4));
    CK(setSym(g_inTriMem, &morton)); CK(setSym(g_inTriIdxMem, &order));
    // calcMortonAndSort: step = (sceneMax - sceneMin) / 1024.0f per component, in fp32
    float step[3];
    for (int i = 0; i < 3; i++) step[i] = (hi[i] - lo[i]) / 1024.0f;
    calcMorton<<<gridFor(triCnt), BLOCK_SIZE>>>((uint)triCnt, lo[0], lo[1], lo[2], step[0], step[1], step[2]);
    CK(cudaDeviceSynchronize());
    radixSortCuda((CUdeviceptr)morton, (CUdeviceptr)order, triCnt);
    CK(cudaDeviceSynchronize());
    CK(cudaMalloc(&inWoop, (size_t)triCnt * 3 * 16));
    CK(setSym(g_inWoopMem, &inWoop));
    calcWoopKernel<<<gridFor(triCnt), BLOCK_SIZE>>>((uint)triCnt);
    CK(cudaDeviceSynchronize());
    CK(cudaMalloc(&outWoop, (size_t)triCnt * 4 * 16)); CK(cudaMalloc(&outIdx, (size_t)triCnt * 4 * 4));
    CK(cudaMemset(outWoop, 0, (size_t)triCnt * 4 * 16)); CK(cudaMemset(outIdx, 0, (size_t)triCnt * 4 * 4));
    CK(setSym(g_outWoopMem, &outWoop));
    CUdeviceptr outIdxPtr = (CUdeviceptr)outIdx;
    CK(setSym(g_outIdxMem, &outIdxPtr));
    unsigned long long zero64 = 0;
    CK(setSym(g_leafsPtr, &zero64));
    // initMemory(q0, q1, min(2, leafSize))
    const long long size = 2LL * (triCnt / (leafSize < 2 ? leafSize : 2));
    CK(cudaMalloc(&nodes, (size_t)size * 64)); CK(cudaMemset(nodes, 0, (size_t)size * 64));
    CUdeviceptr nodesPtr = (CUdeviceptr)nodes;
    CK(setSym(g_outNodes, &nodesPtr));
    CK(cudaMalloc(&q[0], (size_t)size * 12)); CK(cudaMalloc(&q[1], (size_t)size * 12));
    int root[3] = {0, 0, triCnt};
    CK(cudaMemcpy(q[0], root, 12, cudaMemcpyHostToDevice));
    // buildBottomLevel(&q0, &q1, nodeWritten = 1, nodeCreated = 1, bOfs = 0, n_bits)
    std::vector<unsigned> lvlNodes(1, 1u);
    unsigned nodeWritten = 1, nodeCreated = 1, level = 0;
    int in = 0;
    while (level < (unsigned)n_bits && nodeCreated > 0) {
        CK(setSym(g_inQueueMem, &q[in])); CK(setSym(g_outQueueMem, &q[in ^ 1]));
        int zero = 0;
        CK(setSym(g_inQueuePtr, &zero)); CK(setSym(g_outQueuePtr, &zero));
        emitTreeKernel<<<gridFor(nodeCreated), BLOCK_SIZE>>>(n_bits - (int)(level + 1), nodeCreated, (int)nodeWritten);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpyFromSymbol(&nodeCreated, g_outQueuePtr, 4));
        lvlNodes.push_back(nodeCreated);
        nodeWritten += nodeCreated;
        if (lvlNodes.back() == 0) lvlNodes.pop_back();
        in ^= 1;
        level++;
    }
    unsigned long long leafsPtr = 0;
    CK(cudaMemcpyFromSymbol(&leafsPtr, g_leafsPtr, 8));
    const unsigned leafs = (unsigned)(leafsPtr & 0xFFFFFFFFu);
    // calcAABB(nodeWritten): bottom-up, one launch per level
    unsigned nw = nodeWritten;
    for (int lvl = (int)lvlNodes.size() - 1; lvl >= 0; lvl--) {
        nw -= lvlNodes[lvl];
        calcAABB<<<gridFor(lvlNodes[lvl]), BLOCK_SIZE>>>((int)nw, (int)lvlNodes[lvl]);
        CK(cudaDeviceSynchronize());
    }
    cudaFree(q[0]); cudaFree(q[1]); cudaFree(inWoop);
    g_last.nodes = nodes; g_last.woop = outWoop; g_last.idx = outIdx; g_last.morton = morton; g_last.order = order;
    g_last.nodeBytes = (size_t)nodeWritten * 64; g_last.woopBytes = (size_t)triCnt * 48 + (size_t)leafs * 16; g_last.idxBytes = (size_t)triCnt * 12 + (size_t)leafs * 4;
    g_last.numTris = triCnt;
    if (outNodesLeaves) { outNodesLeaves[0] = (int)nodeWritten; outNodesLeaves[1] = (int)leafs; outNodesLeaves[2] = (int)level; }
    return 0;
}

extern "C" int ref_build_sizes(long long* sizes3) { sizes3[0] = (long long)g_last.nodeBytes; sizes3[1] = (long long)g_last.woopBytes; sizes3[2] = (long long)g_last.idxBytes; return g_last.nodes ? 0 : 1; }

// host copies of the last build: nodes, woop, triIndex, sorted Morton codes, sorted triangle order
extern "C" int ref_build_download(void* nodes, void* woop, void* idx, void* morton, void* order)
{
    if (!g_last.nodes) { snprintf(s_err, sizeof(s_err), "no build"); return 1; }
    if (nodes) CK(cudaMemcpy(nodes, g_last.nodes, g_last.nodeBytes, cudaMemcpyDeviceToHost));
    if (woop) CK(cudaMemcpy(woop, g_last.woop, g_last.woopBytes, cudaMemcpyDeviceToHost));
    if (idx) CK(cudaMemcpy(idx, g_last.idx, g_last.idxBytes, cudaMemcpyDeviceToHost));
    if (morton) CK(cudaMemcpy(morton, g_last.morton, (size_t)g_last.numTris * 4, cudaMemcpyDeviceToHost));
    if (order) CK(cudaMemcpy(order, g_last.order, (size_t)g_last.numTris * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// calcMorton alone, unsorted: codes[i] for triangle i
extern "C" int ref_morton(const void* dVerts, const void* dTris, int triCnt, const float* lo, const float* hi, void* hostCodes)
{
    CUdeviceptr tris = (CUdeviceptr)dTris;
    CK(setSym(g_tris, &tris)); CK(setSym(g_verts, &dVerts));
    void *morton, *order;
    CK(cudaMalloc(&morton, (size_t)triCnt * 4)); CK(cudaMalloc(&order, (size_t)triCnt * 4));
    CK(setSym(g_inTriMem, &morton)); CK(setSym(g_inTriIdxMem, &order));
    float step[3];
    for (int i = 0; i < 3; i++) step[i] = (hi[i] - lo[i]) / 1024.0f;
    calcMorton<<<gridFor(triCnt), BLOCK_SIZE>>>((uint)triCnt, lo[0], lo[1], lo[2], step[0], step[1], step[2]);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hostCodes, morton, (size_t)triCnt * 4, cudaMemcpyDeviceToHost));
    cudaFree(morton); cudaFree(order);
    return 0;
}
