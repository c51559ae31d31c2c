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

// the reference's own accounting of build time: m_gpuTime = sum of launchTimed() kernel times + the thrust calls
// (HLBVHBuilder.cpp:84-92, 350, 433); host readbacks between levels are not in it
static float s_gpuMs = 0.0f;
static cudaEvent_t s_ev0 = NULL, s_ev1 = NULL;
static void tick() { if (!s_ev0) { cudaEventCreate(&s_ev0); cudaEventCreate(&s_ev1); } cudaEventRecord(s_ev0); }
static void tock() { cudaEventRecord(s_ev1); cudaEventSynchronize(s_ev1); float ms = 0.0f; cudaEventElapsedTime(&ms, s_ev0, s_ev1); s_gpuMs += ms; }
extern "C" float ref_build_gpu_ms(void) { return s_gpuMs; }

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
    s_gpuMs = 0.0f;
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
    tick(); calcMorton<<<gridFor(triCnt), BLOCK_SIZE>>>((uint)triCnt, lo[0], lo[1], lo[2], step[0], step[1], step[2]); tock();
    CK(cudaDeviceSynchronize());
    tick(); radixSortCuda((CUdeviceptr)morton, (CUdeviceptr)order, triCnt); tock();
    CK(cudaDeviceSynchronize());
    CK(cudaMalloc(&inWoop, (size_t)triCnt * 3 * 16));
    CK(setSym(g_inWoopMem, &inWoop));
    tick(); calcWoopKernel<<<gridFor(triCnt), BLOCK_SIZE>>>((uint)triCnt); tock();
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
        tick(); emitTreeKernel<<<gridFor(nodeCreated), BLOCK_SIZE>>>(n_bits - (int)(level + 1), nodeCreated, (int)nodeWritten); tock();
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
        tick(); calcAABB<<<gridFor(lvlNodes[lvl]), BLOCK_SIZE>>>((int)nw, (int)lvlNodes[lvl]); tock();
        CK(cudaDeviceSynchronize());
    }
    cudaFree(q[0]); cudaFree(q[1]); cudaFree(inWoop);
    g_last.nodes = nodes; g_last.woop = outWoop; g_last.idx = outIdx; g_last.morton = morton; g_last.order = order;
    g_last.nodeBytes = (size_t)nodeWritten * 64; g_last.woopBytes = (size_t)triCnt * 48 + (size_t)leafs * 16; g_last.idxBytes = (size_t)triCnt * 12 + (size_t)leafs * 4;
    g_last.numTris = triCnt;
    if (outNodesLeaves) { outNodesLeaves[0] = (int)nodeWritten; outNodesLeaves[1] = (int)leafs; outNodesLeaves[2] = (int)level; }
    return 0;
}

// HLBVHBuilder::buildHLBVH (HLBVHBuilder.cpp:595-745): Morton + sort, clusters of equal high bits (createClustersC :98-156,
// thrust unique_by_key_copy), binned-SAH top level over the clusters (buildTopLevel :158-317: initBins / fillBins / findSplit /
// distribute per level, counters read back in between), LBVH below every cluster (buildBottomLevel), calcAABB.
extern "C" int ref_hlbvh_build(const void* dVerts, const void* dTris, int triCnt, const float* lo, const float* hi, int hlbvhBits, int leafSize, float epsilon,
                               int* outNodesLeaves)
{
    freeLast();
    s_gpuMs = 0.0f;
    const int n = 10, m = n - hlbvhBits, d = 3 * (n - m), n_bits = 3 * n;
    CK(setSym(c_leafSize, &leafSize)); CK(setSym(c_epsilon, &epsilon));
    CUdeviceptr tris = (CUdeviceptr)dTris;
    CK(setSym(g_tris, &tris)); CK(setSym(g_verts, &dVerts));
    void *morton, *order, *inWoop, *outWoop, *outIdx, *nodes, *q[2];
    CK(cudaMalloc(&morton, (size_t)triCnt * 4)); CK(cudaMalloc(&order, (size_t)triCnt * 4));
    CK(setSym(g_inTriMem, &morton)); CK(setSym(g_inTriIdxMem, &order));
    float step[3];
    for (int i = 0; i < 3; i++) step[i] = (hi[i] - lo[i]) / 1024.0f;
    tick(); calcMorton<<<gridFor(triCnt), BLOCK_SIZE>>>((uint)triCnt, lo[0], lo[1], lo[2], step[0], step[1], step[2]); tock();
    CK(cudaDeviceSynchronize());
    tick(); radixSortCuda((CUdeviceptr)morton, (CUdeviceptr)order, triCnt); tock();
    CK(cudaDeviceSynchronize());
    // createClustersC
    void *clusters, *clsBB, *clsBin, *clsSplit;
    int cluster_cnt = 0;
    CK(cudaMalloc(&clusters, (size_t)(triCnt + 1) * 4));
    tick(); createClusters((CUdeviceptr)morton, triCnt, d, (CUdeviceptr)clusters, cluster_cnt); tock();
    CK(cudaDeviceSynchronize());
    CK(setSym(g_clsStart, &clusters));
    CK(cudaMalloc(&clsBB, (size_t)cluster_cnt * 24)); CK(cudaMalloc(&clsBin, (size_t)cluster_cnt * 12)); CK(cudaMalloc(&clsSplit, (size_t)cluster_cnt * 4));
    CK(cudaMemset(clsSplit, 0, (size_t)cluster_cnt * 4));
    CK(setSym(g_clsAABB, &clsBB)); CK(setSym(g_clsBinId, &clsBin)); CK(setSym(g_clsSplitId, &clsSplit));
    tick(); initClusterAABB<<<gridFor(cluster_cnt), BLOCK_SIZE>>>(cluster_cnt); tock();
    tick(); clusterAABB<<<NUM_BLOCKS, BLOCK_SIZE>>>(cluster_cnt, triCnt); tock();
    CK(cudaDeviceSynchronize());
    // Woop pass, output buffers
    CK(cudaMalloc(&inWoop, (size_t)triCnt * 3 * 16));
    CK(setSym(g_inWoopMem, &inWoop));
    tick(); calcWoopKernel<<<gridFor(triCnt), BLOCK_SIZE>>>((uint)triCnt); tock();
    CK(cudaDeviceSynchronize());
    CK(cudaMalloc(&outWoop, (size_t)triCnt * 4 * 16)); CK(cudaMalloc(&outIdx, (size_t)triCnt * 4 * 4));
    CK(cudaMemset(outWoop, 0, (size_t)triCnt * 4 * 16)); CK(cudaMemset(outIdx, 0, (size_t)triCnt * 4 * 4));
    CK(setSym(g_outWoopMem, &outWoop));
    CUdeviceptr outIdxPtr = (CUdeviceptr)outIdx;
    CK(setSym(g_outIdxMem, &outIdxPtr));
    unsigned long long zero64 = 0;
    CK(setSym(g_leafsPtr, &zero64));
    std::vector<unsigned> lvlNodes(1, 1u);
    const int leafArg = (d == 0) ? 1 : (leafSize < 2 ? leafSize : 2);
    const long long size = 2LL * (triCnt / leafArg);
    CK(cudaMalloc(&nodes, (size_t)size * 64)); CK(cudaMemset(nodes, 0, (size_t)size * 64));
    CUdeviceptr nodesPtr = (CUdeviceptr)nodes;
    CK(setSym(g_outNodes, &nodesPtr));
    CK(cudaMalloc(&q[0], (size_t)size * 12)); CK(cudaMalloc(&q[1], (size_t)size * 12));
    // buildTopLevel
    unsigned sahCreated = 1, sahWritten = 1, sahTerminated = 0, oldTerminated = 0;
    const long long bufferSize = 2LL * cluster_cnt;
    void *qbb[2], *qcls[2], *qid[2], *qplane[2], *qchild[2], *binBB, *binCnt;
    for (int i = 0; i < 2; i++) {
        CK(cudaMalloc(&qbb[i], (size_t)bufferSize * 24)); CK(cudaMalloc(&qcls[i], (size_t)bufferSize * 4)); CK(cudaMalloc(&qid[i], (size_t)bufferSize * 4));
        CK(cudaMalloc(&qplane[i], (size_t)bufferSize * 4)); CK(cudaMalloc(&qchild[i], (size_t)bufferSize * 4));
        CK(cudaMemset(qbb[i], 0, (size_t)bufferSize * 24)); CK(cudaMemset(qcls[i], 0, (size_t)bufferSize * 4)); CK(cudaMemset(qid[i], 0, (size_t)bufferSize * 4));
        CK(cudaMemset(qplane[i], 0, (size_t)bufferSize * 4)); CK(cudaMemset(qchild[i], 0, (size_t)bufferSize * 4));
    }
    {
        int id0 = 0, cls0 = cluster_cnt, child0 = -1;
        float bb0[6] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
        CK(cudaMemcpy(qid[0], &id0, 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(qcls[0], &cls0, 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(qchild[0], &child0, 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(qbb[0], bb0, 24, cudaMemcpyHostToDevice));
    }
    CK(setSym(g_ooq, &q[0]));
    unsigned zero = 0;
    CK(setSym(g_oofs, &zero));
    CK(cudaMalloc(&binBB, (size_t)24 * BIN_CNT * 3 * bufferSize)); CK(cudaMalloc(&binCnt, (size_t)4 * BIN_CNT * 3 * bufferSize));
    CUdeviceptr binBBPtr = (CUdeviceptr)binBB;
    CK(setSym(g_binAABB, &binBBPtr)); CK(setSym(g_binCnt, &binCnt));
    int in = 0, guard = 0;
    while (sahCreated > 0) {
        if (++guard > 4096) { snprintf(s_err, sizeof(s_err), "top-level SAH did not terminate"); return 1; }
        CUdeviceptr pin = (CUdeviceptr)qbb[in], pout = (CUdeviceptr)qbb[in ^ 1];
        CK(setSym(g_qsiAABB, &pin)); CK(setSym(g_qsiCnt, &qcls[in])); CK(setSym(g_qsiId, &qid[in])); CK(setSym(g_qsiPlane, &qplane[in])); CK(setSym(g_qsiChildId, &qchild[in]));
        CK(setSym(g_qsoAABB, &pout)); CK(setSym(g_qsoCnt, &qcls[in ^ 1])); CK(setSym(g_qsoId, &qid[in ^ 1])); CK(setSym(g_qsoPlane, &qplane[in ^ 1])); CK(setSym(g_qsoChildId, &qchild[in ^ 1]));
        tick(); initBins<<<gridFor((long long)sahCreated * BIN_CNT * 3), BLOCK_SIZE>>>(sahCreated * BIN_CNT * 3); tock();
        tick(); fillBins<<<gridFor(cluster_cnt), BLOCK_SIZE>>>((uint)cluster_cnt); tock();
        CK(setSym(g_sahCreated, &zero));
        tick(); findSplit<<<gridFor(sahCreated), BLOCK_SIZE>>>(sahCreated, sahWritten); tock();
        tick(); distribute<<<gridFor(cluster_cnt), BLOCK_SIZE>>>((uint)cluster_cnt, (int)sahWritten); tock();
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpyFromSymbol(&sahTerminated, g_oofs, 4));
        CK(cudaMemcpyFromSymbol(&sahCreated, g_sahCreated, 4));
        const int terminated = (int)(sahTerminated - oldTerminated);
        oldTerminated = sahTerminated;
        if (sahCreated != 0) lvlNodes.push_back(sahCreated);
        sahWritten += sahCreated;
        sahCreated -= (unsigned)terminated;
        in ^= 1;
    }
    unsigned nodeWritten = sahWritten, nodeCreated = 0, level = 0;
    CK(cudaMemcpyFromSymbol(&nodeCreated, g_oofs, 4));
    const int numTopNodes = (int)nodeWritten, numLbvhRoots = (int)nodeCreated;
    // buildBottomLevel(&q0, &q1, nodeWritten, nodeCreated, 3 * m, n_bits)
    if (d != 0) {
        const int bit_ofs = 3 * m;
        int qin = 0;
        while (level < (unsigned)(n_bits - bit_ofs) && nodeCreated > 0) {
            CK(setSym(g_inQueueMem, &q[qin])); CK(setSym(g_outQueueMem, &q[qin ^ 1]));
            CK(setSym(g_inQueuePtr, &zero)); CK(setSym(g_outQueuePtr, &zero));
            tick(); emitTreeKernel<<<gridFor(nodeCreated), BLOCK_SIZE>>>(n_bits - (int)(level + 1 + bit_ofs), nodeCreated, (int)nodeWritten); tock();
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpyFromSymbol(&nodeCreated, g_outQueuePtr, 4));
            lvlNodes.push_back(nodeCreated);
            nodeWritten += nodeCreated;
            if (lvlNodes.back() == 0) lvlNodes.pop_back();
            qin ^= 1;
            level++;
        }
    }
    unsigned long long leafsPtr = 0;
    CK(cudaMemcpyFromSymbol(&leafsPtr, g_leafsPtr, 8));
    const unsigned leafs = (unsigned)(leafsPtr & 0xFFFFFFFFu);
    unsigned nw = nodeWritten;
    for (int lvl = (int)lvlNodes.size() - 1; lvl >= 0; lvl--) {
        nw -= lvlNodes[lvl];
        tick(); calcAABB<<<gridFor(lvlNodes[lvl]), BLOCK_SIZE>>>((int)nw, (int)lvlNodes[lvl]); tock();
        CK(cudaDeviceSynchronize());
    }
    for (int i = 0; i < 2; i++) { cudaFree(qbb[i]); cudaFree(qcls[i]); cudaFree(qid[i]); cudaFree(qplane[i]); cudaFree(qchild[i]); cudaFree(q[i]); }
    cudaFree(binBB); cudaFree(binCnt); cudaFree(clusters); cudaFree(clsBB); cudaFree(clsBin); cudaFree(clsSplit); cudaFree(inWoop);
    g_last.nodes = nodes; g_last.woop = outWoop; g_last.idx = outIdx; g_last.morton = morton; g_last.order = order;
    g_last.nodeBytes = (size_t)nodeWritten * 64; g_last.woopBytes = (size_t)triCnt * 48 + (size_t)leafs * 16; g_last.idxBytes = (size_t)triCnt * 12 + (size_t)leafs * 4;
    g_last.numTris = triCnt;
    if (outNodesLeaves) { outNodesLeaves[0] = (int)nodeWritten; outNodesLeaves[1] = (int)leafs; outNodesLeaves[2] = (int)level; outNodesLeaves[3] = cluster_cnt; outNodesLeaves[4] = numTopNodes; outNodesLeaves[5] = numLbvhRoots; }
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
    tick(); calcMorton<<<gridFor(triCnt), BLOCK_SIZE>>>((uint)triCnt, lo[0], lo[1], lo[2], step[0], step[1], step[2]); tock();
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hostCodes, morton, (size_t)triCnt * 4, cudaMemcpyDeviceToHost));
    cudaFree(morton); cudaFree(order);
    return 0;
}
