// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  Parity PINNED ON THE GPU BOX: the reference's
// HLBVH builder is device code (emitTreeKernel.cu) and cannot run in the build container, but it compiles for sm_100a
// (oracle/ref_gpu.py, `make -C oracle ref_gpu`) and tests/test_gpu_reference_kernels.py runs it on the B200 beside the
// product builder: Morton codes, sorted order, LBVH and HLBVH trees (canonical form), child boxes and Woop rows are
// bit-identical to the IEEE build of the reference kernels; the product builder is in turn bit-identical to this
// restatement (tests/test_gpu_build.py), which closes the triangle.  In the container (no GPU) this file is constrained by
// tests/test_oracle_lbvh.py only.
//
// CPU restatement of the reference GPU LBVH / HLBVH builder, executed with a *serial
// schedule* (threads in queue order), which is one of the schedules the reference's
// atomics allow.  Restated from:
//   src/rt/bvh/HLBVH/emitTreeKernel.cu:647-691   calcMorton / spread
//   src/rt/bvh/HLBVH/radixSort.cu:22-46           thrust::sort_by_key contract (stable asc.)
//   src/rt/bvh/HLBVH/emitTreeKernel.cu:574-645   calcWoop (3x3 form)
//   src/rt/bvh/HLBVH/emitTreeKernel.cu:170-381   createLeaf / emitTreeKernel
//   src/rt/bvh/HLBVH/emitTreeKernel.cu:383-562   calcLeaf / calcAABB
//   src/rt/bvh/HLBVH/emitTreeKernel.cu:699-1027  initBins / fillBins / findSplit / distribute
//   src/rt/bvh/HLBVH/HLBVHBuilder.cpp:67-750     host driver
#pragma once
#include "orc_bvh.hpp"

namespace orc {

struct HLBVHParams { bool hlbvh = true; int hlbvhBits = 4; int leafSize = 8; float epsilon = 0.001f; };

void morton_codes(const Scene& sc, V3 lo, V3 hi, uint32_t* codes);
void sort_pairs_stable(uint32_t* keys, int32_t* idx, int n);
void calc_woop_gpu(V3 v0, V3 v1, V3 v2, float out[12]);

struct LBVHResult {
    CompactBVH bvh;
    std::vector<uint32_t> sortedKeys;
    std::vector<int32_t> sortedIdx;
    std::vector<int> levelNodes;    // lvlNodes of the reference (nodes created per level)
    int numNodes = 0, numLeaves = 0;
    int numClusters = 0;            // HLBVH only
    float buildSeconds = 0.0f;
};

void build_lbvh(const Scene& sc, V3 lo, V3 hi, const HLBVHParams& p, LBVHResult& out);

// Canonical (numbering-independent) serialisation of a Compact tree, preorder, child 0 first:
//   tokens: inner -> (leftTris, rightTris, word14) ; boxes: 12 floats per inner node (node words 0-11)
//   leaf tri ids appended to `tris` in traversal order; `leafSizes` one entry per leaf.
struct Canonical {
    std::vector<int32_t> inner;      // 3 per inner node
    std::vector<float> boxes;        // 12 per inner node
    std::vector<int32_t> leafSizes;
    std::vector<int32_t> tris;
    std::vector<float> woop;         // 12 per leaf triangle, traversal order
};
void canonicalize(const int32_t* nodes, const int32_t* woop, const int32_t* triIndex, Canonical& out);

} // namespace orc
