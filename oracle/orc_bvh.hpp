// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  Parity PINNED against the
// compiled reference (oracle/_ref, tests/test_reference_pin.py): bit-identical SAH, trees, Woop data and trace results.
//
// CPU pointer-tree BVH (array-backed), the reference's two CPU builders, its SAH metric and
// its two CPU tracers, restated from:
//   src/rt/Util.cpp:34-127            RayBox / RayTriangle / RayTriangleWoop
//   src/rt/bvh/Platform.hpp:53-139    SAH cost model
//   src/rt/bvh/SAHBVHBuilder.cpp:51-254, SplitBVHBuilder.cpp:49-393, base/Sort.cpp:62-160
//   src/rt/bvh/BVHNode.cpp:79-94      SAH metric
//   src/rt/bvh/BVH.cpp:90-186         BVH::trace
//   src/rt/cuda/CudaBVH.cpp:579-687   createCompact / woopifyTri (no shuffle)
//   src/rt/cuda/CudaBVH.cpp:698-784,1083-1126,1183-1225  flat Compact trace
#pragma once
#include "orc_math.hpp"

namespace orc {

// ---- intersections (Util.cpp) ---------------------------------------------------------
struct Span { float tmin, tmax; };
Span  ray_box(const AABB& box, const Ray& ray);                       // Util.cpp:34-46
// returns t (F32_MAX on miss); u,v optional
float ray_triangle(V3 v0, V3 v1, V3 v2, const Ray& ray, float* u = nullptr, float* v = nullptr);   // Util.cpp:50-94
float ray_triangle_woop(const float* zpleq, const float* upleq, const float* vpleq, const Ray& ray,
                        float* u = nullptr, float* v = nullptr);                                       // Util.cpp:99-127

// ---- platform (Platform.hpp) -----------------------------------------------------------
struct Platform {
    float nodeCost = 1.0f, triCost = 1.0f;
    int nodeBatch = 1, triBatch = 1;
    int minLeaf = 1, maxLeaf = 0x7FFFFFF;
    int roundTri(int n) const { return ((n + triBatch - 1) / triBatch) * triBatch; }
    int roundNode(int n) const { return ((n + nodeBatch - 1) / nodeBatch) * nodeBatch; }
    float triangleCost(int n) const { return roundTri(n) * triCost; }
    float nodeCostN(int n) const { return roundNode(n) * nodeCost; }
    float cost(int nChildren, int nTris) const { return nodeCostN(nChildren) + triangleCost(nTris); }
};

// ---- tree ------------------------------------------------------------------------------
struct Node {
    AABB bounds;
    int child[2] = {-1, -1};   // node indices, -1 for leaves
    int lo = 0, hi = 0;        // leaf: range in triIndices
    int axis = 0, splitType = 0;  // SplitInfo: axis | type<<2  (BVHNode.hpp:87)
    bool leaf = false;
};

struct Scene {
    const V3* verts = nullptr; int numVerts = 0;
    const int32_t* tris = nullptr; int numTris = 0;   // 3 ints per triangle
    V3 v(int tri, int k) const { return verts[tris[3 * tri + k]]; }
};

struct BVH {
    Scene scene;
    Platform platform;
    std::vector<Node> nodes;
    std::vector<int32_t> triIndices;
    int root = -1;
    int numDuplicates = 0;
    float buildSeconds = 0.0f;
};

enum BuilderKind { BUILDER_SAH = 0, BUILDER_SPLIT = 1 };

void build_bvh(BVH& bvh, BuilderKind kind, float splitAlpha);

struct TreeStats { float sah; int numInner, numLeaf, numTris, maxDepth; };
TreeStats tree_stats(const BVH& bvh);                                  // BVHNode.cpp:36-94

// per-ray counters: [0]=inner nodes whose two child boxes were tested, [1]=triangles tested, [2]=leaves entered
void trace_tree(const BVH& bvh, const Ray* rays, RayResult* results, int n, bool needClosestHit,
                uint32_t* counters /* n*3 or null */, int nthreads);    // BVH.cpp:90-186

// ---- flat Compact layout (CudaBVH.cpp) ---------------------------------------------------
struct CompactBVH {
    std::vector<int32_t> nodes;     // 16 words per inner node
    std::vector<int32_t> woop;      // 4 words per float4
    std::vector<int32_t> triIndex;  // one int per woop float4
};
void woopify_tri(V3 v0, V3 v1, V3 v2, float out[12]);                  // CudaBVH.cpp:667-687
void create_compact(const BVH& bvh, CompactBVH& out, int nodeOffsetSizeDiv = 1);   // CudaBVH.cpp:579-664
// layout: 0 AOS_AOS, 1 AOS_SOA, 2 SOA_AOS, 3 SOA_SOA (CudaTracerKernels.hpp:54-57).  Buffers 4096-byte padded, padding zero.
void create_basic(const BVH& bvh, int layout, CompactBVH& out);                    // CudaBVH.cpp:453-575

void trace_compact(const int32_t* nodes, const int32_t* woop, const int32_t* triIndex,
                   const Ray* rays, RayResult* results, int n, bool needClosestHit,
                   uint32_t* counters, int nthreads);                   // CudaBVH.cpp:213-302,698-784

// brute force over all triangles (not in the reference; independent cross-check of both tracers)
void trace_brute(const Scene& scene, const Ray* rays, RayResult* results, int n, bool needClosestHit, int nthreads);

// SAH metric evaluated on a flat Compact tree with the BVHNode.cpp:79-94 formula
// (used to put LBVH/HLBVH outputs and SplitBVH on the same scale).  Box of the root is the
// union of its two child boxes.
float compact_sah(const int32_t* nodes, const int32_t* woop, const Platform& p,
                  int* numInner, int* numLeaf, int* numTris, int* maxDepth);

} // namespace orc
