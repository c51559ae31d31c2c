// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  Parity status per module: see orc_math.hpp.
// Plain-C entry points for the ctypes loader in oracle/__init__.py.
#include "orc_bvh.hpp"
#include "orc_wide.hpp"
#include "orc_lbvh.hpp"
#include "orc_raygen.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {
struct BvhHandle {
    std::vector<V3> verts;
    std::vector<int32_t> tris;
    BVH bvh;
    CompactBVH compact;
    bool haveCompact = false;
};
struct LbvhHandle {
    std::vector<V3> verts;
    std::vector<int32_t> tris;
    LBVHResult res;
    Canonical canon;
    bool haveCanon = false;
};
Scene make_scene(const std::vector<V3>& v, const std::vector<int32_t>& t)
{
    Scene s; s.verts = v.data(); s.numVerts = (int)v.size(); s.tris = t.data(); s.numTris = (int)t.size() / 3; return s;
}
}

extern "C" {

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---------------- CPU BVH builders + tracers ------------------------------------------------
void* orc_bvh_build(const float* vtx, int nv, const int32_t* tri, int nt, int builder, int minLeaf, int maxLeaf, float splitAlpha)
{
    BvhHandle* h = new BvhHandle;
    h->verts.resize(nv);
    std::memcpy((void*)h->verts.data(), vtx, sizeof(float) * 3 * (size_t)nv);
    h->tris.assign(tri, tri + 3 * (size_t)nt);
    h->bvh.scene = make_scene(h->verts, h->tris);
    h->bvh.platform.minLeaf = minLeaf;
    h->bvh.platform.maxLeaf = maxLeaf;
    build_bvh(h->bvh, builder == 1 ? BUILDER_SPLIT : BUILDER_SAH, splitAlpha);
    return h;
}
void orc_bvh_free(void* p) { delete (BvhHandle*)p; }

// out: [sah, numInner, numLeaf, numTris, maxDepth, duplicates, buildSeconds]
void orc_bvh_stats(void* p, double* out)
{
    BvhHandle* h = (BvhHandle*)p;
    TreeStats s = tree_stats(h->bvh);
    out[0] = s.sah; out[1] = s.numInner; out[2] = s.numLeaf; out[3] = s.numTris; out[4] = s.maxDepth;
    out[5] = h->bvh.numDuplicates; out[6] = h->bvh.buildSeconds;
}
void orc_bvh_trace(void* p, const float* rays, int n, int needClosest, int32_t* results, uint32_t* counters, int nthreads)
{
    BvhHandle* h = (BvhHandle*)p;
    trace_tree(h->bvh, (const Ray*)rays, (RayResult*)results, n, needClosest != 0, counters, nthreads);
}
void orc_bvh_compact_sizes(void* p, int64_t* sizes)
{
    BvhHandle* h = (BvhHandle*)p;
    if (!h->haveCompact) { create_compact(h->bvh, h->compact, 1); h->haveCompact = true; }
    sizes[0] = (int64_t)h->compact.nodes.size() * 4;
    sizes[1] = (int64_t)h->compact.woop.size() * 4;
    sizes[2] = (int64_t)h->compact.triIndex.size() * 4;
}
void orc_bvh_compact_copy(void* p, void* nodes, void* woop, void* triIndex)
{
    BvhHandle* h = (BvhHandle*)p;
    std::memcpy(nodes, h->compact.nodes.data(), h->compact.nodes.size() * 4);
    std::memcpy(woop, h->compact.woop.data(), h->compact.woop.size() * 4);
    std::memcpy(triIndex, h->compact.triIndex.data(), h->compact.triIndex.size() * 4);
}

// CudaBVH(bvh, layout) for the basic layouts; call with null buffers to get the sizes
void orc_bvh_basic(void* p, int layout, int64_t* sizes, void* nodes, void* woop, void* triIndex)
{
    BvhHandle* h = (BvhHandle*)p;
    CompactBVH b;
    create_basic(h->bvh, layout, b);
    sizes[0] = (int64_t)b.nodes.size() * 4; sizes[1] = (int64_t)b.woop.size() * 4; sizes[2] = (int64_t)b.triIndex.size() * 4;
    if (nodes) std::memcpy(nodes, b.nodes.data(), b.nodes.size() * 4);
    if (woop) std::memcpy(woop, b.woop.data(), b.woop.size() * 4);
    if (triIndex) std::memcpy(triIndex, b.triIndex.data(), b.triIndex.size() * 4);
}

void orc_compact_trace(const int32_t* nodes, const int32_t* woop, const int32_t* triIndex,
                       const float* rays, int n, int needClosest, int32_t* results, uint32_t* counters, int nthreads)
{
    trace_compact(nodes, woop, triIndex, (const Ray*)rays, (RayResult*)results, n, needClosest != 0, counters, nthreads);
}
void orc_wide4_trace(const uint32_t* wnodes, const int32_t* woop, const int32_t* triIndex,
                     const float* rays, int n, int needClosest, int32_t* results, uint32_t* counters, int nthreads)
{
    trace_wide4(wnodes, woop, triIndex, (const Ray*)rays, (RayResult*)results, n, needClosest != 0, counters, nthreads);
}
int orc_wide4_check(const uint32_t* wnodes, int64_t numWide, const int32_t* nodes, int64_t nodeBytes, int layout, double* out4)
{
    return check_wide4(wnodes, (size_t)numWide, nodes, (size_t)nodeBytes, layout, out4);
}
void orc_brute_trace(const float* vtx, int nv, const int32_t* tri, int nt, const float* rays, int n, int needClosest, int32_t* results, int nthreads)
{
    Scene s; s.verts = (const V3*)vtx; s.numVerts = nv; s.tris = tri; s.numTris = nt;
    trace_brute(s, (const Ray*)rays, (RayResult*)results, n, needClosest != 0, nthreads);
}
// out: [sah, numInner, numLeaf, numTris, maxDepth]
void orc_compact_sah(const int32_t* nodes, const int32_t* woop, double* out)
{
    Platform p;
    int a, b, c, d;
    out[0] = compact_sah(nodes, woop, p, &a, &b, &c, &d);
    out[1] = a; out[2] = b; out[3] = c; out[4] = d;
}
void orc_woopify(const float* v9, float* out12, int gpuForm)
{
    V3 v0(v9[0], v9[1], v9[2]), v1(v9[3], v9[4], v9[5]), v2(v9[6], v9[7], v9[8]);
    if (gpuForm) calc_woop_gpu(v0, v1, v2, out12); else woopify_tri(v0, v1, v2, out12);
}
// single-ray primitives for unit tests; out3 = (t, u, v)
void orc_ray_triangle(const float* v9, const float* ray8, float* out3)
{
    out3[1] = out3[2] = 0.0f;
    out3[0] = ray_triangle(V3(v9[0], v9[1], v9[2]), V3(v9[3], v9[4], v9[5]), V3(v9[6], v9[7], v9[8]), *(const Ray*)ray8, out3 + 1, out3 + 2);
}
void orc_ray_triangle_woop(const float* woop12, const float* ray8, float* out3)
{
    out3[1] = out3[2] = 0.0f;
    out3[0] = ray_triangle_woop(woop12, woop12 + 4, woop12 + 8, *(const Ray*)ray8, out3 + 1, out3 + 2);
}
void orc_ray_box(const float* box6, const float* ray8, float* out2)
{
    Span s = ray_box(AABB(V3(box6[0], box6[1], box6[2]), V3(box6[3], box6[4], box6[5])), *(const Ray*)ray8);
    out2[0] = s.tmin; out2[1] = s.tmax;
}

// ---------------- LBVH / HLBVH ---------------------------------------------------------------
void orc_morton(const float* vtx, int nv, const int32_t* tri, int nt, const float* lo, const float* hi, uint32_t* codes)
{
    Scene s; s.verts = (const V3*)vtx; s.numVerts = nv; s.tris = tri; s.numTris = nt;
    morton_codes(s, V3(lo[0], lo[1], lo[2]), V3(hi[0], hi[1], hi[2]), codes);
}
void orc_sort_pairs(uint32_t* keys, int32_t* idx, int n) { sort_pairs_stable(keys, idx, n); }

void* orc_lbvh_build(const float* vtx, int nv, const int32_t* tri, int nt, const float* lo, const float* hi,
                     int hlbvh, int hlbvhBits, int leafSize, float epsilon)
{
    LbvhHandle* h = new LbvhHandle;
    h->verts.resize(nv);
    std::memcpy((void*)h->verts.data(), vtx, sizeof(float) * 3 * (size_t)nv);
    h->tris.assign(tri, tri + 3 * (size_t)nt);
    HLBVHParams p; p.hlbvh = hlbvh != 0; p.hlbvhBits = hlbvhBits; p.leafSize = leafSize; p.epsilon = epsilon;
    build_lbvh(make_scene(h->verts, h->tris), V3(lo[0], lo[1], lo[2]), V3(hi[0], hi[1], hi[2]), p, h->res);
    return h;
}
void orc_lbvh_free(void* p) { delete (LbvhHandle*)p; }
// info: [nodeBytes, woopBytes, idxBytes, numNodes, numLeaves, numClusters, numLevels]
void orc_lbvh_info(void* p, int64_t* info, double* seconds)
{
    LbvhHandle* h = (LbvhHandle*)p;
    info[0] = (int64_t)h->res.bvh.nodes.size() * 4; info[1] = (int64_t)h->res.bvh.woop.size() * 4; info[2] = (int64_t)h->res.bvh.triIndex.size() * 4;
    info[3] = h->res.numNodes; info[4] = h->res.numLeaves; info[5] = h->res.numClusters; info[6] = (int64_t)h->res.levelNodes.size();
    if (seconds) *seconds = h->res.buildSeconds;
}
void orc_lbvh_copy(void* p, void* nodes, void* woop, void* triIndex, uint32_t* sortedKeys, int32_t* sortedIdx)
{
    LbvhHandle* h = (LbvhHandle*)p;
    if (nodes) std::memcpy(nodes, h->res.bvh.nodes.data(), h->res.bvh.nodes.size() * 4);
    if (woop) std::memcpy(woop, h->res.bvh.woop.data(), h->res.bvh.woop.size() * 4);
    if (triIndex) std::memcpy(triIndex, h->res.bvh.triIndex.data(), h->res.bvh.triIndex.size() * 4);
    if (sortedKeys) std::memcpy(sortedKeys, h->res.sortedKeys.data(), h->res.sortedKeys.size() * 4);
    if (sortedIdx) std::memcpy(sortedIdx, h->res.sortedIdx.data(), h->res.sortedIdx.size() * 4);
}

// canonical form of any Compact tree: call once with null outputs to get sizes
// sizes: [numInner, numLeaves, numTris]
void orc_canonical(const int32_t* nodes, const int32_t* woop, const int32_t* triIndex, int64_t* sizes,
                   int32_t* inner, float* boxes, int32_t* leafSizes, int32_t* tris, float* woopOut)
{
    Canonical c;
    canonicalize(nodes, woop, triIndex, c);
    sizes[0] = (int64_t)c.inner.size() / 3; sizes[1] = (int64_t)c.leafSizes.size(); sizes[2] = (int64_t)c.tris.size();
    if (inner) std::memcpy(inner, c.inner.data(), c.inner.size() * 4);
    if (boxes) std::memcpy(boxes, c.boxes.data(), c.boxes.size() * 4);
    if (leafSizes) std::memcpy(leafSizes, c.leafSizes.data(), c.leafSizes.size() * 4);
    if (tris) std::memcpy(tris, c.tris.data(), c.tris.size() * 4);
    if (woopOut) std::memcpy(woopOut, c.woop.data(), c.woop.size() * 4);
}

// ---------------- ray generation --------------------------------------------------------------
void orc_pixel_table(int w, int h, int32_t* indexToPixel, int32_t* pixelToIndex) { pixel_table(w, h, indexToPixel, pixelToIndex); }
void orc_raygen_primary(float* rays, int32_t* idToSlot, int32_t* slotToID, const float* origin, const float* n2w16,
                        int w, int h, float maxDist, uint32_t seed)
{
    M4 m; std::memcpy(m.m, n2w16, 64);
    raygen_primary((Ray*)rays, idToSlot, slotToID, V3(origin[0], origin[1], origin[2]), m, w, h, maxDist, seed);
}
void orc_raygen_ao(float* outRays, int32_t* outIDToSlot, int32_t* outSlotToID, const float* inRays, const int32_t* inResults,
                   const float* normals, int firstInputSlot, int numInputRays, int numSamples, float maxDist, uint32_t seed)
{
    raygen_ao((Ray*)outRays, outIDToSlot, outSlotToID, (const Ray*)inRays, (const RayResult*)inResults, (const V3*)normals,
              firstInputSlot, numInputRays, numSamples, maxDist, seed);
}
void orc_raygen_shadow(float* outRays, int32_t* outIDToSlot, int32_t* outSlotToID, const float* inRays, const int32_t* inResults,
                       int firstInputSlot, int numInputRays, int numSamples, const float* lightPos, float lightRadius, uint32_t seed)
{
    raygen_shadow((Ray*)outRays, outIDToSlot, outSlotToID, (const Ray*)inRays, (const RayResult*)inResults,
                  firstInputSlot, numInputRays, numSamples, V3(lightPos[0], lightPos[1], lightPos[2]), lightRadius, seed);
}
int orc_count_hits(const int32_t* results, int n) { return count_hits((const RayResult*)results, n); }
void orc_tri_normals(const float* vtx, int nv, const int32_t* tri, int nt, float* out)
{
    Scene s; s.verts = (const V3*)vtx; s.numVerts = nv; s.tris = tri; s.numTris = nt;
    tri_normals(s, (V3*)out);
}
void orc_ray_morton_keys(const float* rays, int n, uint32_t* keys6, float* aabb6)
{
    ray_morton_keys((const Ray*)rays, n, keys6, aabb6);
}
void orc_ray_morton_order(const float* rays, int n, int truncated, int32_t* order, uint64_t* key64)
{
    ray_morton_order((const Ray*)rays, n, truncated, order, key64);
}
void orc_invert4(const float* in16, float* out16)
{
    M4 a; std::memcpy(a.m, in16, 64);
    M4 r = invert(a);
    std::memcpy(out16, r.m, 64);
}

} // extern "C"
