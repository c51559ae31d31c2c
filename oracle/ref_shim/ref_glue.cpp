// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the reference's OWN CPU implementation of the path, compiled
// unmodified from /root/reference (see oracle/Makefile target _ref/libref.so): SAHBVHBuilder / SplitBVHBuilder, the SAH
// metric (BVHNode::computeSubtreeProbabilities via BVH::BVH), BVH::trace, CudaBVH::createCompact + woopifyTri, the flat
// CudaBVH::trace and the Intersect:: primitives.  Only platform glue lives here: the four FW:: runtime functions that
// base/Defs.cpp implements with Win32 calls, and RayBuffer::resize / setRay (ray/RayBuffer.cpp drags in the runtime nvcc
// compiler).  Used by tests/test_reference_pin.py to pin the restated oracle against the reference itself.
#include "bvh/BVH.hpp"
#include "cuda/CudaBVH.hpp"
#include "ray/PixelTable.hpp"
#include "Environment.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <thread>
#include <vector>

namespace FW
{
void* malloc(size_t size) { return ::malloc(size); }
void free(void* ptr) { ::free(ptr); }
void* realloc(void* ptr, size_t size) { return ::realloc(ptr, size); }
void printf(const char* fmt, ...) { (void)fmt; }            // build progress prints are not wanted in tests
void fail(const char* fmt, ...)
{
    char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    fprintf(stderr, "reference fail(): %s\n", buf);
    abort();
}

static bool s_hasError = false;       // (FW:: namespace scope below)
void setError(const char* fmt, ...)
{
    char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    fprintf(stderr, "reference setError(): %s\n", buf);
    s_hasError = true;
}
bool hasError(void) { return s_hasError; }

// ray/RayBuffer.cpp:38-62
void RayBuffer::resize(S32 n)
{
    if (n < m_size) { m_size = n; return; }
    m_size = n;
    m_rays.resize(n * sizeof(Ray)); m_results.resize(n * sizeof(RayResult));
    m_IDToSlot.resize(n * sizeof(S32)); m_slotToID.resize(n * sizeof(S32));
}
void RayBuffer::setRay(S32 slot, const Ray& ray, S32 id)
{
    ((Ray*)m_rays.getMutablePtr())[slot] = ray;
    ((S32*)m_IDToSlot.getMutablePtr())[id] = slot;
    ((S32*)m_slotToID.getMutablePtr())[slot] = id;
}
}

using namespace FW;

struct RefHandle
{
    Scene* scene; Platform platform; BVH::Stats stats; BVH* bvh; CudaBVH* compact;
    std::vector<CudaBVH*> perThread;     // CudaBVH::trace keeps per-call state in members: one object per worker thread
    CudaBVH* byLayout[BVHLayout_Max];    // CudaBVH(bvh, layout) for the basic layouts (createNodeBasic / createTriWoopBasic / createTriIndexBasic)
};

static CudaBVH* layout_bvh(RefHandle* h, int layout)
{
    if (layout < 0 || layout >= BVHLayout_Max) return NULL;
    if (!h->byLayout[layout]) h->byLayout[layout] = new CudaBVH(*h->bvh, (BVHLayout)layout);
    return h->byLayout[layout];
}

static void fill_rays(RayBuffer& rb, const float* rays, int n, bool closest)
{
    rb.resize(n);
    rb.setNeedClosestHit(closest);
    for (int i = 0; i < n; i++) {
        Ray r;
        r.origin = Vec3f(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]); r.tmin = rays[8 * i + 3];
        r.direction = Vec3f(rays[8 * i + 4], rays[8 * i + 5], rays[8 * i + 6]); r.tmax = rays[8 * i + 7];
        rb.setRay(i, r);
        RayResult rr; rr.id = -1; rr.t = 0.0f; rr.padA = 0; rr.padB = 0;
        rb.setResult(i, rr);
    }
}

extern "C" {

void* ref_build(const float* verts, int nv, const int* tris, int nt, int splitBVH, int minLeaf, int maxLeaf, float alpha)
{
    RefHandle* h = new RefHandle;
    h->scene = new Scene((const Vec3f*)verts, nv, (const Vec3i*)tris, nt);
    h->platform = Platform("GPU");
    h->platform.setLeafPreferences(minLeaf, maxLeaf);                // Renderer.cpp:88-89
    Environment::GetSingleton()->builder = splitBVH ? "SplitBVH" : "SAHBVH";
    BVH::BuildParams params;
    params.stats = &h->stats;
    params.enablePrints = false;
    params.splitAlpha = alpha;
    h->bvh = new BVH(h->scene, h->platform, params);
    h->compact = NULL;
    for (int i = 0; i < BVHLayout_Max; i++) h->byLayout[i] = NULL;
    return h;
}
void ref_free(void* p) { RefHandle* h = (RefHandle*)p; for (size_t i = 0; i < h->perThread.size(); i++) delete h->perThread[i]; for (int i = 0; i < BVHLayout_Max; i++) delete h->byLayout[i]; delete h->compact; delete h->bvh; delete h->scene; delete h; }

// out: [SAHCost, numInner, numLeaf, numTris, maxDepth, numTriIndices]
void ref_stats(void* p, double* out)
{
    RefHandle* h = (RefHandle*)p;
    out[0] = h->stats.SAHCost; out[1] = h->stats.numInnerNodes; out[2] = h->stats.numLeafNodes; out[3] = h->stats.numTris;
    out[4] = h->stats.maxDepth; out[5] = h->bvh->getTriIndices().getSize();
}
void ref_tri_indices(void* p, int* out)
{
    RefHandle* h = (RefHandle*)p;
    const Array<S32>& t = h->bvh->getTriIndices();
    for (int i = 0; i < t.getSize(); i++) out[i] = t[i];
}
void ref_trace(void* p, const float* rays, int n, int closest, int* results)
{
    RefHandle* h = (RefHandle*)p;
    RayBuffer rb;
    fill_rays(rb, rays, n, closest != 0);
    h->bvh->trace(rb, NULL);                                          // BVH.cpp:90-110
    for (int i = 0; i < n; i++) memcpy(results + 4 * i, &rb.getResultForSlot(i), 16);
}
// CudaBVH(bvh, BVHLayout_Compact): createCompact + (this fork) random node shuffle
void ref_compact_sizes(void* p, long long* sizes)
{
    RefHandle* h = (RefHandle*)p;
    if (!h->compact) h->compact = new CudaBVH(*h->bvh, BVHLayout_Compact);
    sizes[0] = h->compact->getNodeBuffer().getSize(); sizes[1] = h->compact->getTriWoopBuffer().getSize(); sizes[2] = h->compact->getTriIndexBuffer().getSize();
}
void ref_compact_copy(void* p, void* nodes, void* woop, void* idx)
{
    RefHandle* h = (RefHandle*)p;
    memcpy(nodes, h->compact->getNodeBuffer().getPtr(), (size_t)h->compact->getNodeBuffer().getSize());
    memcpy(woop, h->compact->getTriWoopBuffer().getPtr(), (size_t)h->compact->getTriWoopBuffer().getSize());
    memcpy(idx, h->compact->getTriIndexBuffer().getPtr(), (size_t)h->compact->getTriIndexBuffer().getSize());
}
void ref_compact_trace(void* p, const float* rays, int n, int closest, int* results)
{
    RefHandle* h = (RefHandle*)p;
    if (!h->compact) h->compact = new CudaBVH(*h->bvh, BVHLayout_Compact);
    RayBuffer rb;
    fill_rays(rb, rays, n, closest != 0);
    Buffer visibility;
    h->compact->trace(rb, visibility, false, NULL);                   // CudaBVH.cpp:213-302
    for (int i = 0; i < n; i++) memcpy(results + 4 * i, &rb.getResultForSlot(i), 16);
}
// CudaBVH(bvh, layout) for any BVHLayout (CudaBVH.cpp:60-101, 453-575): buffers and, for AOS_AOS / Compact, the CPU trace
int ref_layout_sizes(void* p, int layout, long long* sizes)
{
    CudaBVH* b = layout_bvh((RefHandle*)p, layout);
    if (!b) return 1;
    sizes[0] = b->getNodeBuffer().getSize(); sizes[1] = b->getTriWoopBuffer().getSize(); sizes[2] = b->getTriIndexBuffer().getSize();
    return 0;
}
int ref_layout_copy(void* p, int layout, void* nodes, void* woop, void* idx)
{
    CudaBVH* b = layout_bvh((RefHandle*)p, layout);
    if (!b) return 1;
    memcpy(nodes, b->getNodeBuffer().getPtr(), (size_t)b->getNodeBuffer().getSize());
    memcpy(woop, b->getTriWoopBuffer().getPtr(), (size_t)b->getTriWoopBuffer().getSize());
    memcpy(idx, b->getTriIndexBuffer().getPtr(), (size_t)b->getTriIndexBuffer().getSize());
    return 0;
}
int ref_layout_trace(void* p, int layout, const float* rays, int n, int closest, int* results)
{
    CudaBVH* b = layout_bvh((RefHandle*)p, layout);
    if (!b || (layout != BVHLayout_AOS_AOS && layout != BVHLayout_Compact)) return 1;     // the layouts CudaBVH::trace switches on
    RayBuffer rb;
    fill_rays(rb, rays, n, closest != 0);
    Buffer visibility;
    b->trace(rb, visibility, false, NULL);
    for (int i = 0; i < n; i++) memcpy(results + 4 * i, &rb.getResultForSlot(i), 16);
    return 0;
}
// the same call fanned out over host threads (the reference is single-threaded; bench.py --impl reference uses every core
// the box gives it).  Each worker owns a CudaBVH built from the same BVH and a contiguous slice of the rays.
void ref_compact_trace_mt(void* p, const float* rays, int n, int closest, int* results, int nthreads)
{
    RefHandle* h = (RefHandle*)p;
    if (nthreads < 1) nthreads = 1;
    while ((int)h->perThread.size() < nthreads) h->perThread.push_back(new CudaBVH(*h->bvh, BVHLayout_Compact));
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) {
        int lo = (int)((long long)n * t / nthreads), hi = (int)((long long)n * (t + 1) / nthreads);
        pool.push_back(std::thread([=]() {
            if (hi <= lo) return;
            RayBuffer rb;
            fill_rays(rb, rays + 8 * (size_t)lo, hi - lo, closest != 0);
            Buffer visibility;
            h->perThread[t]->trace(rb, visibility, false, NULL);
            for (int i = lo; i < hi; i++) memcpy(results + 4 * (size_t)i, &rb.getResultForSlot(i - lo), 16);
        }));
    }
    for (size_t t = 0; t < pool.size(); t++) pool[t].join();
}
void ref_ray_box(const float* box6, const float* ray8, float* out2)
{
    Ray r; r.origin = Vec3f(ray8[0], ray8[1], ray8[2]); r.tmin = ray8[3]; r.direction = Vec3f(ray8[4], ray8[5], ray8[6]); r.tmax = ray8[7];
    Vec2f s = Intersect::RayBox(AABB(Vec3f(box6[0], box6[1], box6[2]), Vec3f(box6[3], box6[4], box6[5])), r);
    out2[0] = s.x; out2[1] = s.y;
}
void ref_ray_triangle(const float* v9, const float* ray8, float* out3)
{
    Ray r; r.origin = Vec3f(ray8[0], ray8[1], ray8[2]); r.tmin = ray8[3]; r.direction = Vec3f(ray8[4], ray8[5], ray8[6]); r.tmax = ray8[7];
    Vec3f b = Intersect::RayTriangle(Vec3f(v9[0], v9[1], v9[2]), Vec3f(v9[3], v9[4], v9[5]), Vec3f(v9[6], v9[7], v9[8]), r);
    out3[0] = b.z; out3[1] = b.x; out3[2] = b.y;                      // (t, u, v)
}
void ref_ray_triangle_woop(const float* w12, const float* ray8, float* out3)
{
    Ray r; r.origin = Vec3f(ray8[0], ray8[1], ray8[2]); r.tmin = ray8[3]; r.direction = Vec3f(ray8[4], ray8[5], ray8[6]); r.tmax = ray8[7];
    Vec3f b = Intersect::RayTriangleWoop(Vec4f(w12[0], w12[1], w12[2], w12[3]), Vec4f(w12[4], w12[5], w12[6], w12[7]), Vec4f(w12[8], w12[9], w12[10], w12[11]), r);
    out3[0] = b.z; out3[1] = b.x; out3[2] = b.y;
}
// ray/PixelTable.cpp:57-141 (compiled unmodified): the index <-> pixel tables primary ray generation walks
void ref_pixel_table(int w, int h, int* indexToPixel, int* pixelToIndex)
{
    PixelTable pt;
    pt.setSize(Vec2i(w, h));
    memcpy(indexToPixel, pt.getIndexToPixel().getPtr(), (size_t)w * h * sizeof(S32));
    memcpy(pixelToIndex, pt.getPixelToIndex().getPtr(), (size_t)w * h * sizeof(S32));
}
void ref_invert4(const float* in16, float* out16)
{
    Mat4f m;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m(i, j) = in16[i * 4 + j];
    Mat4f r = invert(m);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out16[i * 4 + j] = r(i, j);
}

// FW::hashBuffer / the cache-name formula of Renderer::getCudaBVH (Renderer.cpp:173-178) evaluated by the reference's own Hash.cpp
unsigned ref_hash_buffer(const void* p, int size) { return hashBuffer(p, size); }
unsigned ref_cache_name_hash(unsigned sceneHash, int minLeaf, int maxLeaf, float splitAlpha, int layout, const char* ds)
{
    Platform platform("GPU");
    platform.setLeafPreferences(minLeaf, maxLeaf);
    BVH::BuildParams params;
    params.splitAlpha = splitAlpha;
    return hashBits(sceneHash, platform.computeHash(), params.computeHash(), (U32)layout, hash<String>(String(ds)));
}
unsigned ref_scene_hash(unsigned triVtxIndex, unsigned triNormal, unsigned triMaterialColor, unsigned triShadedColor, unsigned vtxPos)
{
    return hashBits(triVtxIndex, triNormal, triMaterialColor, triShadedColor, vtxPos);      // Scene.cpp:171-179
}

// CudaBVH::serialize (CudaBVH.cpp:116-125) through the reference's own OutputStream operators (io/Stream.cpp) into memory: the byte
// stream Renderer::getCudaBVH writes to bvhcache/<hash>_<builder>.dat (Renderer.cpp:293-299).  Returns the stream length; copies
// at most cap bytes.  (Buffer's "S64 size, then the bytes" is the shim's statement of Buffer.cpp:349-381: the real Buffer.cpp needs
// the CUDA driver API.)
long long ref_serialize(void* p, int layout, void* out, long long cap)
{
    RefHandle* h = (RefHandle*)p;
    if (layout < 0 || layout >= BVHLayout_Max) return -1;
    if (!h->byLayout[layout]) h->byLayout[layout] = new CudaBVH(*h->bvh, (BVHLayout)layout);
    MemoryOutputStream ms;
    h->byLayout[layout]->serialize(ms);
    const Array<U8>& d = ms.getData();
    if (out && cap > 0) memcpy(out, d.getPtr(), (size_t)(d.getSize() < cap ? d.getSize() : cap));
    return d.getSize();
}
// CudaBVH(InputStream&) (CudaBVH.cpp:105-108): read a stream back with the reference's reader.  sizes = [layout, nodeBytes, woopBytes, idxBytes]
int ref_deserialize(const void* bytes, long long n, long long* sizes, void* nodes, void* woop, void* idx)
{
    MemoryInputStream in(bytes, (int)n);
    CudaBVH bvh(in);
    if (hasError()) { s_hasError = false; return 1; }
    sizes[0] = (long long)bvh.getLayout(); sizes[1] = bvh.getNodeBuffer().getSize(); sizes[2] = bvh.getTriWoopBuffer().getSize(); sizes[3] = bvh.getTriIndexBuffer().getSize();
    if (nodes) memcpy(nodes, bvh.getNodeBuffer().getPtr(), (size_t)sizes[1]);
    if (woop) memcpy(woop, bvh.getTriWoopBuffer().getPtr(), (size_t)sizes[2]);
    if (idx) memcpy(idx, bvh.getTriIndexBuffer().getPtr(), (size_t)sizes[3]);
    return 0;
}

} // extern "C"
