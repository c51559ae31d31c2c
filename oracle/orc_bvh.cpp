// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
#include "orc_bvh.hpp"
#include <chrono>
#include <climits>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// =========================================================================================
// Intersections — src/rt/Util.cpp
// =========================================================================================

// Util.cpp:34-46.  IEEE division by the direction (no reciprocal), min/max reductions.
Span ray_box(const AABB& box, const Ray& ray)
{
    V3 t0 = (box.mn - ray.o) / ray.d;
    V3 t1 = (box.mx - ray.o) / ray.d;
    Span s;
    s.tmin = hmax(vmin(t0, t1));
    s.tmax = hmin(vmax(t0, t1));
    return s;
}

// Util.cpp:50-94.  Moller-Trumbore, EPSILON = 0, two-sided, strict (tmin, tmax) acceptance.
float ray_triangle(V3 v0, V3 v1, V3 v2, const Ray& ray, float* uo, float* vo)
{
    V3 e1 = v1 - v0;
    V3 e2 = v2 - v0;
    V3 pvec = cross(ray.d, e2);
    float det = dot(e1, pvec);
    V3 tvec = ray.o - v0;
    float u = dot(tvec, pvec);
    V3 qvec = cross(tvec, e1);
    float v = dot(ray.d, qvec);

    if (det > 0.0f) {
        if (u < 0.0f || u > det) return F32_MAX;
        if (v < 0.0f || u + v > det) return F32_MAX;
    } else if (det < -0.0f) {
        if (u > 0.0f || u < det) return F32_MAX;
        if (v > 0.0f || u + v < det) return F32_MAX;
    } else
        return F32_MAX;

    float inv = 1.0f / det;
    float t = dot(e2, qvec) * inv;
    u *= inv;
    v *= inv;
    if (t > ray.tmin && t < ray.tmax) {
        if (uo) *uo = u;
        if (vo) *vo = v;
        return t;
    }
    return F32_MAX;
}

// Util.cpp:99-127.  4-component dots with w = 1 for the origin and 0 for the direction.
float ray_triangle_woop(const float* z, const float* up, const float* vp, const Ray& ray, float* uo, float* vo)
{
    const float ox = ray.o.x, oy = ray.o.y, oz = ray.o.z;
    const float dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    auto dot4 = [](const float* a, float b0, float b1, float b2, float b3) {
        float r = 0.0f; r += a[0] * b0; r += a[1] * b1; r += a[2] * b2; r += a[3] * b3; return r;
    };
    float Oz = z[3] - ox * z[0] - oy * z[1] - oz * z[2];
    // dot(dir, zpleq) with the vector on the left: r += dir[i]*zpleq[i]
    float dd = 0.0f; dd += dx * z[0]; dd += dy * z[1]; dd += dz * z[2]; dd += 0.0f * z[3];
    float ooDz = 1.0f / dd;
    float t = Oz * ooDz;
    if (t > ray.tmin && t < ray.tmax) {
        float Ou = dot4(up, ox, oy, oz, 1.0f);
        float Du = dot4(up, dx, dy, dz, 0.0f);
        float u = Ou + t * Du;
        if (u >= 0.0f) {
            float Ov = dot4(vp, ox, oy, oz, 1.0f);
            float Dv = dot4(vp, dx, dy, dz, 0.0f);
            float v = Ov + t * Dv;
            if (v >= 0.0f && (u + v) <= 1.0f) {
                if (uo) *uo = u;
                if (vo) *vo = v;
                return t;
            }
        }
    }
    return F32_MAX;
}

// =========================================================================================
// Sort — src/framework/base/Sort.cpp:62-160.  The SAH comparator is a total order except
// for duplicated references with equal centroids, so the exact algorithm (median-of-3
// quicksort, insertion sort below 16, bounded explicit stack) is restated to keep the
// same tie order.
// =========================================================================================
namespace {

template <class Less, class Swap>
struct IndexSorter {
    Less less;
    Swap swp;

    void insertion(int start, int size) const {
        for (int i = 1; i < size; i++) {
            int j = start + i - 1;
            while (j >= start && less(j + 1, j)) { swp(j, j + 1); j--; }
        }
    }
    int median3(int low, int high) const {
        int l = low, c = (low + high) >> 1, h = high - 2;
        if (less(h, l)) std::swap(l, h);
        if (less(c, l)) c = l;
        return less(h, c) ? h : c;
    }
    int partition(int low, int high) const {
        swp(median3(low, high), high - 1);
        int i = low - 1, j = high - 1;
        for (;;) {
            do i++; while (less(i, high - 1));
            do j--; while (less(high - 1, j));
            if (i >= j) break;
            swp(i, j);
        }
        swp(i, high - 1);
        return i;
    }
    void sort(int low, int high) const {
        if (high - low < 2) return;
        const int kStack = 32, kMin = 16;
        int stack[kStack];
        int sp = 0;
        stack[sp++] = high;
        while (sp) {
            high = stack[--sp];
            if (high - low < kMin || sp + 2 > kStack) {
                insertion(low, high - low);
                low = high + 1;
                continue;
            }
            int i = partition(low, high);
            if (high - i > 2) stack[sp++] = high;
            if (i - low > 1) stack[sp++] = i;
            else low = i + 1;
        }
    }
};
template <class Less, class Swap>
static inline void index_sort(int low, int high, Less less, Swap swp) { IndexSorter<Less, Swap>{less, swp}.sort(low, high); }

// =========================================================================================
// Builders — SAHBVHBuilder.cpp / SplitBVHBuilder.cpp
// =========================================================================================
struct Ref { int tri = -1; AABB b; };
struct Spec { int numRef = 0; AABB b; };
struct ObjSplit { float sah = F32_MAX; int dim = 0; int numLeft = 0; AABB lb, rb; };
struct SpatSplit { float sah = F32_MAX; int dim = 0; float pos = 0.0f; };
struct Bin { AABB b; int enter = 0, exit = 0; };

static inline float min3f(float a, float b, float c) { return fw_min(fw_min(a, b), c); }

// float -> int as the x86 cvttss2si the reference compiles to: out-of-range / NaN -> INT_MIN.
static inline int trunc_i(float f) {
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT_MIN;
    return (int)f;
}
static inline int clampi(int v, int lo, int hi) { return std::min(std::max(v, lo), hi); }

class Builder {
public:
    static constexpr int kMaxDepth = 64;          // SAHBVHBuilder.hpp:51
    static constexpr int kMaxSpatialDepth = 48;   // SplitBVHBuilder.hpp:54
    static constexpr int kBins = 128;             // SplitBVHBuilder.hpp:55

    Builder(BVH& bvh, BuilderKind kind, float alpha) : m_bvh(bvh), m_kind(kind), m_alpha(alpha), m_p(bvh.platform) {}

    int run() {
        const Scene& sc = m_bvh.scene;
        Spec root;
        root.numRef = sc.numTris;
        m_refs.resize(sc.numTris);
        for (int i = 0; i < sc.numTris; i++) {
            m_refs[i].tri = i;
            for (int j = 0; j < 3; j++) m_refs[i].b.grow(sc.v(i, j));
            root.b.grow(m_refs[i].b);
        }
        if (m_kind == BUILDER_SPLIT) {
            m_minOverlap = root.b.area() * m_alpha;                              // SplitBVHBuilder.cpp:70
            m_right.assign(std::max(root.numRef, (int)kBins) - 1 + 1, AABB());
        } else
            m_right.assign(std::max(root.numRef, 1), AABB());
        m_dups = 0;
        int r = buildNode(root, 0);
        m_bvh.numDuplicates = m_dups;
        return r;
    }

private:
    int newLeaf(const Spec& spec) {                                              // SAHBVHBuilder.cpp:190-196
        auto& tris = m_bvh.triIndices;
        for (int i = 0; i < spec.numRef; i++) { tris.push_back(m_refs.back().tri); m_refs.pop_back(); }
        Node n; n.leaf = true; n.bounds = spec.b; n.lo = (int)tris.size() - spec.numRef; n.hi = (int)tris.size();
        m_bvh.nodes.push_back(n);
        return (int)m_bvh.nodes.size() - 1;
    }

    void sortRefs(int start, int end, int dim) {                                 // SAHBVHBuilder.cpp:106-124
        Ref* refs = m_refs.data();
        index_sort(start, end,
            [refs, dim](int a, int b) {
                const Ref& ra = refs[a]; const Ref& rb = refs[b];
                float ca = ra.b.mn[dim] + ra.b.mx[dim];
                float cb = rb.b.mn[dim] + rb.b.mx[dim];
                return (ca < cb || (ca == cb && ra.tri < rb.tri));
            },
            [refs](int a, int b) { std::swap(refs[a], refs[b]); });
    }

    ObjSplit findObjectSplit(const Spec& spec, float nodeSAH) {                  // SAHBVHBuilder.cpp:200-241
        ObjSplit split;
        int base = (int)m_refs.size() - spec.numRef;
        float bestTie = F32_MAX;
        for (int dim = 0; dim < 3; dim++) {
            sortRefs(base, (int)m_refs.size(), dim);
            AABB rb;
            for (int i = spec.numRef - 1; i > 0; i--) { rb.grow(m_refs[base + i].b); m_right[i - 1] = rb; }
            AABB lb;
            for (int i = 1; i < spec.numRef; i++) {
                lb.grow(m_refs[base + i - 1].b);
                float sah = nodeSAH + lb.area() * m_p.triangleCost(i) + m_right[i - 1].area() * m_p.triangleCost(spec.numRef - i);
                float fi = (float)i, fr = (float)(spec.numRef - i);
                float tie = fi * fi + fr * fr;
                if (sah < split.sah || (sah == split.sah && tie < bestTie)) {
                    split.sah = sah; split.dim = dim; split.numLeft = i; split.lb = lb; split.rb = m_right[i - 1];
                    bestTie = tie;
                }
            }
        }
        return split;
    }

    void performObjectSplit(Spec& l, Spec& r, const Spec& spec, const ObjSplit& s) {   // SAHBVHBuilder.cpp:245-254
        sortRefs((int)m_refs.size() - spec.numRef, (int)m_refs.size(), s.dim);
        l.numRef = s.numLeft; l.b = s.lb;
        r.numRef = spec.numRef - s.numLeft; r.b = s.rb;
    }

    void splitReference(Ref& left, Ref& right, const Ref& ref, int dim, float pos) {   // SplitBVHBuilder.cpp:349-393
        left.tri = right.tri = ref.tri;
        left.b = right.b = AABB();
        const Scene& sc = m_bvh.scene;
        V3 v1 = sc.v(ref.tri, 2);
        for (int i = 0; i < 3; i++) {
            V3 v0 = v1;
            v1 = sc.v(ref.tri, i);
            float v0p = v0[dim], v1p = v1[dim];
            if (v0p <= pos) left.b.grow(v0);
            if (v0p >= pos) right.b.grow(v0);
            if ((v0p < pos && v1p > pos) || (v0p > pos && v1p < pos)) {
                V3 t = lerp(v0, v1, clampf((pos - v0p) / (v1p - v0p), 0.0f, 1.0f));
                left.b.grow(t);
                right.b.grow(t);
            }
        }
        left.b.mx[dim] = pos;
        right.b.mn[dim] = pos;
        left.b.intersect(ref.b);
        right.b.intersect(ref.b);
    }

    SpatSplit findSpatialSplit(const Spec& spec, float nodeSAH) {                // SplitBVHBuilder.cpp:170-249
        V3 origin = spec.b.mn;
        V3 binSize = (spec.b.mx - origin) * (1.0f / (float)kBins);
        V3 invBin = V3(1.0f / binSize.x, 1.0f / binSize.y, 1.0f / binSize.z);
        for (int d = 0; d < 3; d++) for (int i = 0; i < kBins; i++) m_bins[d][i] = Bin();

        for (int ri = (int)m_refs.size() - spec.numRef; ri < (int)m_refs.size(); ri++) {
            const Ref ref = m_refs[ri];
            V3 fa = (ref.b.mn - origin) * invBin;
            V3 fb = (ref.b.mx - origin) * invBin;
            int first[3], last[3];
            for (int d = 0; d < 3; d++) {
                first[d] = clampi(trunc_i(fa[d]), 0, kBins - 1);
                last[d] = clampi(trunc_i(fb[d]), first[d], kBins - 1);
            }
            for (int d = 0; d < 3; d++) {
                Ref cur = ref;
                for (int i = first[d]; i < last[d]; i++) {
                    Ref l, r;
                    splitReference(l, r, cur, d, origin[d] + binSize[d] * (float)(i + 1));
                    m_bins[d][i].b.grow(l.b);
                    cur = r;
                }
                m_bins[d][last[d]].b.grow(cur.b);
                m_bins[d][first[d]].enter++;
                m_bins[d][last[d]].exit++;
            }
        }

        SpatSplit split;
        for (int d = 0; d < 3; d++) {
            AABB rb;
            for (int i = kBins - 1; i > 0; i--) { rb.grow(m_bins[d][i].b); m_right[i - 1] = rb; }
            AABB lb;
            int ln = 0, rn = spec.numRef;
            for (int i = 1; i < kBins; i++) {
                lb.grow(m_bins[d][i - 1].b);
                ln += m_bins[d][i - 1].enter;
                rn -= m_bins[d][i - 1].exit;
                float sah = nodeSAH + lb.area() * m_p.triangleCost(ln) + m_right[i - 1].area() * m_p.triangleCost(rn);
                if (sah < split.sah) { split.sah = sah; split.dim = d; split.pos = origin[d] + binSize[d] * (float)i; }
            }
        }
        return split;
    }

    void performSpatialSplit(Spec& left, Spec& right, const Spec& spec, const SpatSplit& split) {   // SplitBVHBuilder.cpp:253-345
        auto& refs = m_refs;
        int leftStart = (int)refs.size() - spec.numRef;
        int leftEnd = leftStart;
        int rightStart = (int)refs.size();
        left.b = right.b = AABB();

        for (int i = leftEnd; i < rightStart; i++) {
            if (refs[i].b.mx[split.dim] <= split.pos) {
                left.b.grow(refs[i].b);
                std::swap(refs[i], refs[leftEnd++]);
            } else if (refs[i].b.mn[split.dim] >= split.pos) {
                right.b.grow(refs[i].b);
                std::swap(refs[i--], refs[--rightStart]);
            }
        }

        while (leftEnd < rightStart) {
            Ref lref, rref;
            splitReference(lref, rref, refs[leftEnd], split.dim, split.pos);

            AABB lub = left.b, rub = right.b, ldb = left.b, rdb = right.b;
            lub.grow(refs[leftEnd].b);
            rub.grow(refs[leftEnd].b);
            ldb.grow(lref.b);
            rdb.grow(rref.b);

            float lac = m_p.triangleCost(leftEnd - leftStart);
            float rac = m_p.triangleCost((int)refs.size() - rightStart);
            float lbc = m_p.triangleCost(leftEnd - leftStart + 1);
            float rbc = m_p.triangleCost((int)refs.size() - rightStart + 1);

            float unsplitL = lub.area() * lbc + right.b.area() * rac;
            float unsplitR = left.b.area() * lac + rub.area() * rbc;
            float dupSAH = ldb.area() * lbc + rdb.area() * rbc;
            float mn = min3f(unsplitL, unsplitR, dupSAH);

            if (mn == unsplitL) { left.b = lub; leftEnd++; }
            else if (mn == unsplitR) { right.b = rub; std::swap(refs[leftEnd], refs[--rightStart]); }
            else { left.b = ldb; right.b = rdb; refs[leftEnd++] = lref; refs.push_back(rref); }
        }
        left.numRef = leftEnd - leftStart;
        right.numRef = (int)refs.size() - rightStart;
    }

    int buildNode(Spec spec, int level) {        // SAHBVHBuilder.cpp:127-186 / SplitBVHBuilder.cpp:90-166
        // Remove degenerates.
        {
            int firstRef = (int)m_refs.size() - spec.numRef;
            for (int i = (int)m_refs.size() - 1; i >= firstRef; i--) {
                V3 size = m_refs[i].b.mx - m_refs[i].b.mn;
                if (hmin(size) < 0.0f || hsum(size) == hmax(size)) {
                    // Array::removeSwap (Array.hpp:473-482)
                    int lastIdx = (int)m_refs.size() - 1;
                    if (i < lastIdx) m_refs[i] = m_refs[lastIdx];
                    m_refs.pop_back();
                }
            }
            spec.numRef = (int)m_refs.size() - firstRef;
        }

        if ((level != 0 && spec.numRef <= m_p.minLeaf) || level >= kMaxDepth)
            return newLeaf(spec);

        float area = spec.b.area();
        float leafSAH = area * m_p.triangleCost(spec.numRef);
        float nodeSAH = area * m_p.nodeCostN(2);

        int splitType = 0, axis = 0;
        ObjSplit object = findObjectSplit(spec, nodeSAH);

        SpatSplit spatial;
        if (m_kind == BUILDER_SPLIT && level < kMaxSpatialDepth) {
            AABB overlap = object.lb;
            overlap.intersect(object.rb);
            if (overlap.area() >= m_minOverlap)
                spatial = findSpatialSplit(spec, nodeSAH);
        }

        float minSAH = (m_kind == BUILDER_SPLIT) ? min3f(leafSAH, object.sah, spatial.sah) : fw_min(leafSAH, object.sah);
        if (level != 0 && minSAH == leafSAH && spec.numRef <= m_p.maxLeaf)
            return newLeaf(spec);

        Spec left, right;
        if (m_kind == BUILDER_SPLIT) {
            if (minSAH == spatial.sah)
                performSpatialSplit(left, right, spec, spatial);
            if (!left.numRef || !right.numRef) { performObjectSplit(left, right, spec, object); axis = object.dim; }
            else { splitType = 1; axis = spatial.dim; }
        } else {
            performObjectSplit(left, right, spec, object);
            axis = object.dim;
        }

        m_dups += left.numRef + right.numRef - spec.numRef;
        int rn = buildNode(right, level + 1);     // right first: refs live on a stack
        int ln = buildNode(left, level + 1);
        Node n; n.leaf = false; n.bounds = spec.b; n.child[0] = ln; n.child[1] = rn; n.axis = axis; n.splitType = splitType;
        m_bvh.nodes.push_back(n);
        return (int)m_bvh.nodes.size() - 1;
    }

    BVH& m_bvh;
    BuilderKind m_kind;
    float m_alpha;
    const Platform& m_p;
    std::vector<Ref> m_refs;
    std::vector<AABB> m_right;
    Bin m_bins[3][kBins];
    float m_minOverlap = 0.0f;
    int m_dups = 0;
};

} // namespace

void build_bvh(BVH& bvh, BuilderKind kind, float splitAlpha)
{
    auto t0 = std::chrono::steady_clock::now();
    bvh.nodes.clear();
    bvh.triIndices.clear();
    Builder b(bvh, kind, splitAlpha);
    bvh.root = b.run();
    bvh.buildSeconds = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count();
}

// =========================================================================================
// SAH metric — BVHNode.cpp:79-94 (fp32 accumulation in DFS order, children 0 then 1)
// =========================================================================================
namespace {
void sah_rec(const BVH& bvh, int ni, float prob, float& sah, int depth, TreeStats& st)
{
    const Node& n = bvh.nodes[ni];
    int nc = n.leaf ? 0 : 2;
    int nt = n.leaf ? (n.hi - n.lo) : 0;
    sah += prob * bvh.platform.cost(nc, nt);
    st.maxDepth = std::max(st.maxDepth, depth);
    if (n.leaf) { st.numLeaf++; st.numTris += nt; return; }
    st.numInner++;
    for (int i = 0; i < 2; i++) {
        const Node& c = bvh.nodes[n.child[i]];
        float cp = 0.0f;
        if (prob > 0.0f) cp = prob * c.bounds.area() / n.bounds.area();
        sah_rec(bvh, n.child[i], cp, sah, depth + 1, st);
    }
}
}

TreeStats tree_stats(const BVH& bvh)
{
    TreeStats st{0.0f, 0, 0, 0, 0};
    float sah = 0.0f;
    sah_rec(bvh, bvh.root, 1.0f, sah, 1, st);    // getSubtreeSize(MAX_DEPTH) counts the root as 1
    st.sah = sah;
    return st;
}

// =========================================================================================
// BVH::trace — BVH.cpp:90-186
// =========================================================================================
namespace {
struct TraceCtx { const BVH* bvh; bool closest; uint32_t* cnt; };

void trace_rec(const TraceCtx& c, int ni, Ray& ray, RayResult& res)
{
    const BVH& bvh = *c.bvh;
    const Node& node = bvh.nodes[ni];
    if (node.leaf) {
        if (c.cnt) { c.cnt[2]++; c.cnt[1] += bvh.platform.roundTri(node.hi - node.lo); }
        for (int i = node.lo; i < node.hi; i++) {
            int index = bvh.triIndices[i];
            float t = ray_triangle(bvh.scene.v(index, 0), bvh.scene.v(index, 1), bvh.scene.v(index, 2), ray);
            if (t > ray.tmin && t < ray.tmax) {
                ray.tmax = t;
                res.t = t;
                res.id = index;
                if (!c.closest) return;
            }
        }
    } else {
        if (c.cnt) c.cnt[0]++;
        int c0 = node.child[0], c1 = node.child[1];
        Span s0 = ray_box(bvh.nodes[c0].bounds, ray);
        Span s1 = ray_box(bvh.nodes[c1].bounds, ray);
        bool i0 = (s0.tmin <= s0.tmax) && (s0.tmax >= ray.tmin) && (s0.tmin <= ray.tmax);
        bool i1 = (s1.tmin <= s1.tmax) && (s1.tmax >= ray.tmin) && (s1.tmin <= ray.tmax);
        if (i0 && i1 && s0.tmin > s1.tmin) { std::swap(s0, s1); std::swap(c0, c1); std::swap(i0, i1); }
        // note: after the swap both flags are true, so swapping them is a no-op kept for symmetry
        if (i0) trace_rec(c, c0, ray, res);
        if (res.id != -1 && !c.closest) return;
        if (i1) trace_rec(c, c1, ray, res);
    }
}
}

void trace_tree(const BVH& bvh, const Ray* rays, RayResult* results, int n, bool closest, uint32_t* counters, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
#endif
    for (int i = 0; i < n; i++) {
        Ray ray = rays[i];
        RayResult& res = results[i];
        res.id = -1;                                // RayResult::clear touches id only (Util.hpp:81)
        if (counters) { counters[3 * i] = counters[3 * i + 1] = counters[3 * i + 2] = 0; }
        TraceCtx c{&bvh, closest, counters ? counters + 3 * i : nullptr};
        trace_rec(c, bvh.root, ray, res);
    }
}

// =========================================================================================
// createCompact / woopifyTri — CudaBVH.cpp:579-687 (shuffle() deliberately not restated)
// =========================================================================================
void woopify_tri(V3 v0, V3 v1, V3 v2, float out[12])
{
    M4 mtx;
    V3 c0 = v0 - v2, c1 = v1 - v2, c2 = cross(v0 - v2, v1 - v2);
    for (int r = 0; r < 3; r++) { mtx.m[r][0] = c0[r]; mtx.m[r][1] = c1[r]; mtx.m[r][2] = c2[r]; mtx.m[r][3] = v2[r]; }
    mtx.m[3][0] = mtx.m[3][1] = mtx.m[3][2] = 0.0f; mtx.m[3][3] = 1.0f;
    M4 inv = invert(mtx);
    out[0] = inv.m[2][0]; out[1] = inv.m[2][1]; out[2] = inv.m[2][2]; out[3] = -inv.m[2][3];
    for (int j = 0; j < 4; j++) { out[4 + j] = inv.m[0][j]; out[8 + j] = inv.m[1][j]; }
}

void create_compact(const BVH& bvh, CompactBVH& out, int div)
{
    out.nodes.assign(16, 0);
    out.woop.clear();
    out.triIndex.clear();
    struct Entry { int node; int idx; };          // idx in units of int4
    std::vector<Entry> stack;
    stack.push_back({bvh.root, 0});

    while (!stack.empty()) {
        Entry e = stack.back();
        stack.pop_back();
        const Node& n = bvh.nodes[e.node];
        const AABB* cbox[2];
        int cidx[2];
        for (int i = 0; i < 2; i++) {
            const Node& ch = bvh.nodes[n.child[i]];
            cbox[i] = &ch.bounds;
            if (!ch.leaf) {
                cidx[i] = (int)(out.nodes.size() * 4) / div;
                stack.push_back({n.child[i], (int)out.nodes.size() / 4});
                out.nodes.resize(out.nodes.size() + 16, 0);
                continue;
            }
            cidx[i] = ~((int)out.woop.size() / 4);
            for (int j = ch.lo; j < ch.hi; j++) {
                int tri = bvh.triIndices[j];
                float w[12];
                woopify_tri(bvh.scene.v(tri, 0), bvh.scene.v(tri, 1), bvh.scene.v(tri, 2), w);
                if (w[0] == 0.0f) w[0] = 0.0f;     // -0 -> +0 so it cannot alias the terminator
                for (int k = 0; k < 12; k++) out.woop.push_back((int32_t)f2u(w[k]));
                out.triIndex.push_back(tri);
                out.triIndex.push_back(0);
                out.triIndex.push_back(0);
            }
            for (int k = 0; k < 4; k++) out.woop.push_back((int32_t)0x80000000u);
            out.triIndex.push_back(0);
        }
        int32_t* dst = &out.nodes[(size_t)e.idx * 4];
        auto fb = [](float f) { return (int32_t)f2u(f); };
        dst[0] = fb(cbox[0]->mn.x); dst[1] = fb(cbox[0]->mx.x); dst[2] = fb(cbox[0]->mn.y); dst[3] = fb(cbox[0]->mx.y);
        dst[4] = fb(cbox[1]->mn.x); dst[5] = fb(cbox[1]->mx.x); dst[6] = fb(cbox[1]->mn.y); dst[7] = fb(cbox[1]->mx.y);
        dst[8] = fb(cbox[0]->mn.z); dst[9] = fb(cbox[0]->mx.z); dst[10] = fb(cbox[1]->mn.z); dst[11] = fb(cbox[1]->mx.z);
        dst[12] = cidx[0]; dst[13] = cidx[1]; dst[14] = (n.axis | (n.splitType << 2)); dst[15] = 0;
    }
}

// createNodeBasic / createTriWoopBasic / createTriIndexBasic — CudaBVH.cpp:453-575.  Every node, leaves included, is a
// 64-byte record numbered in the order the depth-first stack hands indices out (both children at once, right child
// popped first); leaves carry their bounds twice and (lo, hi) in the link words.  Note the -0.0f guard of createCompact
// is NOT applied here (the basic layouts have no terminator to alias).
void create_basic(const BVH& bvh, int layout, CompactBVH& out)
{
    const bool nodeSOA = (layout == 2 || layout == 3), triSOA = (layout == 1 || layout == 3);
    const size_t Align = 4096;
    size_t numNodes = 0;
    {   // subtree node count from the root (the node pool may hold unreachable entries)
        std::vector<int> st{bvh.root};
        while (!st.empty()) { int n = st.back(); st.pop_back(); numNodes++; if (!bvh.nodes[n].leaf) { st.push_back(bvh.nodes[n].child[0]); st.push_back(bvh.nodes[n].child[1]); } }
    }
    const size_t nodeBytes = (numNodes * 64 + Align - 1) & ~(Align - 1);
    out.nodes.assign(nodeBytes / 4, 0);
    struct Entry { int node; int idx; };
    int next = 0;
    std::vector<Entry> stack;
    stack.push_back({bvh.root, next++});
    auto fb = [](float f) { return (int32_t)f2u(f); };
    while (!stack.empty()) {
        Entry e = stack.back();
        stack.pop_back();
        const Node& n = bvh.nodes[e.node];
        const AABB *b0, *b1;
        int c0, c1, split = 0;
        if (n.leaf) { b0 = b1 = &n.bounds; c0 = n.lo; c1 = n.hi; }
        else {
            Entry e0{n.child[0], next++}, e1{n.child[1], next++};
            stack.push_back(e0); stack.push_back(e1);
            b0 = &bvh.nodes[e0.node].bounds; b1 = &bvh.nodes[e1.node].bounds;
            c0 = bvh.nodes[e0.node].leaf ? ~e0.idx : e0.idx;
            c1 = bvh.nodes[e1.node].leaf ? ~e1.idx : e1.idx;
            split = (n.axis | (n.splitType << 2));
        }
        const int32_t data[16] = {fb(b0->mn.x), fb(b0->mx.x), fb(b0->mn.y), fb(b0->mx.y), fb(b1->mn.x), fb(b1->mx.x), fb(b1->mn.y), fb(b1->mx.y),
                                  fb(b0->mn.z), fb(b0->mx.z), fb(b1->mn.z), fb(b1->mx.z), c0, c1, split, 0};
        for (int j = 0; j < 4; j++) {
            int32_t* dst = nodeSOA ? &out.nodes[(size_t)e.idx * 4 + (nodeBytes / 16) * j] : &out.nodes[(size_t)e.idx * 16 + 4 * j];
            for (int k = 0; k < 4; k++) dst[k] = data[4 * j + k];
        }
    }
    const size_t R = bvh.triIndices.size();
    const size_t woopBytes = (R * 64 + Align - 1) & ~(Align - 1);
    out.woop.assign(woopBytes / 4, 0);
    out.triIndex.assign(R, 0);
    for (size_t i = 0; i < R; i++) {
        int tri = bvh.triIndices[i];
        float w[12];
        woopify_tri(bvh.scene.v(tri, 0), bvh.scene.v(tri, 1), bvh.scene.v(tri, 2), w);
        for (int j = 0; j < 3; j++) {
            int32_t* dst = triSOA ? &out.woop[i * 4 + (woopBytes / 16) * j] : &out.woop[i * 16 + 4 * j];
            for (int k = 0; k < 4; k++) dst[k] = (int32_t)f2u(w[4 * j + k]);
        }
        out.triIndex[i] = tri;
    }
}

// =========================================================================================
// CudaBVH::trace<BVHLayout_Compact> — CudaBVH.cpp:698-784, 1083-1126, 1183-1225, 1251-1265
// =========================================================================================
namespace {
inline void compact_children(const int32_t* nodes, int addr, AABB& b0, AABB& b1, int& a0, int& a1)
{
    const int32_t* w = nodes + addr / 4;
    auto f = [&](int i) { return u2f((uint32_t)w[i]); };
    b0.mn = V3(f(0), f(2), f(8));  b0.mx = V3(f(1), f(3), f(9));
    b1.mn = V3(f(4), f(6), f(10)); b1.mx = V3(f(5), f(7), f(11));
    a0 = w[12]; a1 = w[13];
}

void trace_compact_one(const int32_t* nodes, const int32_t* woop, const int32_t* triIndex,
                       Ray& ray, RayResult& res, bool closest, uint32_t* cnt)
{
    int stack[100];
    int sp = 1;
    int node = 0;
    while (sp > 0) {
        for (;;) {
            if (node < 0) {
                if (cnt) cnt[2]++;
                bool end = false;
                for (int triAddr = (-node - 1);; triAddr += 3) {
                    if ((uint32_t)woop[triAddr * 4] == 0x80000000u) break;
                    if (cnt) cnt[1]++;
                    int index = triIndex[triAddr];
                    const float* z = reinterpret_cast<const float*>(woop + triAddr * 4);
                    float t = ray_triangle_woop(z, z + 4, z + 8, ray);
                    if (t > ray.tmin && t < ray.tmax) {       // updateHit
                        ray.tmax = t;
                        res.t = t;
                        res.id = index;
                        if (!closest) { end = true; break; }
                    }
                }
                if (end) return;
                break;
            }
            AABB b0, b1; int a0, a1;
            compact_children(nodes, node, b0, b1, a0, a1);
            Span s0 = ray_box(b0, ray), s1 = ray_box(b1, ray);
            bool i0 = (s0.tmin <= s0.tmax) && (s0.tmax >= ray.tmin) && (s0.tmin <= ray.tmax);
            bool i1 = (s1.tmin <= s1.tmax) && (s1.tmax >= ray.tmin) && (s1.tmin <= ray.tmax);
            if (cnt) cnt[0]++;
            if (i0 && i1) {
                if (s0.tmin > s1.tmin) { std::swap(s0, s1); std::swap(a0, a1); }
                node = a0;
                stack[sp++] = a1;
            } else if (i0) node = a0;
            else if (i1) node = a1;
            else break;
        }
        sp--;
        node = stack[sp];
    }
}
}

void trace_compact(const int32_t* nodes, const int32_t* woop, const int32_t* triIndex,
                   const Ray* rays, RayResult* results, int n, bool closest, uint32_t* counters, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
#endif
    for (int i = 0; i < n; i++) {
        Ray ray = rays[i];
        RayResult& res = results[i];
        res.id = -1;
        res.t = ray.tmax;                          // CudaBVH.cpp:271-272
        if (counters) { counters[3 * i] = counters[3 * i + 1] = counters[3 * i + 2] = 0; }
        trace_compact_one(nodes, woop, triIndex, ray, res, closest, counters ? counters + 3 * i : nullptr);
    }
}

void trace_brute(const Scene& scene, const Ray* rays, RayResult* results, int n, bool closest, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
#endif
    for (int i = 0; i < n; i++) {
        Ray ray = rays[i];
        RayResult& res = results[i];
        res.id = -1; res.t = ray.tmax;
        for (int k = 0; k < scene.numTris; k++) {
            float t = ray_triangle(scene.v(k, 0), scene.v(k, 1), scene.v(k, 2), ray);
            if (t > ray.tmin && t < ray.tmax) {
                ray.tmax = t; res.t = t; res.id = k;
                if (!closest) break;
            }
        }
    }
}

// =========================================================================================
// SAH of a flat Compact tree with the BVHNode.cpp:79-94 formula.
// =========================================================================================
namespace {
struct FlatSahCtx { const int32_t* nodes; const int32_t* woop; const Platform* p; int inner, leaf, tris, maxDepth; };

void flat_sah(FlatSahCtx& c, int addr, const AABB& box, float prob, float& sah, int depth)
{
    c.maxDepth = std::max(c.maxDepth, depth);
    if (addr < 0) {
        int nt = 0;
        for (int a = ~addr; (uint32_t)c.woop[a * 4] != 0x80000000u; a += 3) nt++;
        sah += prob * c.p->cost(0, nt);
        c.leaf++; c.tris += nt;
        return;
    }
    sah += prob * c.p->cost(2, 0);
    c.inner++;
    AABB b[2]; int a[2];
    compact_children(c.nodes, addr, b[0], b[1], a[0], a[1]);
    for (int i = 0; i < 2; i++) {
        float cp = 0.0f;
        if (prob > 0.0f) cp = prob * b[i].area() / box.area();
        flat_sah(c, a[i], b[i], cp, sah, depth + 1);
    }
}
}

float compact_sah(const int32_t* nodes, const int32_t* woop, const Platform& p, int* numInner, int* numLeaf, int* numTris, int* maxDepth)
{
    FlatSahCtx c{nodes, woop, &p, 0, 0, 0, 0};
    AABB b0, b1; int a0, a1;
    compact_children(nodes, 0, b0, b1, a0, a1);
    AABB root = b0; root.grow(b1);
    float sah = 0.0f;
    flat_sah(c, 0, root, 1.0f, sah, 1);
    if (numInner) *numInner = c.inner;
    if (numLeaf) *numLeaf = c.leaf;
    if (numTris) *numTris = c.tris;
    if (maxDepth) *maxDepth = c.maxDepth;
    return sah;
}

} // namespace orc
