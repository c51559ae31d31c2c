// ntrace_b200 — GPU LBVH builder for sm_100a, emitting the Compact traversal layout directly.
//
// Replaces the reference pipeline
//   calcMorton            src/rt/bvh/HLBVH/emitTreeKernel.cu:655-691
//   thrust::sort_by_key   src/rt/bvh/HLBVH/radixSort.cu:22-46          (hand-written stable LSD radix sort here)
//   emitTreeKernel x <=30 src/rt/bvh/HLBVH/emitTreeKernel.cu:233-381   (one launch per level + host readback each)
//   createLeaf            src/rt/bvh/HLBVH/emitTreeKernel.cu:170-231
//   calcAABB x levels     src/rt/bvh/HLBVH/emitTreeKernel.cu:417-562
//   calcWoopKernel        src/rt/bvh/HLBVH/emitTreeKernel.cu:574-645
// driven by HLBVHBuilder::buildLBVH (src/rt/bvh/HLBVH/HLBVHBuilder.cpp:451-593).
//
// The reference emits the tree level by level through queues.  Here the same tree (the same
// (start, split, end) range hierarchy) is produced with one thread per *gap*
// between consecutive sorted keys (Karras-style, fully parallel, no per-level launches):
//   * a gap whose two keys differ in top bit b is the split of the node spanning the maximal range
//     around it whose keys agree above bit b (found by galloping + binary search on the keys);
//   * gaps inside a run of identical keys follow the reference's recursive median rule
//     split = (s + e) >> 1, which is a pure function of the run bounds;
//   * a child range with <= leafSize triangles is a leaf; nodes 29 levels below the root force both
//     children to leaves (reference: `oldLevel == 0`), which only long duplicate runs can reach;
//   * inner nodes and leaves are numbered by one 64-bit exclusive scan (nodes by gap index, root
//     moved to slot 0; leaves by sorted position), so the output is deterministic and Morton-coherent;
//   * AABBs are fitted bottom-up in the same kernel that emits the nodes, with one arrival counter per
//     node (one acq_rel atomic per arrival), writing child boxes straight into the parent's node words.
// One host readback (node / leaf counts, to size the output buffers) instead of one per level.
#include "nt_common.cuh"
#include "nt_sort.cuh"
#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>

namespace nt {

namespace {

constexpr float kF32Max = 3.402823466e+38f;

// ------------------------------------------------------------------------------------------------
// Morton codes — bit-exact with the IEEE restatement: explicit _rn intrinsics, no FMA contraction.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint spread10(uint n)
{
    n &= 0x3ffu;
    n = (n ^ (n << 16)) & 0xff0000ffu;
    n = (n ^ (n << 8)) & 0x0300f00fu;
    n = (n ^ (n << 4)) & 0x030c30c3u;
    return (n ^ (n << 2)) & 0x09249249u;
}

__device__ __forceinline__ int quantise(float mid, float lo, float step, int cells)
{
    // (int)floorf((mid - lo) / step), then the reference's clamp through the float overload
    const int q = (int)floorf(__fdiv_rn(__fsub_rn(mid, lo), step));      // cvt.rzi saturates, NaN -> 0
    return (int)fmaxf(0.0f, fminf((float)q, (float)(cells - 1)));
}

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 ld3(const float* p, int i) { F3 r; r.x = __ldg(p + 3 * i); r.y = __ldg(p + 3 * i + 1); r.z = __ldg(p + 3 * i + 2); return r; }
__device__ __forceinline__ F3 min3v(F3 a, F3 b) { F3 r; r.x = fminf(a.x, b.x); r.y = fminf(a.y, b.y); r.z = fminf(a.z, b.z); return r; }
__device__ __forceinline__ F3 max3v(F3 a, F3 b) { F3 r; r.x = fmaxf(a.x, b.x); r.y = fmaxf(a.y, b.y); r.z = fmaxf(a.z, b.z); return r; }

__global__ void __launch_bounds__(256) morton_kernel(const float* __restrict__ verts, const int* __restrict__ tris, int n,
                                                      float lox, float loy, float loz, float sx, float sy, float sz,
                                                      uint* __restrict__ keys, int* __restrict__ idx,
                                                      u64* __restrict__ zeroPack, int* __restrict__ zeroCounters)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    zeroPack[t] = 0; zeroCounters[t] = 0;        // scan input and refit arrival counters of this position (saves two memsets)
    const F3 a = ld3(verts, __ldg(tris + 3 * t)), b = ld3(verts, __ldg(tris + 3 * t + 1)), c = ld3(verts, __ldg(tris + 3 * t + 2));
    const F3 lo = min3v(a, min3v(b, c)), hi = max3v(a, max3v(b, c));
    const float mx = __fadd_rn(lo.x, __fdiv_rn(__fsub_rn(hi.x, lo.x), 2.0f));
    const float my = __fadd_rn(lo.y, __fdiv_rn(__fsub_rn(hi.y, lo.y), 2.0f));
    const float mz = __fadd_rn(lo.z, __fdiv_rn(__fsub_rn(hi.z, lo.z), 2.0f));
    const uint qx = (uint)quantise(mx, lox, sx, 1024), qy = (uint)quantise(my, loy, sy, 1024), qz = (uint)quantise(mz, loz, sz, 1024);
    keys[t] = spread10(qx) | (spread10(qy) << 1) | (spread10(qz) << 2);
    idx[t] = t;
}

// ------------------------------------------------------------------------------------------------
// Topology: one thread per gap g in [0, N-2] (between sorted positions g and g+1).
//
// `hb` = number of Morton bits the gap-parallel emitter owns: 30 for plain LBVH (whole array is one "cluster"),
// 3*hlbvhBits for HLBVH, where gaps whose keys differ in a bit >= hb are cluster boundaries handled by the
// top-level SAH stage and the per-cluster subtrees start at bit hb-1 (HLBVHBuilder.cpp:720, level arg :344).
// ------------------------------------------------------------------------------------------------
enum : uint { F_KEPT = 1u, F_LEFT_LEAF = 2u, F_RIGHT_LEAF = 4u, F_NEED_DEPTH = 8u };

// parent codes: >= 0 : (parent gap)*2 + side ; -1 : global root (no parent) ; <= -3 : -3 - (topNode*2 + side)
__device__ __forceinline__ int top_parent_code(int topNode, int side) { return -3 - (topNode * 2 + side); }

__device__ __forceinline__ int topbit(uint x) { return 31 - __clz(x); }

// smallest f <= from such that pred holds on [f, from]; pred is monotone (true near `from`)
template <class Pred>
__device__ __forceinline__ int gallop_left(int from, Pred pred)
{
    int good = from, step = 1;
    while (good - step >= 0 && pred(good - step)) { good -= step; step <<= 1; }
    int bad = max(good - step, -1);
    while (good - bad > 1) { const int mid = (good + bad) >> 1; if (pred(mid)) good = mid; else bad = mid; }
    return good;
}
template <class Pred>
__device__ __forceinline__ int gallop_right(int from, int n, Pred pred)
{
    int good = from, step = 1;
    while (good + step < n && pred(good + step)) { good += step; step <<= 1; }
    int bad = min(good + step, n);
    while (bad - good > 1) { const int mid = (good + bad) >> 1; if (pred(mid)) good = mid; else bad = mid; }
    return good;
}

__global__ void __launch_bounds__(256) topology_kernel(const uint* __restrict__ K, int n, int leafSize, int hb,
                                                        const uint* __restrict__ clusterOf,   // exclusive scan of cluster heads (HLBVH) or null
                                                        const int* __restrict__ clsParent,    // per cluster: topNode*2+side (HLBVH) or null
                                                        const int* __restrict__ clsStart, int clusterLeaf,   // HLBVH: clusters with <= clusterLeaf triangles are leaves
                                                        int* __restrict__ nodeS, int* __restrict__ nodeE, int* __restrict__ parent,
                                                        uint* __restrict__ flags, int* __restrict__ rootGap)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n - 1) return;
    const uint kg = __ldg(K + g), kn = __ldg(K + g + 1);
    const uint x = kg ^ kn;
    int s, e, b, par = -2;
    if (x) {
        b = topbit(x);
        if (b >= hb) { nodeS[g] = g; nodeE[g] = g + 1; parent[g] = -2; flags[g] = 0; return; }   // cluster boundary
        s = gallop_left(g, [&](int i) { return ((__ldg(K + i) ^ kg) >> b) == 0u; });
        e = gallop_right(g + 1, n, [&](int i) { return ((__ldg(K + i) ^ kn) >> b) == 0u; }) + 1;
    } else {
        b = -1;
        s = gallop_left(g, [&](int i) { return __ldg(K + i) == kg; });
        e = gallop_right(g + 1, n, [&](int i) { return __ldg(K + i) == kg; }) + 1;
        // recursive median rule inside the run (emitTreeKernel.cu:281-283); the last split passed is the parent
        for (;;) {
            const int m = (s + e) >> 1;
            if (m == g + 1) break;
            if (g + 1 < m) { e = m; par = (m - 1) * 2 + 0; }     // we are in the left part: left child of gap m-1
            else           { s = m; par = (m - 1) * 2 + 1; }
        }
    }
    bool root = false;
    if (par == -2) {
        // bounded by radix gaps: the one with the lower split bit is the parent; cluster boundaries (or the array
        // ends) on both sides mean this node is the root of its cluster
        const int bl = (s == 0) ? 32 : topbit(__ldg(K + s - 1) ^ __ldg(K + s));
        const int br = (e == n) ? 32 : topbit(__ldg(K + e - 1) ^ __ldg(K + e));
        if (bl >= hb && br >= hb) {
            root = true;
            if (clsParent) par = top_parent_code(0, 0) - clsParent[clusterOf[s]];   // -3 - (topNode*2+side)
            else { par = -1; *rootGap = g; }
        }
        else if (bl >= hb) par = (e - 1) * 2 + 0;
        else if (br >= hb) par = (s - 1) * 2 + 1;
        else par = (bl < br) ? (s - 1) * 2 + 1 : (e - 1) * 2 + 0;
    }
    const int split = g + 1;
    uint f = 0;
    bool keep = (e - s) > leafSize || (root && !clsParent);
    if (keep && clsParent) {                                   // HLBVH: nothing is emitted inside a leaf cluster
        const uint k = clusterOf[s];
        if (clsStart[k + 1] - clsStart[k] <= clusterLeaf) keep = false;
    }
    if (keep) {
        f = F_KEPT;
        if (split - s <= leafSize) f |= F_LEFT_LEAF;
        if (e - split <= leafSize) f |= F_RIGHT_LEAF;
        // A radix node splitting bit b sits at depth <= (hb-1) - b below its cluster root, so only two kinds of node
        // can be affected by the level limit: duplicate-run nodes (b == -1; they may not exist at all if they are
        // too deep) and bit-0 nodes that would otherwise keep an inner child.
        const bool bothLeaves = (f & (F_LEFT_LEAF | F_RIGHT_LEAF)) == (F_LEFT_LEAF | F_RIGHT_LEAF);
        if (b < 0 || (b == 0 && !bothLeaves)) f |= F_NEED_DEPTH;
    }
    f |= (uint)(b + 1) << 8;
    nodeS[g] = s; nodeE[g] = e; parent[g] = par; flags[g] = f;
}

// forced leaves hb-1 levels below the cluster root (emitTreeKernel.cu:289-292 `oldLevel == 0`), then scan inputs:
// pack[i].lo = gap i is an inner node, pack[i].hi = a leaf starts at sorted position i
__global__ void __launch_bounds__(256) finalize_kernel(int n, int hb, const int* __restrict__ nodeS, const int* __restrict__ parent,
                                                        uint* __restrict__ flags, uint* __restrict__ pack32 /* null: flags only */)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n - 1) return;
    uint f = flags[g];
    if (f & F_NEED_DEPTH) {
        int depth = 0;
        for (int p = parent[g]; p >= 0; p = parent[p >> 1]) depth++;
        if (depth >= hb) f &= ~(F_KEPT | F_LEFT_LEAF | F_RIGHT_LEAF);
        else if (depth == hb - 1) f |= F_LEFT_LEAF | F_RIGHT_LEAF;
        f &= ~F_NEED_DEPTH;
        flags[g] = f;
    }
    if (pack32 && (f & F_KEPT)) {
        pack32[2 * g] = 1u;
        if (f & F_LEFT_LEAF) pack32[2 * nodeS[g] + 1] = 1u;
        if (f & F_RIGHT_LEAF) pack32[2 * (g + 1) + 1] = 1u;
    }
}

__global__ void __launch_bounds__(256) pack_kernel(int n, const int* __restrict__ nodeS, const uint* __restrict__ flags, uint* __restrict__ pack32)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n - 1) return;
    const uint f = flags[g];
    if (f & F_KEPT) {
        pack32[2 * g] = 1u;
        if (f & F_LEFT_LEAF) pack32[2 * nodeS[g] + 1] = 1u;
        if (f & F_RIGHT_LEAF) pack32[2 * (g + 1) + 1] = 1u;
    }
}

// Final node numbering.  LBVH: gap nodes by scan rank with the root moved to slot 0.  HLBVH: the numTop top-level
// nodes come first (root = top node 0), gap nodes follow in scan order.
struct Numbering { int numTop; int rootGap; uint rootRank; int linkMul; };   // linkMul: inner-child link per node index, 64 (Compact, byte offset) or 4 (Compact2, offset / 16)
__device__ __forceinline__ int gap_node_id(const Numbering& nb, int g, uint rank)
{
    if (nb.numTop > 0) return nb.numTop + (int)rank;
    return (g == nb.rootGap) ? 0 : (int)(rank < nb.rootRank ? rank + 1 : rank);
}

// Per-triangle boxes in SORTED order (6 floats: min xyz, max xyz of the three vertices, no epsilon).  The geometry is gathered
// once -- vertices are stored in input order, so a gather in Morton order fetches ~270 B of DRAM sectors for the 36 B it wants --
// and every later consumer (cluster boxes, leaf boxes, collapse) streams these 24 B instead.
__device__ __forceinline__ void tri_box_store(float* __restrict__ triBox, int p, F3 a, F3 b, F3 c)
{
    const F3 mn = min3v(a, min3v(b, c)), mx = max3v(a, max3v(b, c));
    float2* o = reinterpret_cast<float2*>(triBox + (size_t)p * 6);
    o[0] = make_float2(mn.x, mn.y); o[1] = make_float2(mn.z, mx.x); o[2] = make_float2(mx.y, mx.z);
}

__global__ void __launch_bounds__(256) tri_box_kernel(int n, const float* __restrict__ verts, const int* __restrict__ tris,
                                                       const int* __restrict__ idx, float* __restrict__ triBox)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int t = __ldg(idx + p);
    tri_box_store(triBox, p, ld3(verts, __ldg(tris + 3 * t)), ld3(verts, __ldg(tris + 3 * t + 1)), ld3(verts, __ldg(tris + 3 * t + 2)));
}

// leaf box = union of triangle vertices -/+ epsilon (calcLeaf, emitTreeKernel.cu:383-408).  The reference subtracts / adds epsilon
// per triangle before the min / max; rounding is monotone, so min_i fl(m_i - eps) == fl(min_i m_i - eps) bit for bit and the
// epsilon is applied once to the union here.
__device__ __forceinline__ void leaf_box(const float* __restrict__ triBox, int a, int b, float eps, F3& lo, F3& hi)
{
    lo.x = lo.y = lo.z = kF32Max; hi.x = hi.y = hi.z = -kF32Max;
    for (int i = a; i < b; i++) {
        const float2* q = reinterpret_cast<const float2*>(triBox + (size_t)i * 6);
        const float2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        lo.x = fminf(lo.x, q0.x); lo.y = fminf(lo.y, q0.y); lo.z = fminf(lo.z, q1.x);
        hi.x = fmaxf(hi.x, q1.y); hi.y = fmaxf(hi.y, q2.x); hi.z = fmaxf(hi.z, q2.y);
    }
    if (b > a) {
        lo.x = __fsub_rn(lo.x, eps); lo.y = __fsub_rn(lo.y, eps); lo.z = __fsub_rn(lo.z, eps);
        hi.x = __fadd_rn(hi.x, eps); hi.y = __fadd_rn(hi.y, eps); hi.z = __fadd_rn(hi.z, eps);
    }
}

__device__ __forceinline__ void store_child_box(float* node, int side, F3 lo, F3 hi)
{
    // node words: c0 -> 0..3, 8, 9 ; c1 -> 4..7, 10, 11  (CudaBVH.hpp:43-47)
    *reinterpret_cast<float4*>(node + 4 * side) = make_float4(lo.x, hi.x, lo.y, hi.y);
    *reinterpret_cast<float2*>(node + 8 + 2 * side) = make_float2(lo.z, hi.z);
}

struct ClimbCtx {
    int* nodes; const int* parent; const u64* ex; int* gapCounters;
    const int* topParent; int* topCounters; Numbering nb;
};

// One arrival at a refit counter: a single acq_rel RMW releases this thread's child-box stores and acquires the sibling's
// (replaces membar.gl + relaxed atom; the fused form is cheaper and is the exact ordering the climb needs).
__device__ __forceinline__ int arrive(int* counter)
{
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
    return old;
}

// Bottom-up refit: node `curId` (both child boxes present) carries its box to its parent; whoever arrives second at a
// node continues (one arrival counter per node).  parCode is the parent code of the current node.
__device__ __forceinline__ void climb(const ClimbCtx& c, int curId, int parCode)
{
    for (;;) {
        if (parCode == -1) return;
        const float* cb = reinterpret_cast<const float*>(c.nodes + (size_t)curId * 16);
        const float4 b0 = __ldcg(reinterpret_cast<const float4*>(cb));
        const float4 b1 = __ldcg(reinterpret_cast<const float4*>(cb + 4));
        const float4 bz = __ldcg(reinterpret_cast<const float4*>(cb + 8));
        F3 lo, hi;
        lo.x = fminf(b0.x, b1.x); hi.x = fmaxf(b0.y, b1.y);
        lo.y = fminf(b0.z, b1.z); hi.y = fmaxf(b0.w, b1.w);
        lo.z = fminf(bz.x, bz.z); hi.z = fmaxf(bz.y, bz.w);
        int pid, side, next; int* counter;
        if (parCode >= 0) {
            const int pg = parCode >> 1; side = parCode & 1;
            pid = gap_node_id(c.nb, pg, (uint)c.ex[pg]);
            counter = c.gapCounters + pg; next = c.parent[pg];
        } else {
            const int code = -3 - parCode; side = code & 1;
            pid = code >> 1;
            counter = c.topCounters + pid; next = c.topParent[pid];
        }
        store_child_box(reinterpret_cast<float*>(c.nodes + (size_t)pid * 16), side, lo, hi);
        if (arrive(counter) == 0) return;
        curId = pid; parCode = next;
    }
}

// One thread per KEPT gap (= inner node), taken from the rank-ordered list leaf_emit_kernel leaves behind: with leafSize 8
// only ~18 % of the gaps survive, and a thread per gap ran this latency-bound kernel with 6 of 32 lanes alive.
// The number of kept gaps and of top-level nodes is read from the device scalars (no host readback before the launch: the grid
// covers every gap and the surplus threads leave).
__global__ void __launch_bounds__(256) emit_kernel(const int* __restrict__ keptGap, const int* __restrict__ nodeS,
                                                    const int* __restrict__ nodeE, const uint* __restrict__ flags,
                                                    const int* __restrict__ scalars, const float* __restrict__ triBox, float eps, ClimbCtx c)
{
    const int rank = blockIdx.x * blockDim.x + threadIdx.x;
    if (rank >= scalars[2]) return;              // low word of the scan total = kept gaps
    const int g = __ldg(keptGap + rank);
    const uint f = flags[g];
    c.nb.rootGap = scalars[0];
    c.nb.numTop = c.topParent ? scalars[5] : 0;
    c.nb.rootRank = (c.nb.numTop > 0) ? 0u : (uint)c.ex[c.nb.rootGap];
    const int id = gap_node_id(c.nb, g, (uint)rank);
    const int s = nodeS[g], e = nodeE[g], split = g + 1;
    int* node = c.nodes + (size_t)id * 16;
    float* nodef = reinterpret_cast<float*>(node);

    const int b = (int)((f >> 8) & 0xffu) - 1;
    node[14] = (b < 0) ? -1 : (b % 3);           // `level % 3` with C remainder semantics (emitTreeKernel.cu:378)
    node[15] = 0;
    const int par = c.parent[g];
    if (par >= 0) {
        const int pg = par >> 1;
        c.nodes[(size_t)gap_node_id(c.nb, pg, (uint)c.ex[pg]) * 16 + 12 + (par & 1)] = id * c.nb.linkMul;   // byte offset (Compact) or / 16 (Compact2)
    } else if (par <= -3) {
        const int code = -3 - par;
        c.nodes[(size_t)(code >> 1) * 16 + 12 + (code & 1)] = id * c.nb.linkMul;                               // cluster root under a top-level node
    }

    int arrivals = 0;
    if (f & F_LEFT_LEAF) {
        node[12] = ~(3 * s + (int)(c.ex[s] >> 32));
        F3 lo, hi; leaf_box(triBox, s, split, eps, lo, hi);
        store_child_box(nodef, 0, lo, hi);
        arrivals++;
    }
    if (f & F_RIGHT_LEAF) {
        node[13] = ~(3 * split + (int)(c.ex[split] >> 32));
        F3 lo, hi; leaf_box(triBox, split, e, eps, lo, hi);
        store_child_box(nodef, 1, lo, hi);
        arrivals++;
    }
    if (arrivals == 0) return;
    if (arrivals == 1) {
        if (arrive(c.gapCounters + g) == 0) return;          // the inner child has not arrived yet
    }
    climb(c, id, par);
}

// HLBVH: clusters with <= leafSize triangles hang directly under a top-level node as leaves
// (distribute, emitTreeKernel.cu:990-996): link, leaf box, then join the refit.
__global__ void __launch_bounds__(256) cluster_leaf_emit_kernel(int numClusters, int leafSize, const int* __restrict__ clsStart,
                                                                 const int* __restrict__ clsParent, const int* __restrict__ scalars,
                                                                 const float* __restrict__ triBox, float eps, ClimbCtx c)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= numClusters) return;
    const int cs = clsStart[k], ce = clsStart[k + 1];
    if (ce - cs > leafSize) return;
    c.nb.numTop = scalars[5];
    const int code = clsParent[k];
    const int pid = code >> 1, side = code & 1;
    c.nodes[(size_t)pid * 16 + 12 + side] = ~(3 * cs + (int)(c.ex[cs] >> 32));
    F3 lo, hi; leaf_box(triBox, cs, ce, eps, lo, hi);
    store_child_box(reinterpret_cast<float*>(c.nodes + (size_t)pid * 16), side, lo, hi);
    if (arrive(c.topCounters + pid) == 0) return;
    climb(c, pid, c.topParent[pid]);
}

__global__ void __launch_bounds__(256) cluster_leaf_flag_kernel(int numClusters, int leafSize, const int* __restrict__ clsStart, uint* __restrict__ pack32)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= numClusters) return;
    if (clsStart[k + 1] - clsStart[k] <= leafSize) pack32[2 * clsStart[k] + 1] = 1u;
}

// ------------------------------------------------------------------------------------------------
// HLBVH: clusters = maximal runs of equal (key >> d) (createClusters, radixSort.cu:74-120), their boxes, and the
// top-level binned SAH over cluster boxes (initBins / fillBins / findSplit / distribute, emitTreeKernel.cu:699-1027)
// executed level by level inside ONE cooperative kernel (grid.sync between phases, no host readbacks).
// Deterministic: output slots come from scans in task order and the object-split fallback ranks clusters by index,
// i.e. exactly the serial schedule the CPU restatement uses.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int f2i_ord(float f) { const int i = __float_as_int(f); return (i >= 0) ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float i2f_ord(int i) { return __int_as_float((i >= 0) ? i : i ^ 0x7FFFFFFF); }

__global__ void __launch_bounds__(256) cluster_mark_kernel(const uint* __restrict__ K, int n, int d, uint* __restrict__ head)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    head[p] = (p == 0 || (K[p] >> d) != (K[p - 1] >> d)) ? 1u : 0u;
}

// clusterOf[p] (inclusive count - 1) is written in place of the exclusive scan; clsStart[k] = first position of cluster k
__global__ void __launch_bounds__(256) cluster_start_kernel(int n, const uint* __restrict__ head, uint* __restrict__ exToClusterOf,
                                                             int* __restrict__ clsStart, int numClusters)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint k = exToClusterOf[p] + head[p] - 1u;
    exToClusterOf[p] = k;
    if (head[p]) clsStart[k] = p;
    if (p == 0) clsStart[numClusters] = n;
}

__global__ void __launch_bounds__(256) cluster_box_init_kernel(int numClusters, int* __restrict__ boxI)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numClusters * 6) return;
    boxI[i] = ((i % 6) < 3) ? f2i_ord(kF32Max) : f2i_ord(-kF32Max);
}

__global__ void __launch_bounds__(256) cluster_box_kernel(int n, const uint* __restrict__ clusterOf, const float* __restrict__ triBox,
                                                           int* __restrict__ boxI)
{
    // Sorted positions of one cluster are contiguous, so the lanes of a warp form a few contiguous segments: reduce each
    // segment with shuffles and let only its first lane touch memory (ordered-int atomics: the result is order independent).
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = p < n;
    uint cid = 0xffffffffu;
    float lo[3] = {kF32Max, kF32Max, kF32Max}, hi[3] = {-kF32Max, -kF32Max, -kF32Max};
    if (valid) {
        cid = clusterOf[p];
        const float2* q = reinterpret_cast<const float2*>(triBox + (size_t)p * 6);
        const float2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        lo[0] = q0.x; lo[1] = q0.y; lo[2] = q1.x; hi[0] = q1.y; hi[1] = q2.x; hi[2] = q2.y;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint oc = __shfl_down_sync(0xffffffffu, cid, o);
        const bool take = (lane + o < 32) && (oc == cid);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float ol = __shfl_down_sync(0xffffffffu, lo[k], o), oh = __shfl_down_sync(0xffffffffu, hi[k], o);
            if (take) { lo[k] = fminf(lo[k], ol); hi[k] = fmaxf(hi[k], oh); }
        }
    }
    const uint pc = __shfl_up_sync(0xffffffffu, cid, 1);
    if (valid && (lane == 0 || pc != cid)) {
        int* bx = boxI + (size_t)cid * 6;
        atomicMin(bx + 0, f2i_ord(lo[0])); atomicMin(bx + 1, f2i_ord(lo[1])); atomicMin(bx + 2, f2i_ord(lo[2]));
        atomicMax(bx + 3, f2i_ord(hi[0])); atomicMax(bx + 4, f2i_ord(hi[1])); atomicMax(bx + 5, f2i_ord(hi[2]));
    }
}

constexpr int kBins = 8;            // BIN_CNT (emitTreeKernel.cuh:9)
constexpr int kTopThreads = 256;
constexpr int kSmemTasksDefault = 256;   // levels with at most this many tasks bin through (dynamic) shared memory first: 168 ints per task
constexpr int kTrackFirst = 64;     // tasks with at most this many clusters know their lowest cluster index (one atomicMin per member)

struct TopArgs {
    int C, leafSize, linkMul, smemTasks;
    const int* clsStart; const int* clsBoxI;
    int* clsTask[2];          // current / next task of each cluster (-1 = done)
    int* clsBin;              // C * 3
    int* clsParent;           // C : topNode*2 + side once the cluster terminates
    float* tBox[2];           // task boxes, 6 floats per task (lo xyz, hi xyz)
    int* tCnt[2]; int* tId[2];
    int* tFirst[2];           // lowest cluster index of each task: where the object-split fallback starts looking for its members
    int* rSplit; int* rAxis; int* rCntL; int* rCntR; int* rLocalOfs; int* rChild; float* rBoxes;   // per task, current level
    int* binBoxI; int* binCnt;
    int* blockSum;
    int* topNodes; int* topParent;
    int* scal;                // [0] numTasks, [1] written (top nodes so far)
    float sceneLo[3], sceneHi[3];
};

__device__ __forceinline__ float area3_rn(float x, float y, float z)
{
    return __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, y), __fmul_rn(y, z)), __fmul_rn(z, x)), 2.0f);
}

__device__ __forceinline__ int block_exclusive_int(int v, int* s_warp, int& total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = (lane < kTopThreads / 32) ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane < kTopThreads / 32) s_warp[lane] = x;
    }
    __syncthreads();
    const int base = (w == 0) ? 0 : s_warp[w - 1];
    total = s_warp[kTopThreads / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

// terminate-or-descend decision of one cluster (distribute, emitTreeKernel.cu:955-1027)
__device__ __forceinline__ int distribute_one(const TopArgs& a, int c, int task, bool goLeft, int cntL, int cntR, int topId)
{
    const int leafs = ((cntL <= 1) ? 2 : 0) | ((cntR <= 1) ? 1 : 0);
    const int childId = a.rChild[task];
    if (leafs == 0) return childId + (goLeft ? 0 : 1);
    const bool single = goLeft ? (leafs & 2) : (leafs & 1);
    if (!single) return childId;                         // the other side terminated: the only child task sits at childId
    a.clsParent[c] = topId * 2 + (goLeft ? 0 : 1);       // this cluster hangs under the top node as a leaf or an LBVH subtree
    return -1;
}

// fillBins for one cluster that belongs to task t (box tb = 6 floats) of the level being prepared (emitTreeKernel.cu:735-790):
// bin of the cluster's centre per axis, kept in clsBin for the coming distribute, and the bin's box / count updated.
// `cb` = the cluster's box (ordered ints), loaded by the caller before anything that depends on the task.  The three bins are computed
// first and read together, so the L2 path costs one round trip for the three axes instead of three (the phase is a chain of dependent
// loads: task label -> task record -> bins; see profiles/r2_summary.md, builder section).
__device__ __forceinline__ void fill_bins_one(const TopArgs& a, int c, int t, const float* tb, bool inSmem, int* s_binBox, int* s_binCnt,
                                              int c0, int c1, int c2, int c3, int c4, int c5)
{
    const float lo[3] = {i2f_ord(c0), i2f_ord(c1), i2f_ord(c2)}, hi[3] = {i2f_ord(c3), i2f_ord(c4), i2f_ord(c5)};
    const float t0 = tb[0], t1 = tb[1], t2 = tb[2], t3 = tb[3], t4 = tb[4], t5 = tb[5];
    const float tl[3] = {t0, t1, t2}, th[3] = {t3, t4, t5};
    int slot[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float mid = __fadd_rn(lo[k], __fdiv_rn(__fsub_rn(hi[k], lo[k]), 2.0f));
        const float step = __fdiv_rn(__fsub_rn(th[k], tl[k]), 8.0f);
        const int bid = quantise(mid, tl[k], step, kBins);
        a.clsBin[c * 3 + k] = bid;
        slot[k] = (t * 3 + k) * kBins + bid;
    }
    if (inSmem) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            int* b = s_binBox + slot[k] * 6;
            atomicMin(b + 0, c0); atomicMin(b + 1, c1); atomicMin(b + 2, c2);
            atomicMax(b + 3, c3); atomicMax(b + 4, c4); atomicMax(b + 5, c5);
            atomicAdd(s_binCnt + slot[k], 1);
        }
    } else {
        // a bin only ever shrinks / grows within a level, so a value read from L2 that already covers ours makes the atomic
        // redundant (a stale read can only cost an unnecessary atomic, never skip a necessary one)
        int2 q[3][3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int2* b = reinterpret_cast<const int2*>(a.binBoxI + (size_t)slot[k] * 6);
            q[k][0] = __ldcg(b); q[k][1] = __ldcg(b + 1); q[k][2] = __ldcg(b + 2);
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            int* b = a.binBoxI + (size_t)slot[k] * 6;
            if (c0 < q[k][0].x) atomicMin(b + 0, c0); if (c1 < q[k][0].y) atomicMin(b + 1, c1); if (c2 < q[k][1].x) atomicMin(b + 2, c2);
            if (c3 > q[k][1].y) atomicMax(b + 3, c3); if (c4 > q[k][2].x) atomicMax(b + 4, c4); if (c5 > q[k][2].y) atomicMax(b + 5, c5);
            atomicAdd(a.binCnt + slot[k], 1);
        }
    }
}

// Measured and dropped (scripts/top_level_bench.py, profiles/r2_summary.md): a hand-written grid barrier (monotonic counter, one arrival
// per CTA, acquire spin) instead of cooperative_groups' grid.sync(): 0.906 vs 0.879 ms for the bench tree -- the ~85 barriers of a build are
// not what the kernel waits for, the dependent loads inside its phases are; two CTAs per SM: 0.914-0.935 ms.
__global__ void __launch_bounds__(kTopThreads) hlbvh_top_kernel(TopArgs a)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_warp[kTopThreads / 32];
    __shared__ int s_red[2];
    extern __shared__ int s_bins[];                                  // a.smemTasks x (3 x kBins x 6 box words, then 3 x kBins counts)
    int* const s_binBox = s_bins;
    int* const s_binCnt = s_bins + a.smemTasks * 3 * kBins * 6;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gsize = gridDim.x * blockDim.x;
    int cur = 0;

    if (gtid == 0) {
        for (int k = 0; k < 3; k++) { a.tBox[0][k] = a.sceneLo[k]; a.tBox[0][3 + k] = a.sceneHi[k]; }
        a.tCnt[0][0] = a.C; a.tId[0][0] = 0; a.tFirst[0][0] = 0;
        a.topParent[0] = -1;
        a.scal[0] = 1; a.scal[1] = 1;
    }
    for (int c = gtid; c < a.C; c += gsize) a.clsTask[0][c] = 0;
    for (int i = gtid; i < 3 * kBins; i += gsize) {                       // initBins for the root task
        int* b = a.binBoxI + (size_t)i * 6;
        b[0] = b[1] = b[2] = f2i_ord(kF32Max); b[3] = b[4] = b[5] = f2i_ord(-kF32Max);
        a.binCnt[i] = 0;
    }
    grid.sync();

    // Shared-memory staging of a level's bins: while a level has few tasks every cluster hammers the same handful of bins, so
    // the block accumulates into its own copy and flushes one atomic per non-empty bin; deeper levels go straight to L2.
    auto smemBinsInit = [&](int tasks) {
        for (int i = threadIdx.x; i < tasks * 3 * kBins; i += kTopThreads) {
            int* b = s_binBox + i * 6;
            b[0] = b[1] = b[2] = f2i_ord(kF32Max); b[3] = b[4] = b[5] = f2i_ord(-kF32Max);
            s_binCnt[i] = 0;
        }
        __syncthreads();
    };
    auto smemBinsFlush = [&](int tasks) {
        __syncthreads();
        for (int i = threadIdx.x; i < tasks * 3 * kBins; i += kTopThreads) {
            const int cnt = s_binCnt[i];
            if (cnt == 0) continue;
            const int* sb = s_binBox + i * 6;
            int* b = a.binBoxI + (size_t)i * 6;
            atomicMin(b + 0, sb[0]); atomicMin(b + 1, sb[1]); atomicMin(b + 2, sb[2]);
            atomicMax(b + 3, sb[3]); atomicMax(b + 4, sb[4]); atomicMax(b + 5, sb[5]);
            atomicAdd(a.binCnt + i, cnt);
        }
    };

    // ---- fillBins of the root task (every later level is binned by the distribute phase of its parent level)
    smemBinsInit(1);
    for (int c = gtid; c < a.C; c += gsize) {
        const int2* cb = reinterpret_cast<const int2*>(a.clsBoxI + (size_t)c * 6);
        const int2 b0 = __ldg(cb), b1 = __ldg(cb + 1), b2 = __ldg(cb + 2);
        fill_bins_one(a, c, 0, a.tBox[0], true, s_binBox, s_binCnt, b0.x, b0.y, b1.x, b1.y, b2.x, b2.y);
    }
    smemBinsFlush(1);
    grid.sync();

#ifdef NT_TOP_TIMING
    unsigned long long tA = 0, tB = 0, tD = 0, t_prev, t_now; int levels = 0, smallLevels = 0; unsigned long long tSmall = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_prev));
    const unsigned long long t_start = t_prev;
#define NT_MARK(acc) do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now)); acc += t_now - t_prev; t_prev = t_now; } while (0)
#else
#define NT_MARK(acc) do { } while (0)
#endif
    for (;;) {
        const int T = *(volatile int*)&a.scal[0];
#ifdef NT_TOP_TIMING
        if (T == 0 && gtid == 0) printf("top: levels %d (T<=148: %d, %llu ns) phaseA %llu ns phaseB %llu ns distribute+fill %llu ns total %llu ns\n", levels, smallLevels, tSmall, tA, tB, tD, t_prev - t_start);
        levels++; const unsigned long long t_lvl = t_prev;
#endif
        if (T == 0) break;
        const int written = *(volatile int*)&a.scal[1];
        const int* tCnt = a.tCnt[cur]; const int* tId = a.tId[cur];
        const int* clsTask = a.clsTask[cur]; int* clsNext = a.clsTask[cur ^ 1];

        // (the bins of this level's tasks were filled by the previous level's distribute phase / the prologue)

        // ---- findSplit, phase a: per task decision + block-local offsets of the child tasks
        const int perBlock = (T + gridDim.x - 1) / gridDim.x;
        const int t0 = min(T, (int)blockIdx.x * perBlock), t1 = min(T, t0 + perBlock);
        int carry = 0;
        for (int base = t0; base < t1; base += kTopThreads) {
            const int t = base + threadIdx.x;
            int nodesNew = 0;
            if (t < t1) {
                const int* bb = a.binBoxI + (size_t)t * 3 * kBins * 6;
                const int* bc = a.binCnt + t * 3 * kBins;
                float best = kF32Max;
                int split = -1, axis = 0, cntL = 0, cntR = 0;
                float bx[12];                              // mnL, mxL, mnR, mxR
                for (int ax = 0; ax < 3; ax++) {
                    float mn[kBins - 1][3], mx[kBins - 1][3]; int cnt[kBins - 1];
                    float rl[3] = {kF32Max, kF32Max, kF32Max}, rh[3] = {-kF32Max, -kF32Max, -kF32Max};
                    int cc = 0;
                    for (int b = kBins - 1; b > 0; b--) {
                        const int* q = bb + (size_t)(ax * kBins + b) * 6;
                        for (int k = 0; k < 3; k++) { rl[k] = fminf(rl[k], i2f_ord(q[k])); rh[k] = fmaxf(rh[k], i2f_ord(q[3 + k])); mn[b - 1][k] = rl[k]; mx[b - 1][k] = rh[k]; }
                        cc += bc[ax * kBins + b]; cnt[b - 1] = cc;
                    }
                    float ll[3] = {kF32Max, kF32Max, kF32Max}, lh[3] = {-kF32Max, -kF32Max, -kF32Max};
                    cc = 0;
                    for (int b = 0; b < kBins - 1; b++) {
                        const int* q = bb + (size_t)(ax * kBins + b) * 6;
                        for (int k = 0; k < 3; k++) { ll[k] = fminf(ll[k], i2f_ord(q[k])); lh[k] = fmaxf(lh[k], i2f_ord(q[3 + k])); }
                        cc += bc[ax * kBins + b];
                        const float aL = area3_rn(__fsub_rn(lh[0], ll[0]), __fsub_rn(lh[1], ll[1]), __fsub_rn(lh[2], ll[2]));
                        const float aR = area3_rn(__fsub_rn(mx[b][0], mn[b][0]), __fsub_rn(mx[b][1], mn[b][1]), __fsub_rn(mx[b][2], mn[b][2]));
                        const float s = __fadd_rn(__fmul_rn((float)cc, aL), __fmul_rn((float)cnt[b], aR));
                        if (s < best) {
                            best = s; split = b; axis = ax; cntL = cc; cntR = cnt[b];
                            for (int k = 0; k < 3; k++) { bx[k] = ll[k]; bx[3 + k] = lh[k]; bx[6 + k] = mn[b][k]; bx[9 + k] = mx[b][k]; }
                        }
                    }
                }
                if (split == -1) {
                    // no plane separates the clusters: halve them by index (object split).  The reference leaves the
                    // right box half-assigned here (emitTreeKernel.cu:844-845); both children get the occupied bin's box.
                    for (int i = 0; i < kBins; i++)
                        if (bc[i] != 0) {
                            for (int k = 0; k < 3; k++) { bx[k] = bx[6 + k] = i2f_ord(bb[(size_t)i * 6 + k]); bx[3 + k] = bx[9 + k] = i2f_ord(bb[(size_t)i * 6 + 3 + k]); }
                            break;
                        }
                    cntR = tCnt[t] / 2; cntL = tCnt[t] - cntR;
                    split = -cntL; axis = 0;
                }
                nodesNew = (cntL > 1) + (cntR > 1);
                a.rSplit[t] = split; a.rAxis[t] = axis; a.rCntL[t] = cntL; a.rCntR[t] = cntR;
                for (int k = 0; k < 12; k++) a.rBoxes[(size_t)t * 12 + k] = bx[k];
            }
            int total;
            const int ex = block_exclusive_int(nodesNew, s_warp, total);
            if (t < t1) a.rLocalOfs[t] = carry + ex;
            carry += total;
        }
        if (threadIdx.x == 0) a.blockSum[blockIdx.x] = carry;
        grid.sync();
        NT_MARK(tA);

        // ---- findSplit, phase b: global offsets, child tasks, node links
        {
            int mine = 0, all = 0;
            for (int b = threadIdx.x; b < (int)gridDim.x; b += kTopThreads) { const int v = a.blockSum[b]; all += v; if (b < (int)blockIdx.x) mine += v; }
            int tot;
            block_exclusive_int(mine, s_warp, tot);
            if (threadIdx.x == 0) s_red[0] = tot;
            __syncthreads();
            block_exclusive_int(all, s_warp, tot);
            if (threadIdx.x == 0) s_red[1] = tot;
            __syncthreads();
        }
        const int blockBase = s_red[0], created = s_red[1];
        float* oBox = a.tBox[cur ^ 1]; int* oCnt = a.tCnt[cur ^ 1]; int* oId = a.tId[cur ^ 1]; int* oFirst = a.tFirst[cur ^ 1];
        for (int t = t0 + threadIdx.x; t < t1; t += kTopThreads) {
            const int ofs = blockBase + a.rLocalOfs[t];
            const int idN = written + ofs;
            const int cntL = a.rCntL[t], cntR = a.rCntR[t];
            const float* bx = a.rBoxes + (size_t)t * 12;
            int val = 0, l = 0, r = 0;
            if (cntL > 1) {
                l = idN * a.linkMul;
                for (int k = 0; k < 6; k++) oBox[(size_t)ofs * 6 + k] = bx[k];
                oCnt[ofs] = cntL; oId[ofs] = idN; oFirst[ofs] = (cntL <= kTrackFirst) ? 0x7fffffff : 0;
                a.topParent[idN] = top_parent_code(tId[t], 0);
                val = 1;
            }
            if (cntR > 1) {
                r = (idN + val) * a.linkMul;
                for (int k = 0; k < 6; k++) oBox[(size_t)(ofs + val) * 6 + k] = bx[6 + k];
                oCnt[ofs + val] = cntR; oId[ofs + val] = idN + val; oFirst[ofs + val] = (cntR <= kTrackFirst) ? 0x7fffffff : 0;
                a.topParent[idN + val] = top_parent_code(tId[t], 1);
            }
            a.rChild[t] = ofs;
            int* w = a.topNodes + (size_t)tId[t] * 16;
            w[12] = l; w[13] = r; w[14] = a.rAxis[t]; w[15] = 0;
        }
        // initBins for the next level's tasks: this level's bins were last read in phase a (before the barrier above); what
        // distribute consults are the per-task results in rSplit / rAxis / rCnt* and the per-cluster bin ids in clsBin
        for (int i = gtid; i < created * 3 * kBins; i += gsize) {
            int* b = a.binBoxI + (size_t)i * 6;
            b[0] = b[1] = b[2] = f2i_ord(kF32Max); b[3] = b[4] = b[5] = f2i_ord(-kF32Max);
            a.binCnt[i] = 0;
        }
        grid.sync();
        NT_MARK(tB);
        if (gtid == 0) { a.scal[0] = created; a.scal[1] = written + created; }      // read again only after the next grid.sync

        // ---- distribute + fillBins of the next level: plane splits per cluster; object-split fallback per task with clusters
        // ranked by index.  A cluster that descends is binned into its child task right away (the child boxes were written in phase b).
        const bool nextInSmem = (created <= a.smemTasks);
        if (nextInSmem) smemBinsInit(created);
        for (int c = gtid; c < a.C; c += gsize) {
            const int t = clsTask[c];
            // what does not depend on the task is requested at once: the cluster's box and its three bin ids of this level
            const int2* cb = reinterpret_cast<const int2*>(a.clsBoxI + (size_t)c * 6);
            const int2 b0 = __ldg(cb), b1 = __ldg(cb + 1), b2 = __ldg(cb + 2);
            const int bin0 = __ldcg(a.clsBin + c * 3), bin1 = __ldcg(a.clsBin + c * 3 + 1), bin2 = __ldcg(a.clsBin + c * 3 + 2);
            if (t < 0) { clsNext[c] = -1; continue; }
            const int split = a.rSplit[t];
            if (split < 0) continue;                       // handled below
            const int axis = a.rAxis[t];
            const int cntL = a.rCntL[t], cntR = a.rCntR[t], topId = tId[t];
            const bool goLeft = (axis == 0 ? bin0 : axis == 1 ? bin1 : bin2) <= split;
            const int nt = distribute_one(a, c, t, goLeft, cntL, cntR, topId);
            clsNext[c] = nt;
            if (nt >= 0) {
                if ((goLeft ? cntL : cntR) <= kTrackFirst) atomicMin(oFirst + nt, c);
                fill_bins_one(a, c, nt, oBox + (size_t)nt * 6, nextInSmem, s_binBox, s_binCnt, b0.x, b0.y, b1.x, b1.y, b2.x, b2.y);
            }
        }
        {
            const int lane = threadIdx.x & 31;
            const int warp = gtid >> 5, nwarps = gsize >> 5;
            const int* tFirst = a.tFirst[cur];
            for (int t = warp; t < T; t += nwarps) {
                if (a.rSplit[t] >= 0) continue;
                // the task's clusters, in index order, lie between its lowest member and wherever the last of its tCnt members is
                const int cntL = a.rCntL[t], cntR = a.rCntR[t], topId = tId[t], members = tCnt[t];
                int seen = 0;
                for (int base = tFirst[t] & ~31; base < a.C && seen < members; base += 32) {
                    const int c = base + lane;
                    const bool m = (c < a.C) && (clsTask[c] == t);
                    const unsigned mask = __ballot_sync(0xffffffffu, m);
                    if (m) {
                        const int rank = seen + __popc(mask & ((1u << lane) - 1u));      // arrival order == cluster index order
                        const bool goLeft = rank <= cntL - 1;
                        const int nt = distribute_one(a, c, t, goLeft, cntL, cntR, topId);
                        clsNext[c] = nt;
                        if (nt >= 0) {
                            if ((goLeft ? cntL : cntR) <= kTrackFirst) atomicMin(oFirst + nt, c);
                            const int2* cb = reinterpret_cast<const int2*>(a.clsBoxI + (size_t)c * 6);
                            const int2 b0 = __ldg(cb), b1 = __ldg(cb + 1), b2 = __ldg(cb + 2);
                            fill_bins_one(a, c, nt, oBox + (size_t)nt * 6, nextInSmem, s_binBox, s_binCnt, b0.x, b0.y, b1.x, b1.y, b2.x, b2.y);
                        }
                    }
                    seen += __popc(mask);
                }
            }
        }
        if (nextInSmem) smemBinsFlush(created);
        grid.sync();
        NT_MARK(tD);
#ifdef NT_TOP_TIMING
        if (T <= 148) { smallLevels++; tSmall += t_prev - t_lvl; }
#endif
        cur ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------
// SAH-guided collapse (optional; not part of the reference, never used for the parity rows).
// The emitter is run with leaf size 1, then subtrees are folded back into leaves bottom-up wherever a leaf is not
// more expensive than the subtree under the reference's cost model (Platform: Cn = Ct = 1, BVHNode.cpp:79-94):
//   C(leaf) = A * Ct * n ,  C(inner) = A * 2 Cn + C(left) + C(right) ,  collapse iff n <= maxLeaf and C(leaf) <= C(inner).
// One arrival counter per node, as in the refit; the decision is taken by whoever completes the node.
// ------------------------------------------------------------------------------------------------
enum : uint { F_COLLAPSED = 16u };

struct CollapseCtx {
    const int* nodeS; const int* nodeE; const int* parent; uint* flags;
    float* childBox;      // 12 floats per gap node: child 0 lo/hi, child 1 lo/hi
    float* childCost;     // 2 floats per gap node
    int* counters; int maxLeaf;
    float triCost;        // cost of one triangle test relative to one child-box test (Platform: 1)
};

__device__ __forceinline__ float box_area(const float* lo, const float* hi)
{
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

__global__ void __launch_bounds__(256) collapse_analyse_kernel(int n, CollapseCtx c, const float* __restrict__ triBox, float eps)
{
    const int g0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (g0 >= n - 1) return;
    const uint f0 = c.flags[g0];
    if (!(f0 & F_KEPT)) return;
    int arrivals = 0;
    for (int side = 0; side < 2; side++) {
        if (!(f0 & (side ? F_RIGHT_LEAF : F_LEFT_LEAF))) continue;
        const int a = side ? g0 + 1 : c.nodeS[g0], b = side ? c.nodeE[g0] : g0 + 1;
        F3 lo, hi; leaf_box(triBox, a, b, eps, lo, hi);
        float* cb = c.childBox + (size_t)g0 * 12 + side * 6;
        cb[0] = lo.x; cb[1] = lo.y; cb[2] = lo.z; cb[3] = hi.x; cb[4] = hi.y; cb[5] = hi.z;
        c.childCost[(size_t)g0 * 2 + side] = c.triCost * box_area(cb, cb + 3) * (float)(b - a);
        arrivals++;
    }
    if (arrivals == 0) return;
    if (arrivals == 1) { if (arrive(c.counters + g0) == 0) return; }
    int g = g0;
    for (;;) {
        const float* cb = c.childBox + (size_t)g * 12;
        float lo[3], hi[3];
        for (int k = 0; k < 3; k++) { lo[k] = fminf(__ldcg(cb + k), __ldcg(cb + 6 + k)); hi[k] = fmaxf(__ldcg(cb + 3 + k), __ldcg(cb + 9 + k)); }
        const float area = box_area(lo, hi);
        const int count = c.nodeE[g] - c.nodeS[g];
        float cost = 2.0f * area + __ldcg(c.childCost + (size_t)g * 2) + __ldcg(c.childCost + (size_t)g * 2 + 1);
        const int p = c.parent[g];
        if (p >= 0 && count <= c.maxLeaf && c.triCost * area * (float)count <= cost) {      // roots of clusters / of the tree never fold
            cost = c.triCost * area * (float)count;
            c.flags[g] |= F_COLLAPSED;
        }
        if (p < 0) return;
        const int pg = p >> 1, side = p & 1;
        float* pb = c.childBox + (size_t)pg * 12 + side * 6;
        for (int k = 0; k < 3; k++) { pb[k] = lo[k]; pb[3 + k] = hi[k]; }
        c.childCost[(size_t)pg * 2 + side] = cost;
        if (arrive(c.counters + pg) == 0) return;
        g = pg;
    }
}

// top-down resolution: a node below a folded ancestor disappears; a folded node becomes a leaf child of its parent
__global__ void __launch_bounds__(256) collapse_resolve_kernel(int n, const int* __restrict__ parent, const uint* __restrict__ flagsIn,
                                                                uint* __restrict__ flagsOut)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n - 1) return;
    const uint f = flagsIn[g];
    if (!(f & F_KEPT)) return;
    bool below = false;
    for (int p = parent[g]; p >= 0; p = parent[p >> 1]) if (flagsIn[p >> 1] & F_COLLAPSED) { below = true; break; }
    if (below) { atomicAnd(flagsOut + g, ~(F_KEPT | F_LEFT_LEAF | F_RIGHT_LEAF)); return; }
    if (f & F_COLLAPSED) {
        atomicAnd(flagsOut + g, ~(F_KEPT | F_LEFT_LEAF | F_RIGHT_LEAF));
        const int p = parent[g];
        atomicOr(flagsOut + (p >> 1), (p & 1) ? F_RIGHT_LEAF : F_LEFT_LEAF);
    }
}

// Woop rows of one triangle, 3x3 adjugate form of the reference (calcWoop, emitTreeKernel.cu:574-635),
// evaluated without FMA contraction so the rows are reproducible bit for bit.
__device__ __forceinline__ void calc_woop(F3 v0, F3 v1, F3 v2, float4& o0, float4& o1, float4& o2)
{
#define M(a, b) __fmul_rn(a, b)
#define S(a, b) __fsub_rn(a, b)
#define A(a, b) __fadd_rn(a, b)
    F3 c0, c1, c2;
    c0.x = S(v0.x, v2.x); c0.y = S(v0.y, v2.y); c0.z = S(v0.z, v2.z);
    c1.x = S(v1.x, v2.x); c1.y = S(v1.y, v2.y); c1.z = S(v1.z, v2.z);
    c2.x = S(M(c0.y, c1.z), M(c0.z, c1.y)); c2.y = S(M(c0.z, c1.x), M(c0.x, c1.z)); c2.z = S(M(c0.x, c1.y), M(c0.y, c1.x));
    const float m00 = S(M(c2.z, c1.y), M(c1.z, c2.y)), m01 = S(M(c2.z, c1.x), M(c1.z, c2.x)), m02 = S(M(c2.y, c1.x), M(c1.y, c2.x));
    const float dexp = A(S(M(c0.x, m00), M(c0.y, m01)), M(c0.z, m02));
    const float det = (float)(1.0 / (double)dexp);
    F3 i0, i1, i2;
    i0.x = M(m00, det); i0.y = M(-m01, det); i0.z = M(m02, det);
    i1.x = M(-S(M(c2.z, c0.y), M(c0.z, c2.y)), det); i1.y = M(S(M(c2.z, c0.x), M(c0.z, c2.x)), det); i1.z = M(-S(M(c2.y, c0.x), M(c0.y, c2.x)), det);
    i2.x = M(S(M(c1.z, c0.y), M(c0.z, c1.y)), det); i2.y = M(-S(M(c1.z, c0.x), M(c0.z, c1.x)), det); i2.z = M(S(M(c1.y, c0.x), M(c0.y, c1.x)), det);
    auto ndot = [](F3 a, F3 v) { return A(A(M(-a.x, v.x), M(-a.y, v.y)), M(-a.z, v.z)); };   // fdot(-a, v)
    o0 = make_float4(i2.x, i2.y, i2.z, -ndot(i2, v2));
    o1 = make_float4(i0.x, i0.y, i0.z, ndot(i0, v2));
    o2 = make_float4(i1.x, i1.y, i1.z, ndot(i1, v2));
    if (o0.x == 0.0f) o0.x = 0.0f;                       // -0 must not alias the terminator
#undef M
#undef S
#undef A
}

// one thread per sorted position: Woop triple + index into the leaf-ordered arrays, terminator after the last
// triangle of each leaf (createLeaf, emitTreeKernel.cu:170-231)
__global__ void __launch_bounds__(256) leaf_emit_kernel(int n, const u64* __restrict__ ex, const uint* __restrict__ pack32,
                                                         const float* __restrict__ verts, const int* __restrict__ tris, const int* __restrict__ idx,
                                                         float4* __restrict__ woop, int* __restrict__ triIndex, float* __restrict__ triBoxOut,
                                                         int* __restrict__ keptGap)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const u64 e = ex[p];
    const uint2 pk = *reinterpret_cast<const uint2*>(pack32 + 2 * p);           // (gap p kept, a leaf starts at p)
    if (pk.x) keptGap[(uint)e] = p;                                            // kept gaps in rank order, for emit_kernel
    const int leafRank = (int)(e >> 32) + (int)pk.y - 1;                       // leaves started at or before p, minus one
    const int out = 3 * p + leafRank;
    const int t = __ldg(idx + p);
    float4 o0, o1, o2;
    const F3 va = ld3(verts, __ldg(tris + 3 * t)), vb = ld3(verts, __ldg(tris + 3 * t + 1)), vc = ld3(verts, __ldg(tris + 3 * t + 2));
    if (triBoxOut) tri_box_store(triBoxOut, p, va, vb, vc);      // plain LBVH: this is the only pass that gathers the geometry
    calc_woop(va, vb, vc, o0, o1, o2);
    woop[out] = o0; woop[out + 1] = o1; woop[out + 2] = o2;
    triIndex[out] = t; triIndex[out + 1] = 0; triIndex[out + 2] = 0;
    if (p == n - 1 || pack32[2 * (p + 1) + 1]) {
        const float z = __int_as_float((int)0x80000000);
        woop[out + 3] = make_float4(z, z, z, z);
        triIndex[out + 3] = 0;
    }
}

// N == 1: the reference still emits an inner root: split = (0+1)>>1 = 0, empty left leaf, right leaf = the triangle.
__global__ void single_triangle_kernel(const float* __restrict__ verts, const int* __restrict__ tris, const int* __restrict__ idx, float eps,
                                        int* __restrict__ nodes, float4* __restrict__ woop, int* __restrict__ triIndex)
{
    const float z = __int_as_float((int)0x80000000);
    float4 o0, o1, o2;
    calc_woop(ld3(verts, tris[0]), ld3(verts, tris[1]), ld3(verts, tris[2]), o0, o1, o2);
    woop[0] = make_float4(z, z, z, z); triIndex[0] = 0;
    woop[1] = o0; woop[2] = o1; woop[3] = o2; triIndex[1] = 0; triIndex[2] = 0; triIndex[3] = 0;
    woop[4] = make_float4(z, z, z, z); triIndex[4] = 0;
    (void)idx;                                               // idx[0] == 0 (written by morton_kernel)
    const F3 a = ld3(verts, tris[0]), b = ld3(verts, tris[1]), c3 = ld3(verts, tris[2]);
    const F3 mn = min3v(a, min3v(b, c3)), mx = max3v(a, max3v(b, c3));
    F3 lo, hi;
    lo.x = __fsub_rn(mn.x, eps); lo.y = __fsub_rn(mn.y, eps); lo.z = __fsub_rn(mn.z, eps);
    hi.x = __fadd_rn(mx.x, eps); hi.y = __fadd_rn(mx.y, eps); hi.z = __fadd_rn(mx.z, eps);
    float* nf = reinterpret_cast<float*>(nodes);
    F3 elo, ehi; elo.x = elo.y = elo.z = kF32Max; ehi.x = ehi.y = ehi.z = -kF32Max;
    store_child_box(nf, 0, elo, ehi);
    store_child_box(nf, 1, lo, hi);
    nodes[12] = ~0; nodes[13] = ~1; nodes[14] = -1; nodes[15] = 0;
}

constexpr int kOneSweepMaxKeys = 2400000;

struct Scratch {
    DevBuf keysB, idxB, hist, blockSums, nodeS, nodeE, parent, flags, pack, ex, counters;
    DevBuf childBox, childCost, flags2;      // SAH collapse
    DevBuf triBox;                           // 6 floats per sorted position
    DevBuf keptGap;                          // kept gaps (inner nodes) in rank order
    DevBuf zero;                             // scalars + scan descriptors + the one-sweep radix sort's zone: cleared by one memset per build
    // HLBVH
    DevBuf clsHead, clusterOf, clsStart, clsBox, clsTask0, clsTask1, clsBin, clsParent;
    DevBuf tBox0, tBox1, tCnt0, tCnt1, tId0, tId1, tFirst0, tFirst1, rInts, rBoxes, binBox, binCnt, blockSum, topNodes, topParent, topCounters;
};
Scratch g_scratch;
static_assert(sizeof(Scratch) % sizeof(DevBuf) == 0, "Scratch holds DevBuf members only");

} // namespace

// nt_shutdown: the grow-only scratch belongs to the device it was allocated on
void release_build_scratch()
{
    DevBuf* b = reinterpret_cast<DevBuf*>(&g_scratch);
    for (size_t i = 0; i < sizeof(Scratch) / sizeof(DevBuf); i++) b[i].release();
}

cudaError_t build_bvh_device(const float* dVerts, int numVerts, const int* dTris, int n,
                             const BuildParams& p, BuildOutput& out, cudaStream_t stream,
                             int numSMs, int* outLaunches, std::string* err, cudaEvent_t doneEvent)
{
    (void)numVerts;
    int launches = 0;
    cudaError_t e;
    const int linkMul = (p.layout == Layout_Compact2) ? 4 : 64;
#define NT_TRY(call) do { e = (call); if (e != cudaSuccess) { *outLaunches = launches; return e; } } while (0)

    bool hl = (p.builder != 0) && (p.hlbvhBits != 10);            // HLBVHBuilder.cpp:44-47
    if (hl && (p.hlbvhBits < 1 || p.hlbvhBits > 9)) {
        if (err) *err = "HLBVH: hlbvhBits must be in [1, 9] (10 selects plain LBVH)";
        return cudaErrorInvalidValue;
    }
    if ((long long)n * 4 >= 0x7fffffffLL) {
        if (err) *err = "scene too large for 32-bit Woop offsets";
        return cudaErrorInvalidValue;
    }

    Scratch& sc = g_scratch;
    NT_TRY(out.sortedKeys->reserve((size_t)n * 4));
    NT_TRY(out.sortedIdx->reserve((size_t)n * 4));
    uint* keysA = out.sortedKeys->as<uint>();
    int* idxA = out.sortedIdx->as<int>();

    // ---- everything the pipeline needs zeroed, in ONE allocation cleared by ONE memset: the scalars, the descriptors / tickets of
    // the two single-pass scans and the radix sort's zone (digit histograms, tickets, tile descriptors of the four passes)
    // One-sweep sort while all of a pass's tiles are resident at once (<= ~1200 tiles of 2048 keys): fewer launches, 0.19 -> 0.15 ms
    // at 283 K triangles.  Beyond that the serial per-digit look-back over thousands of tiles costs more than the histogram + scan
    // launches it replaces (10 M soup: 1.81 vs 1.74 ms), so large inputs keep the four-pass form.  NT_SORT_ONESWEEP=0/1 forces either.
    static const int oneSweepMode = [] { const char* e = getenv("NT_SORT_ONESWEEP"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    const bool oneSweep = oneSweepMode < 0 ? (n <= kOneSweepMaxKeys) : (oneSweepMode != 0);
    static const bool chainedScan = [] { const char* e = getenv("NT_SCAN_CHAINED"); return !e || atoi(e) != 0; }();
    const size_t zScalars = 0, zDescU = 64, zDescP = zDescU + ((scan_desc_bytes(n) + 15) & ~(size_t)15),
                 zSort = zDescP + ((scan_desc_bytes(n) + 15) & ~(size_t)15), zBytes = zSort + (oneSweep ? onesweep_zone_bytes(n, 4) : 0);
    NT_TRY(sc.zero.reserve(zBytes));
    NT_TRY(cudaMemsetAsync(sc.zero.p, 0, zBytes, stream));
    char* zone = sc.zero.as<char>();
    int* scalars = reinterpret_cast<int*>(zone + zScalars);
    int* rootGap = scalars;                              // int [0]
    u64* totals = reinterpret_cast<u64*>(scalars) + 1;   // bytes 8..15: (leaves << 32 | inner gap nodes)
    int* topScal = scalars + 4;                          // ints [4] numTasks, [5] top-level nodes written
    uint* clusterCount = reinterpret_cast<uint*>(scalars) + 6;   // int [6]
    uint* scanTickets = reinterpret_cast<uint*>(scalars) + 8;    // ints [8], [9]

    // ---- Morton codes (HLBVHBuilder.cpp:67-83: step = (hi - lo) / 1024 on the host, in fp32)
    NT_TRY(sc.pack.reserve((size_t)n * 8)); NT_TRY(sc.counters.reserve((size_t)n * 4));
    const float sx = (p.hi[0] - p.lo[0]) / 1024.0f, sy = (p.hi[1] - p.lo[1]) / 1024.0f, sz = (p.hi[2] - p.lo[2]) / 1024.0f;
    morton_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dVerts, dTris, n, p.lo[0], p.lo[1], p.lo[2], sx, sy, sz, keysA, idxA,
                                                        sc.pack.as<u64>(), sc.counters.as<int>());
    launches++;
    NT_TRY(cudaGetLastError());

    if (n == 1) {
        NT_TRY(out.nodes->reserve(64)); NT_TRY(out.woop->reserve(5 * 16)); NT_TRY(out.triIndex->reserve(5 * 4));
        single_triangle_kernel<<<1, 1, 0, stream>>>(dVerts, dTris, idxA, p.epsilon, out.nodes->as<int>(), out.woop->as<float4>(), out.triIndex->as<int>());
        launches++;
        NT_TRY(cudaGetLastError());
        if (doneEvent) NT_TRY(cudaEventRecord(doneEvent, stream));
        out.nodeBytes = 64; out.woopBytes = 80; out.idxBytes = 20;
        *outLaunches = launches;
        return cudaSuccess;
    }

    // ---- stable LSD radix sort, 4 x 8-bit digits over the 30-bit codes (four passes: sorted data ends in keysA / idxA)
    NT_TRY(sc.keysB.reserve((size_t)n * 4));
    NT_TRY(sc.idxB.reserve((size_t)n * 4));
    if (!oneSweep) NT_TRY(sc.hist.reserve(radix_hist_bytes(n)));
    NT_TRY(sc.blockSums.reserve(scan_block_sums_bytes((long long)radix_hist_bytes(n) / 4) + scan_block_sums_bytes(n)));
    NT_TRY(radix_sort_pairs<uint>(keysA, idxA, sc.keysB.as<uint>(), sc.idxB.as<int>(), n, 4, sc.hist.as<uint>(), sc.blockSums.as<uint>(), stream, &launches,
                                  oneSweep ? reinterpret_cast<uint*>(zone + zSort) : nullptr, true));

    // per-triangle boxes in sorted order.  HLBVH (cluster boxes) and the SAH collapse need them before the leaves are
    // numbered: one early gather pass; the plain LBVH gets them from leaf_emit_kernel, its only gather.
    NT_TRY(sc.triBox.reserve((size_t)n * 24));
    const bool earlyTriBox = hl || (p.collapse != 0);
    if (earlyTriBox) {
        tri_box_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, dVerts, dTris, idxA, sc.triBox.as<float>());
        launches++;
    }

    // ---- HLBVH: clusters + top-level SAH (one cooperative kernel)
    int hb = 30, C = 0;
    if (hl) {
        const int d = 3 * p.hlbvhBits;                    // low bits dropped: clusters are cells of the 3m-bit grid
        NT_TRY(sc.clsHead.reserve((size_t)n * 4)); NT_TRY(sc.clusterOf.reserve((size_t)n * 4));
        cluster_mark_kernel<<<(n + 255) / 256, 256, 0, stream>>>(keysA, n, d, sc.clsHead.as<uint>());
        launches++;
        if (chainedScan) NT_TRY(exclusive_scan_chained<uint>(sc.clsHead.as<uint>(), sc.clusterOf.as<uint>(), n, reinterpret_cast<uint*>(zone + zDescU), scanTickets, clusterCount, stream, &launches));
        else NT_TRY(exclusive_scan<uint>(sc.clsHead.as<uint>(), sc.clusterOf.as<uint>(), n, sc.blockSums.as<uint>(), clusterCount, stream, &launches));
        uint hc = 0;
        NT_TRY(cudaMemcpyAsync(&hc, clusterCount, 4, cudaMemcpyDeviceToHost, stream));     // readback 1 of 2: cluster count sizes the top-level buffers
        NT_TRY(cudaStreamSynchronize(stream));
        C = (int)hc;
        if (C < 2) hl = false;       // the whole scene sits in one grid cell: nothing for the SAH stage to do, plain LBVH
    }
    if (hl) {
        hb = 3 * p.hlbvhBits;
        const size_t maxTasks = (size_t)C / 2 + 2;
        NT_TRY(sc.clsStart.reserve(((size_t)C + 1) * 4)); NT_TRY(sc.clsBox.reserve((size_t)C * 24));
        NT_TRY(sc.clsTask0.reserve((size_t)C * 4)); NT_TRY(sc.clsTask1.reserve((size_t)C * 4));
        NT_TRY(sc.clsBin.reserve((size_t)C * 12)); NT_TRY(sc.clsParent.reserve((size_t)C * 4));
        NT_TRY(sc.tBox0.reserve(maxTasks * 24)); NT_TRY(sc.tBox1.reserve(maxTasks * 24));
        NT_TRY(sc.tCnt0.reserve(maxTasks * 4)); NT_TRY(sc.tCnt1.reserve(maxTasks * 4));
        NT_TRY(sc.tId0.reserve(maxTasks * 4)); NT_TRY(sc.tId1.reserve(maxTasks * 4));
        NT_TRY(sc.tFirst0.reserve(maxTasks * 4)); NT_TRY(sc.tFirst1.reserve(maxTasks * 4));
        NT_TRY(sc.rInts.reserve(maxTasks * 6 * 4)); NT_TRY(sc.rBoxes.reserve(maxTasks * 48));
        NT_TRY(sc.binBox.reserve(maxTasks * 3 * kBins * 24)); NT_TRY(sc.binCnt.reserve(maxTasks * 3 * kBins * 4));
        NT_TRY(sc.topNodes.reserve((size_t)C * 64)); NT_TRY(sc.topParent.reserve((size_t)C * 4)); NT_TRY(sc.topCounters.reserve((size_t)C * 4));
        NT_TRY(cudaMemsetAsync(sc.topNodes.p, 0, (size_t)C * 64, stream));
        NT_TRY(cudaMemsetAsync(sc.topCounters.p, 0, (size_t)C * 4, stream));
        NT_TRY(cudaMemsetAsync(sc.clsParent.p, 0, (size_t)C * 4, stream));

        cluster_start_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, sc.clsHead.as<uint>(), sc.clusterOf.as<uint>(), sc.clsStart.as<int>(), C);
        cluster_box_init_kernel<<<(C * 6 + 255) / 256, 256, 0, stream>>>(C, sc.clsBox.as<int>());
        cluster_box_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, sc.clusterOf.as<uint>(), sc.triBox.as<float>(), sc.clsBox.as<int>());
        launches += 3;
        NT_TRY(cudaGetLastError());

        // Shared-memory staging of the bins while a level has at most `smemTasks` tasks (measured on the bench tree, scripts/top_level_bench.py:
        // 16 tasks 0.932 ms, 32 0.890, 64 0.864, 128 0.855, 256 0.841 (172 KB, one CTA per SM anyway), 320 0.867).  NT_TOP_SMEM_TASKS overrides.
        static int topBlocksPerSM = 0, smemTasks = kSmemTasksDefault;
        if (!topBlocksPerSM) {
            if (const char* e = getenv("NT_TOP_SMEM_TASKS")) { const int v = atoi(e); if (v >= 1 && v <= 320) smemTasks = v; }
            const int smemBytes = smemTasks * 3 * kBins * 7 * 4;
            NT_TRY(cudaFuncSetAttribute(hlbvh_top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
            NT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&topBlocksPerSM, hlbvh_top_kernel, kTopThreads, smemBytes));
            if (topBlocksPerSM < 1) topBlocksPerSM = 1;
            int want = 1;                                     // fewer CTAs = cheaper grid barrier; the phases are latency bound
            if (const char* e = getenv("NT_TOP_CTAS")) want = atoi(e);
            if (want < 1) want = 1;
            if (topBlocksPerSM > want) topBlocksPerSM = want;
        }
        const int topSmemBytes = smemTasks * 3 * kBins * 7 * 4;
        const int grid = numSMs * topBlocksPerSM;
        NT_TRY(sc.blockSum.reserve((size_t)grid * 4));
        TopArgs ta;
        ta.C = C; ta.leafSize = p.leafSize; ta.linkMul = linkMul; ta.smemTasks = smemTasks;
        ta.clsStart = sc.clsStart.as<int>(); ta.clsBoxI = sc.clsBox.as<int>();
        ta.clsTask[0] = sc.clsTask0.as<int>(); ta.clsTask[1] = sc.clsTask1.as<int>();
        ta.clsBin = sc.clsBin.as<int>(); ta.clsParent = sc.clsParent.as<int>();
        ta.tBox[0] = sc.tBox0.as<float>(); ta.tBox[1] = sc.tBox1.as<float>();
        ta.tCnt[0] = sc.tCnt0.as<int>(); ta.tCnt[1] = sc.tCnt1.as<int>();
        ta.tId[0] = sc.tId0.as<int>(); ta.tId[1] = sc.tId1.as<int>();
        ta.tFirst[0] = sc.tFirst0.as<int>(); ta.tFirst[1] = sc.tFirst1.as<int>();
        int* ri = sc.rInts.as<int>();
        ta.rSplit = ri; ta.rAxis = ri + maxTasks; ta.rCntL = ri + 2 * maxTasks; ta.rCntR = ri + 3 * maxTasks;
        ta.rLocalOfs = ri + 4 * maxTasks; ta.rChild = ri + 5 * maxTasks;
        ta.rBoxes = sc.rBoxes.as<float>();
        ta.binBoxI = sc.binBox.as<int>(); ta.binCnt = sc.binCnt.as<int>();
        ta.blockSum = sc.blockSum.as<int>();
        ta.topNodes = sc.topNodes.as<int>(); ta.topParent = sc.topParent.as<int>();
        ta.scal = topScal;
        for (int k = 0; k < 3; k++) { ta.sceneLo[k] = p.lo[k]; ta.sceneHi[k] = p.hi[k]; }
        void* kargs[] = {&ta};
        NT_TRY(cudaLaunchCooperativeKernel((const void*)hlbvh_top_kernel, dim3(grid), dim3(kTopThreads), kargs, topSmemBytes, stream));
        launches++;
    }

    // ---- topology, forced leaves, numbering
    const int gaps = n - 1;
    NT_TRY(sc.nodeS.reserve((size_t)n * 4)); NT_TRY(sc.nodeE.reserve((size_t)n * 4)); NT_TRY(sc.parent.reserve((size_t)n * 4));
    NT_TRY(sc.flags.reserve((size_t)n * 4)); NT_TRY(sc.ex.reserve((size_t)n * 8)); NT_TRY(sc.keptGap.reserve((size_t)n * 4));
    const bool collapse = (p.collapse != 0);
    const int emitLeaf = collapse ? 1 : p.leafSize;          // SAH collapse starts from single-triangle leaves
    topology_kernel<<<(gaps + 255) / 256, 256, 0, stream>>>(keysA, n, emitLeaf, hb, hl ? sc.clusterOf.as<uint>() : nullptr,
                                                             hl ? sc.clsParent.as<int>() : nullptr, hl ? sc.clsStart.as<int>() : nullptr, p.leafSize,
                                                             sc.nodeS.as<int>(), sc.nodeE.as<int>(), sc.parent.as<int>(), sc.flags.as<uint>(), rootGap);
    finalize_kernel<<<(gaps + 255) / 256, 256, 0, stream>>>(n, hb, sc.nodeS.as<int>(), sc.parent.as<int>(), sc.flags.as<uint>(),
                                                             collapse ? nullptr : sc.pack.as<uint>());
    launches += 2;
    if (collapse) {
        NT_TRY(sc.childBox.reserve((size_t)n * 48)); NT_TRY(sc.childCost.reserve((size_t)n * 8)); NT_TRY(sc.flags2.reserve((size_t)n * 4));
        CollapseCtx cx;
        cx.nodeS = sc.nodeS.as<int>(); cx.nodeE = sc.nodeE.as<int>(); cx.parent = sc.parent.as<int>(); cx.flags = sc.flags.as<uint>();
        cx.childBox = sc.childBox.as<float>(); cx.childCost = sc.childCost.as<float>(); cx.counters = sc.counters.as<int>();
        cx.maxLeaf = (p.collapseMaxLeaf > 0) ? p.collapseMaxLeaf : p.leafSize;
        cx.triCost = p.collapseTriCost;
        collapse_analyse_kernel<<<(gaps + 255) / 256, 256, 0, stream>>>(n, cx, sc.triBox.as<float>(), p.epsilon);
        NT_TRY(cudaMemcpyAsync(sc.flags2.p, sc.flags.p, (size_t)gaps * 4, cudaMemcpyDeviceToDevice, stream));
        collapse_resolve_kernel<<<(gaps + 255) / 256, 256, 0, stream>>>(n, sc.parent.as<int>(), sc.flags2.as<uint>(), sc.flags.as<uint>());
        pack_kernel<<<(gaps + 255) / 256, 256, 0, stream>>>(n, sc.nodeS.as<int>(), sc.flags.as<uint>(), sc.pack.as<uint>());
        NT_TRY(cudaMemsetAsync(sc.counters.p, 0, (size_t)n * 4, stream));      // the refit in emit_kernel counts arrivals again
        launches += 3;
    }
    if (hl) {
        cluster_leaf_flag_kernel<<<(C + 255) / 256, 256, 0, stream>>>(C, p.leafSize, sc.clsStart.as<int>(), sc.pack.as<uint>());
        launches++;
    }
    NT_TRY(cudaGetLastError());
    if (chainedScan) NT_TRY(exclusive_scan_chained<u64>(sc.pack.as<u64>(), sc.ex.as<u64>(), n, reinterpret_cast<u64*>(zone + zDescP), scanTickets + 1, totals, stream, &launches));
    else NT_TRY(exclusive_scan<u64>(sc.pack.as<u64>(), sc.ex.as<u64>(), n, sc.blockSums.as<u64>(), totals, stream, &launches));

    // No host readback here: the output buffers are sized for the largest tree this input can produce (n - 1 gap nodes plus the
    // top-level nodes; one terminator per triangle), the emit kernels take the node / leaf counts from the device scalars, and
    // the counts come back with the one copy at the end of the pipeline.
    const size_t maxInner = (size_t)(n - 1) + (hl ? (size_t)C : 0), maxRows = (size_t)n * 4;
    NT_TRY(out.nodes->reserve(maxInner * 64));
    NT_TRY(out.woop->reserve(maxRows * 16));
    NT_TRY(out.triIndex->reserve(maxRows * 4));

    ClimbCtx cc;
    cc.nodes = out.nodes->as<int>(); cc.parent = sc.parent.as<int>(); cc.ex = sc.ex.as<u64>(); cc.gapCounters = sc.counters.as<int>();
    cc.topParent = hl ? sc.topParent.as<int>() : nullptr; cc.topCounters = hl ? sc.topCounters.as<int>() : nullptr;
    cc.nb.numTop = 0; cc.nb.rootGap = 0; cc.nb.rootRank = 0; cc.nb.linkMul = linkMul;
    if (hl) NT_TRY(cudaMemcpyAsync(out.nodes->p, sc.topNodes.p, (size_t)C * 64, cudaMemcpyDeviceToDevice, stream));   // >= the top-level nodes written (scalars[5])

    // Woop rows / indices / terminators.  Without an earlier tri_box pass (plain LBVH) this kernel is the one gather of the
    // geometry and also leaves the per-triangle boxes behind for emit_kernel, so it runs first.
    leaf_emit_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, sc.ex.as<u64>(), sc.pack.as<uint>(), dVerts, dTris, idxA,
                                                           out.woop->as<float4>(), out.triIndex->as<int>(), earlyTriBox ? nullptr : sc.triBox.as<float>(),
                                                           sc.keptGap.as<int>());
    emit_kernel<<<(gaps + 255) / 256, 256, 0, stream>>>(sc.keptGap.as<int>(), sc.nodeS.as<int>(), sc.nodeE.as<int>(), sc.flags.as<uint>(),
                                                         scalars, sc.triBox.as<float>(), p.epsilon, cc);
    launches += 2;
    if (hl) {
        cluster_leaf_emit_kernel<<<(C + 255) / 256, 256, 0, stream>>>(C, p.leafSize, sc.clsStart.as<int>(), sc.clsParent.as<int>(), scalars,
                                                                       sc.triBox.as<float>(), p.epsilon, cc);
        launches++;
    }
    NT_TRY(cudaGetLastError());
    if (doneEvent) NT_TRY(cudaEventRecord(doneEvent, stream));

    // the one host readback of an LBVH build (HLBVH: the second of two): node and leaf counts = the sizes of what was written
    int hs[8];
    NT_TRY(cudaMemcpyAsync(hs, scalars, 32, cudaMemcpyDeviceToHost, stream));
    NT_TRY(cudaStreamSynchronize(stream));
    *outLaunches = launches;
    u64 tot; memcpy(&tot, &hs[2], 8);
    const size_t numTop = hl ? (size_t)hs[5] : 0;
    const size_t numInner = (size_t)(tot & 0xffffffffull) + numTop, numLeaves = (size_t)(tot >> 32);
    if (numInner == 0 || numLeaves == 0) { if (err) *err = "internal error: empty tree"; return cudaErrorUnknown; }
    // links must stay below the EntrypointSentinel 0x76543210 the traversal stack uses (CudaTracerKernels.hpp:36-39)
    if (numInner * (size_t)linkMul >= 0x76543210ull) {
        if (err) *err = p.layout == Layout_Compact2 ? "node buffer exceeds the 32-bit offset range of BVHLayout_Compact2"
                                                    : "node buffer exceeds the 32-bit byte-offset range of BVHLayout_Compact (build into BVHLayout_Compact2: nt_bvh_set_build_layout)";
        return cudaErrorInvalidValue;
    }
    out.nodeBytes = numInner * 64;
    out.woopBytes = ((size_t)n * 3 + numLeaves) * 16;
    out.idxBytes = ((size_t)n * 3 + numLeaves) * 4;
    return cudaSuccess;
#undef NT_TRY
}

} // namespace nt
