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
//     node (atomicAdd + __threadfence), writing child boxes straight into the parent's node words.
// One host readback (node / leaf counts, to size the output buffers) instead of one per level.
#include "nt_common.cuh"

namespace nt {

namespace {

typedef unsigned int uint;
typedef unsigned long long u64;

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
                                                      uint* __restrict__ keys, int* __restrict__ idx)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
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
// Exclusive scan (reduce / scan block sums / apply), T = uint or u64.  Hand-written, no CUB/thrust.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <class T>
__device__ __forceinline__ T block_exclusive(T v, T* s_warp, T& total)
{
    // exclusive scan of one value per thread across a 256-thread block
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = (lane < kScanThreads / 32) ? s_warp[lane] : T(0);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { T y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane < kScanThreads / 32) s_warp[lane] = x;          // inclusive over warps
    }
    __syncthreads();
    const T warpBase = (w == 0) ? T(0) : s_warp[w - 1];
    total = s_warp[kScanThreads / 32 - 1];
    __syncthreads();
    return warpBase + inc - v;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const T* __restrict__ in, long long n, T* __restrict__ blockSums)
{
    __shared__ T s_warp[kScanThreads / 32];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    T sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) if (base + i < n) sum += in[base + i];
    T total;
    block_exclusive<T>(sum, s_warp, total);
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(T* __restrict__ blockSums, int numBlocks, T* __restrict__ grandTotal)
{
    __shared__ T s_warp[kScanThreads / 32];
    T carry = 0;
    for (int base = 0; base < numBlocks; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const T v = (i < numBlocks) ? blockSums[i] : T(0);
        T total;
        const T ex = block_exclusive<T>(v, s_warp, total);
        if (i < numBlocks) blockSums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && grandTotal) *grandTotal = carry;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const T* __restrict__ in, T* __restrict__ out, long long n, const T* __restrict__ blockSums)
{
    __shared__ T s_warp[kScanThreads / 32];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    T v[kScanItems];
    T sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) { v[i] = (base + i < n) ? in[base + i] : T(0); sum += v[i]; }
    T total;
    T run = block_exclusive<T>(sum, s_warp, total) + blockSums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; i++) { if (base + i < n) out[base + i] = run; run += v[i]; }
}

template <class T>
cudaError_t exclusive_scan(const T* in, T* out, long long n, T* blockSums /* >= ceil(n/tile) */, T* grandTotal, cudaStream_t s, int* launches)
{
    const int nb = (int)((n + kScanTile - 1) / kScanTile);
    scan_reduce_kernel<T><<<nb, kScanThreads, 0, s>>>(in, n, blockSums);
    scan_sums_kernel<T><<<1, kScanThreads, 0, s>>>(blockSums, nb, grandTotal);
    scan_apply_kernel<T><<<nb, kScanThreads, 0, s>>>(in, out, n, blockSums);
    *launches += 3;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort of (key, index) pairs, 8-bit digits.
// ------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;                       // rounds of 32 consecutive keys per warp
constexpr int kSortTile = kSortThreads * kSortItems;

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint* __restrict__ keys, int n, int shift, uint* __restrict__ hist, int numBlocks)
{
    __shared__ uint s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortItems; r++) {
        const int i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&s_hist[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * numBlocks + blockIdx.x] = s_hist[threadIdx.x];     // digit-major: one scan gives global offsets
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint* __restrict__ keysIn, const int* __restrict__ idxIn,
                                                                     uint* __restrict__ keysOut, int* __restrict__ idxOut,
                                                                     int n, int shift, const uint* __restrict__ histScan, int numBlocks)
{
    __shared__ uint s_cnt[kSortThreads / 32][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();

    const int segBase = blockIdx.x * kSortTile + w * (kSortItems * 32);
    uint key[kSortItems];
    uint rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; r++) {
        const int i = segBase + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keysIn[i] : 0xffffffffu;
        const uint digit = valid ? ((key[r] >> shift) & 255u) : 256u;
        const uint peers = __match_any_sync(0xffffffffu, digit);
        uint pre = 0;
        if (valid) pre = s_cnt[w][digit];
        __syncwarp();
        if (valid && lane == (31 - __clz(peers))) s_cnt[w][digit] = pre + __popc(peers);
        __syncwarp();
        rank[r] = pre + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    {
        // digit = threadIdx.x: exclusive prefix over the warps of this tile + global base of (digit, tile)
        uint run = histScan[threadIdx.x * numBlocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < kSortThreads / 32; ww++) { const uint c = s_cnt[ww][threadIdx.x]; s_cnt[ww][threadIdx.x] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; r++) {
        const int i = segBase + r * 32 + lane;
        if (i < n) {
            const uint pos = s_cnt[w][(key[r] >> shift) & 255u] + rank[r];
            keysOut[pos] = key[r];
            idxOut[pos] = idxIn[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Topology: one thread per gap g in [0, N-2] (between sorted positions g and g+1).
// ------------------------------------------------------------------------------------------------
enum : uint { F_KEPT = 1u, F_LEFT_LEAF = 2u, F_RIGHT_LEAF = 4u, F_NEED_DEPTH = 8u };

__device__ __forceinline__ int topbit(uint x) { return 31 - __clz(x); }

// smallest f <= from such that pred(K[f]) holds on [f, from]; pred is monotone (true near `from`)
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

__global__ void __launch_bounds__(256) topology_kernel(const uint* __restrict__ K, int n, int leafSize,
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
    if (par == -2) {
        // bounded by radix gaps: the one with the lower split bit is the parent
        if (s == 0 && e == n) { par = -1; *rootGap = g; }
        else if (s == 0) par = (e - 1) * 2 + 0;
        else if (e == n) par = (s - 1) * 2 + 1;
        else {
            const int bl = topbit(__ldg(K + s - 1) ^ __ldg(K + s)), br = topbit(__ldg(K + e - 1) ^ __ldg(K + e));
            par = (bl < br) ? (s - 1) * 2 + 1 : (e - 1) * 2 + 0;
        }
    }
    const int split = g + 1;
    const bool root = (par == -1);
    uint f = 0;
    if ((e - s) > leafSize || root) {
        f = F_KEPT;
        if (split - s <= leafSize) f |= F_LEFT_LEAF;
        if (e - split <= leafSize) f |= F_RIGHT_LEAF;
        // A radix node splitting bit b sits at depth <= 29 - b, so only two kinds of node can be affected by the
        // 29-level rule: duplicate-run nodes (b == -1; they may not exist at all if they are >= 30 levels down)
        // and bit-0 nodes that would otherwise keep an inner child.
        const bool bothLeaves = (f & (F_LEFT_LEAF | F_RIGHT_LEAF)) == (F_LEFT_LEAF | F_RIGHT_LEAF);
        if (b < 0 || (b == 0 && !bothLeaves)) f |= F_NEED_DEPTH;
    }
    f |= (uint)(b + 1) << 8;
    nodeS[g] = s; nodeE[g] = e; parent[g] = par; flags[g] = f;
}

// forced leaves 29 levels below the root (emitTreeKernel.cu:289-292 `oldLevel == 0`), then scan inputs:
// pack[i].lo = gap i is an inner node, pack[i].hi = a leaf starts at sorted position i
__global__ void __launch_bounds__(256) finalize_kernel(int n, const int* __restrict__ nodeS, const int* __restrict__ parent,
                                                        uint* __restrict__ flags, uint* __restrict__ pack32)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n - 1) return;
    uint f = flags[g];
    if (f & F_NEED_DEPTH) {
        int depth = 0;
        for (int p = parent[g]; p >= 0; p = parent[p >> 1]) depth++;
        if (depth >= 30) f &= ~(F_KEPT | F_LEFT_LEAF | F_RIGHT_LEAF);
        else if (depth == 29) f |= F_LEFT_LEAF | F_RIGHT_LEAF;
        flags[g] = f;
    }
    if (f & F_KEPT) {
        pack32[2 * g] = 1u;
        if (f & F_LEFT_LEAF) pack32[2 * nodeS[g] + 1] = 1u;
        if (f & F_RIGHT_LEAF) pack32[2 * (g + 1) + 1] = 1u;
    }
}

__device__ __forceinline__ int node_id(uint rank, uint rootRank, bool isRoot) { return isRoot ? 0 : (int)(rank < rootRank ? rank + 1 : rank); }

// leaf box = union of triangle vertices -/+ epsilon (calcLeaf, emitTreeKernel.cu:383-408)
__device__ __forceinline__ void leaf_box(const float* __restrict__ verts, const int* __restrict__ tris, const int* __restrict__ idx,
                                         int a, int b, float eps, F3& lo, F3& hi)
{
    lo.x = lo.y = lo.z = kF32Max; hi.x = hi.y = hi.z = -kF32Max;
    for (int i = a; i < b; i++) {
        const int t = __ldg(idx + i);
        const F3 p = ld3(verts, __ldg(tris + 3 * t)), q = ld3(verts, __ldg(tris + 3 * t + 1)), r = ld3(verts, __ldg(tris + 3 * t + 2));
        const F3 mn = min3v(p, min3v(q, r)), mx = max3v(p, max3v(q, r));
        lo.x = fminf(lo.x, __fsub_rn(mn.x, eps)); lo.y = fminf(lo.y, __fsub_rn(mn.y, eps)); lo.z = fminf(lo.z, __fsub_rn(mn.z, eps));
        hi.x = fmaxf(hi.x, __fadd_rn(mx.x, eps)); hi.y = fmaxf(hi.y, __fadd_rn(mx.y, eps)); hi.z = fmaxf(hi.z, __fadd_rn(mx.z, eps));
    }
}

__device__ __forceinline__ void store_child_box(float* node, int side, F3 lo, F3 hi)
{
    // node words: c0 -> 0..3, 8, 9 ; c1 -> 4..7, 10, 11  (CudaBVH.hpp:43-47)
    *reinterpret_cast<float4*>(node + 4 * side) = make_float4(lo.x, hi.x, lo.y, hi.y);
    *reinterpret_cast<float2*>(node + 8 + 2 * side) = make_float2(lo.z, hi.z);
}

__global__ void __launch_bounds__(256) emit_kernel(int n, const int* __restrict__ nodeS, const int* __restrict__ nodeE,
                                                    const int* __restrict__ parent, const uint* __restrict__ flags,
                                                    const u64* __restrict__ ex, const int* __restrict__ rootGapPtr,
                                                    const float* __restrict__ verts, const int* __restrict__ tris, const int* __restrict__ idx,
                                                    float eps, int* __restrict__ nodes, int* __restrict__ counters)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n - 1) return;
    const uint f = flags[g];
    if (!(f & F_KEPT)) return;
    const int rootGap = *rootGapPtr;
    const uint rootRank = (uint)ex[rootGap];
    const int id = node_id((uint)ex[g], rootRank, g == rootGap);
    const int s = nodeS[g], e = nodeE[g], split = g + 1;
    int* node = nodes + (size_t)id * 16;
    float* nodef = reinterpret_cast<float*>(node);

    const int b = (int)((f >> 8) & 0xffu) - 1;
    node[14] = (b < 0) ? -1 : (b % 3);           // `level % 3` with C remainder semantics (emitTreeKernel.cu:378)
    node[15] = 0;
    const int par = parent[g];
    if (par >= 0) {
        const int pg = par >> 1;
        const int pid = node_id((uint)ex[pg], rootRank, pg == rootGap);
        nodes[(size_t)pid * 16 + 12 + (par & 1)] = id * 64;        // byte offset (Compact)
    }

    int arrivals = 0;
    if (f & F_LEFT_LEAF) {
        node[12] = ~(3 * s + (int)(ex[s] >> 32));
        F3 lo, hi; leaf_box(verts, tris, idx, s, split, eps, lo, hi);
        store_child_box(nodef, 0, lo, hi);
        arrivals++;
    }
    if (f & F_RIGHT_LEAF) {
        node[13] = ~(3 * split + (int)(ex[split] >> 32));
        F3 lo, hi; leaf_box(verts, tris, idx, split, e, eps, lo, hi);
        store_child_box(nodef, 1, lo, hi);
        arrivals++;
    }
    if (arrivals == 0) return;
    if (arrivals == 1) {
        __threadfence();
        if (atomicAdd(counters + g, 1) == 0) return;          // the inner child has not arrived yet
    }

    // bottom-up refit: this node is complete; carry its box to the parent until we are first somewhere
    int cur = g, curId = id;
    for (;;) {
        const int p = parent[cur];
        if (p < 0) break;
        const float* c = reinterpret_cast<const float*>(nodes + (size_t)curId * 16);
        const float4 b0 = __ldcg(reinterpret_cast<const float4*>(c));
        const float4 b1 = __ldcg(reinterpret_cast<const float4*>(c + 4));
        const float4 bz = __ldcg(reinterpret_cast<const float4*>(c + 8));
        F3 lo, hi;
        lo.x = fminf(b0.x, b1.x); hi.x = fmaxf(b0.y, b1.y);
        lo.y = fminf(b0.z, b1.z); hi.y = fmaxf(b0.w, b1.w);
        lo.z = fminf(bz.x, bz.z); hi.z = fmaxf(bz.y, bz.w);
        const int pg = p >> 1;
        const int pid = node_id((uint)ex[pg], rootRank, pg == rootGap);
        store_child_box(reinterpret_cast<float*>(nodes + (size_t)pid * 16), p & 1, lo, hi);
        __threadfence();
        if (atomicAdd(counters + pg, 1) == 0) break;
        cur = pg; curId = pid;
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
                                                         float4* __restrict__ woop, int* __restrict__ triIndex)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int leafRank = (int)(ex[p] >> 32) + (int)pack32[2 * p + 1] - 1;      // leaves started at or before p, minus one
    const int out = 3 * p + leafRank;
    const int t = __ldg(idx + p);
    float4 o0, o1, o2;
    calc_woop(ld3(verts, __ldg(tris + 3 * t)), ld3(verts, __ldg(tris + 3 * t + 1)), ld3(verts, __ldg(tris + 3 * t + 2)), o0, o1, o2);
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
    F3 lo, hi;
    leaf_box(verts, tris, idx, 0, 1, eps, lo, hi);          // idx[0] == 0 (written by morton_kernel)
    float* nf = reinterpret_cast<float*>(nodes);
    F3 elo, ehi; elo.x = elo.y = elo.z = kF32Max; ehi.x = ehi.y = ehi.z = -kF32Max;
    store_child_box(nf, 0, elo, ehi);
    store_child_box(nf, 1, lo, hi);
    nodes[12] = ~0; nodes[13] = ~1; nodes[14] = -1; nodes[15] = 0;
}

struct Scratch {
    DevBuf keysB, idxB, hist, blockSums, nodeS, nodeE, parent, flags, pack, ex, counters, scalars;
};
Scratch g_scratch;

} // namespace

cudaError_t build_bvh_device(const float* dVerts, int numVerts, const int* dTris, int n,
                             const BuildParams& p, BuildOutput& out, cudaStream_t stream,
                             int numSMs, int* outLaunches, std::string* err)
{
    (void)numVerts; (void)numSMs;
    int launches = 0;
    cudaError_t e;
#define NT_TRY(call) do { e = (call); if (e != cudaSuccess) { *outLaunches = launches; return e; } } while (0)

    if (p.builder != 0 && !(p.hlbvhBits == 10)) {
        if (err) *err = "HLBVH top level (hlbvhBits != 10) is not implemented in this build; use NT_BUILDER_LBVH";
        return cudaErrorNotSupported;
    }
    if ((long long)n * 3 + n >= 0x7fffffffLL / 1) {
        if (err) *err = "scene too large for 32-bit Woop offsets";
        return cudaErrorInvalidValue;
    }

    Scratch& sc = g_scratch;
    NT_TRY(out.sortedKeys->reserve((size_t)n * 4));
    NT_TRY(out.sortedIdx->reserve((size_t)n * 4));
    uint* keysA = out.sortedKeys->as<uint>();
    int* idxA = out.sortedIdx->as<int>();

    // ---- Morton codes (HLBVHBuilder.cpp:67-83: step = (hi - lo) / 1024 on the host, in fp32)
    const float sx = (p.hi[0] - p.lo[0]) / 1024.0f, sy = (p.hi[1] - p.lo[1]) / 1024.0f, sz = (p.hi[2] - p.lo[2]) / 1024.0f;
    morton_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dVerts, dTris, n, p.lo[0], p.lo[1], p.lo[2], sx, sy, sz, keysA, idxA);
    launches++;
    NT_TRY(cudaGetLastError());

    if (n == 1) {
        NT_TRY(out.nodes->reserve(64)); NT_TRY(out.woop->reserve(5 * 16)); NT_TRY(out.triIndex->reserve(5 * 4));
        single_triangle_kernel<<<1, 1, 0, stream>>>(dVerts, dTris, idxA, p.epsilon, out.nodes->as<int>(), out.woop->as<float4>(), out.triIndex->as<int>());
        launches++;
        NT_TRY(cudaGetLastError());
        out.nodeBytes = 64; out.woopBytes = 80; out.idxBytes = 20;
        *outLaunches = launches;
        return cudaSuccess;
    }

    // ---- stable LSD radix sort, 4 x 8-bit digits over the 30-bit codes
    {
        const int nb = (n + kSortTile - 1) / kSortTile;
        NT_TRY(sc.keysB.reserve((size_t)n * 4));
        NT_TRY(sc.idxB.reserve((size_t)n * 4));
        NT_TRY(sc.hist.reserve((size_t)nb * 256 * 4));
        const long long histLen = (long long)nb * 256;
        NT_TRY(sc.blockSums.reserve(((size_t)(histLen + kScanTile - 1) / kScanTile + (size_t)(n + kScanTile - 1) / kScanTile + 16) * 8));
        uint* kin = keysA; int* iin = idxA;
        uint* kout = sc.keysB.as<uint>(); int* iout = sc.idxB.as<int>();
        for (int pass = 0; pass < 4; pass++) {
            const int shift = pass * 8;
            radix_hist_kernel<<<nb, kSortThreads, 0, stream>>>(kin, n, shift, sc.hist.as<uint>(), nb);
            launches++;
            NT_TRY(exclusive_scan<uint>(sc.hist.as<uint>(), sc.hist.as<uint>(), histLen, sc.blockSums.as<uint>(), nullptr, stream, &launches));
            radix_scatter_kernel<<<nb, kSortThreads, 0, stream>>>(kin, iin, kout, iout, n, shift, sc.hist.as<uint>(), nb);
            launches++;
            NT_TRY(cudaGetLastError());
            uint* tk = kin; kin = kout; kout = tk;
            int* ti = iin; iin = iout; iout = ti;
        }
        // four passes: the sorted data is back in keysA / idxA
    }

    // ---- topology, forced leaves, numbering
    const int gaps = n - 1;
    NT_TRY(sc.nodeS.reserve((size_t)n * 4)); NT_TRY(sc.nodeE.reserve((size_t)n * 4)); NT_TRY(sc.parent.reserve((size_t)n * 4));
    NT_TRY(sc.flags.reserve((size_t)n * 4)); NT_TRY(sc.pack.reserve((size_t)n * 8)); NT_TRY(sc.ex.reserve((size_t)n * 8));
    NT_TRY(sc.counters.reserve((size_t)n * 4)); NT_TRY(sc.scalars.reserve(64));
    NT_TRY(cudaMemsetAsync(sc.pack.p, 0, (size_t)n * 8, stream));
    NT_TRY(cudaMemsetAsync(sc.counters.p, 0, (size_t)n * 4, stream));
    NT_TRY(cudaMemsetAsync(sc.scalars.p, 0, 64, stream));
    int* rootGap = sc.scalars.as<int>();                 // [0]
    u64* totals = sc.scalars.as<u64>() + 1;              // bytes 8..15: (leaves << 32 | inner nodes)
    topology_kernel<<<(gaps + 255) / 256, 256, 0, stream>>>(keysA, n, p.leafSize, sc.nodeS.as<int>(), sc.nodeE.as<int>(), sc.parent.as<int>(),
                                                             sc.flags.as<uint>(), rootGap);
    finalize_kernel<<<(gaps + 255) / 256, 256, 0, stream>>>(n, sc.nodeS.as<int>(), sc.parent.as<int>(), sc.flags.as<uint>(), sc.pack.as<uint>());
    launches += 2;
    NT_TRY(cudaGetLastError());
    NT_TRY(exclusive_scan<u64>(sc.pack.as<u64>(), sc.ex.as<u64>(), n, sc.blockSums.as<u64>(), totals, stream, &launches));

    // the only host readback of the build: inner-node and leaf counts size the output buffers
    u64 tot = 0;
    NT_TRY(cudaMemcpyAsync(&tot, totals, 8, cudaMemcpyDeviceToHost, stream));
    NT_TRY(cudaStreamSynchronize(stream));
    const size_t numInner = (size_t)(tot & 0xffffffffull), numLeaves = (size_t)(tot >> 32);
    if (numInner == 0 || numLeaves == 0) { if (err) *err = "internal error: empty tree"; return cudaErrorUnknown; }
    if (numInner * 64 >= 0x76543210ull) { if (err) *err = "node buffer exceeds the 32-bit byte-offset range of BVHLayout_Compact"; return cudaErrorInvalidValue; }
    out.nodeBytes = numInner * 64;
    out.woopBytes = ((size_t)n * 3 + numLeaves) * 16;
    out.idxBytes = ((size_t)n * 3 + numLeaves) * 4;
    NT_TRY(out.nodes->reserve(out.nodeBytes));
    NT_TRY(out.woop->reserve(out.woopBytes));
    NT_TRY(out.triIndex->reserve(out.idxBytes));

    emit_kernel<<<(gaps + 255) / 256, 256, 0, stream>>>(n, sc.nodeS.as<int>(), sc.nodeE.as<int>(), sc.parent.as<int>(), sc.flags.as<uint>(),
                                                         sc.ex.as<u64>(), rootGap, dVerts, dTris, idxA, p.epsilon,
                                                         out.nodes->as<int>(), sc.counters.as<int>());
    leaf_emit_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, sc.ex.as<u64>(), sc.pack.as<uint>(), dVerts, dTris, idxA,
                                                           out.woop->as<float4>(), out.triIndex->as<int>());
    launches += 2;
    NT_TRY(cudaGetLastError());
    *outLaunches = launches;
    return cudaSuccess;
#undef NT_TRY
}

} // namespace nt
