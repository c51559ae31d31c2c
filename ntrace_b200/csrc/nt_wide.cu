// ntrace_b200 — 4-wide, 8-bit quantised BVH ("Wide4") derived from the resident binary CudaBVH, and its traversal kernel.
//
// Why (profiles/r1_summary.md): on a B200 the binary kernel (nt_trace.cu) is bound by the L1 data
// pipe (67-75 % busy: two 32-byte sectors per node visit and lane), by dependent L1/L2 round trips (long-scoreboard stalls) and
// by the ALU pipe (20 FMNMX per node), not by HBM.  A wide node attacks the first two: one 64-byte Wide4 node replaces a binary
// node AND its two children (~0.55x the node visits for the same 64 bytes per visit), and knowing the ray's direction signs picks
// the near / far plane of every child box without a min/max pair.
//
// The API does not change: the host still hands over / builds the reference's Compact (or Compact2) CudaBVH
// (src/rt/cuda/CudaBVH.hpp:43-56).  The Wide4 node array is derived from it when a "b200_wide4*" kernel is selected, exactly like
// nt_layout.cu derives the Compact form from AOS/SOA uploads.  Leaves are NOT rewritten: a Wide4 child link < 0 is the same
// ~woopIndex as in the binary tree, so the Woop triangles, the terminators and the triIndex remap stay the reference's
// (CudaBVH.cpp:620-638) and the triangle test is the one of nt_trace.cu — hit t/u/v are bit-identical per (ray, triangle).
//
// Wide4 node, 64 bytes (16 words), node i at byte i * 64:
//   w0..2   p        float3  origin of the quantisation grid = lo corner of the union of the child boxes
//   w3..5   s'       float3  grid step * 2^15 per axis (see decode)
//   w6,w7   qlo.x, qlo.y      4 x uint8 each: child i in byte i
//   w8      qlo.z    w9..11  qhi.x, qhi.y, qhi.z
//   w12..15 link[4]  int     >= 0: Wide4 node index;  < 0: ~woopIndex of a leaf (as in the Compact layout);
//                            an unused slot has an inverted box (qlo 255, qhi 0) and links to the last terminator of the Woop array
// Decode: plane = p + q * s.  In the kernel a quantised plane enters the slab test as
//   t = fma(as_float(0x3F800000 | q << 8), s' * idir, (p * idir - ood) - s' * idir)       [as_float(..) = 1 + q * 2^-15, s' = s * 2^15]
// i.e. one PRMT and one FFMA per plane.  Boxes are quantised outwards with a margin of 1/8 grid step, which covers the rounding of
// that expression, so a child box always contains the binary tree's box: traversal visits a superset of what the reference visits.
#include "nt_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace nt {

// ================================================================================================================
// host: Compact / Compact2 -> Wide4
// ================================================================================================================
namespace {

struct Cand { float lo[3], hi[3]; int link; };

inline float cand_area(const Cand& c)
{
    const float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// children of the binary node at link `addr` (Compact: byte offset, Compact2: float4 index); CudaBVH.cpp:646-651
inline bool read_binary_children(const int32_t* nodes, size_t nodeBytes, int layout, int addr, Cand out[2])
{
    const size_t byteOfs = (layout == Layout_Compact2) ? (size_t)(uint32_t)addr * 16 : (size_t)(uint32_t)addr;
    if (byteOfs % 64 || byteOfs + 64 > nodeBytes) return false;
    const int32_t* w = nodes + byteOfs / 4;
    const float* f = reinterpret_cast<const float*>(w);
    out[0].lo[0] = f[0]; out[0].hi[0] = f[1]; out[0].lo[1] = f[2]; out[0].hi[1] = f[3];
    out[1].lo[0] = f[4]; out[1].hi[0] = f[5]; out[1].lo[1] = f[6]; out[1].hi[1] = f[7];
    out[0].lo[2] = f[8]; out[0].hi[2] = f[9]; out[1].lo[2] = f[10]; out[1].hi[2] = f[11];
    out[0].link = w[12]; out[1].link = w[13];
    return true;
}

// one axis of a Wide4 node: grid origin, step and the outward-rounded plane bytes of n children
void quantise_axis(const Cand* c, int n, int axis, float& p, float& sPrime, uint8_t qlo[4], uint8_t qhi[4])
{
    float mn = c[0].lo[axis], mx = c[0].hi[axis];
    for (int i = 1; i < n; i++) { mn = std::min(mn, c[i].lo[axis]); mx = std::max(mx, c[i].hi[axis]); }
    p = mn;
    const double ext = (double)mx - (double)mn;
    float s = (float)(ext / 253.0);                       // 253 steps for the extent: room for the two 1/8-step margins and the rounding below
    if (!(s > 1.0e-30f)) s = 1.0e-30f;                    // flat or degenerate axis (also NaN / negative extents)
    while ((double)s * 253.0 < ext) s = std::nextafterf(s, INFINITY);
    for (int i = 0; i < 4; i++) {
        if (i >= n) { qlo[i] = 255; qhi[i] = 0; continue; }               // unused slot: inverted box
        int a = (int)std::floor(((double)c[i].lo[axis] - (double)p) / (double)s - 0.125);
        a = std::min(std::max(a, 0), 255);
        while (a > 0 && (double)p + (double)a * (double)s > (double)c[i].lo[axis]) a--;
        int b = (int)std::ceil(((double)c[i].hi[axis] - (double)p) / (double)s + 0.125);
        b = std::min(std::max(b, 0), 255);
        while (b < 255 && (double)p + (double)b * (double)s < (double)c[i].hi[axis]) b++;
        qlo[i] = (uint8_t)a; qhi[i] = (uint8_t)b;
    }
    sPrime = s * 32768.0f;
}

} // namespace

int convert_compact_to_wide4_host(const int32_t* nodes, size_t nodeBytes, int layout, size_t woopRows,
                                  std::vector<uint32_t>& out, int* outMaxDepth, std::string* err)
{
    auto fail = [&](const char* m) { if (err) *err = m; return 1; };
    if (layout != Layout_Compact && layout != Layout_Compact2) return fail("Wide4 is derived from BVHLayout_Compact / Compact2");
    if (!nodes || nodeBytes < 64 || nodeBytes % 64 || woopRows < 1) return fail("malformed BVH: empty node or triangle buffer");
    const size_t numBinary = nodeBytes / 64;
    const int emptyLink = ~(int)(woopRows - 1);           // the Woop array ends with the terminator of its last leaf
    out.clear();
    out.reserve((numBinary / 2 + 8) * 16);
    struct Work { int wide; int binary; int depth; };
    std::vector<Work> work;
    out.resize(16);
    work.push_back({0, 0, 1});
    size_t consumed = 0;                                  // binary nodes folded into wide nodes (cycle guard)
    int maxDepth = 0;
    while (!work.empty()) {
        const Work wk = work.back();
        work.pop_back();
        maxDepth = std::max(maxDepth, wk.depth);
        Cand c[4];
        int n = 2;
        if (!read_binary_children(nodes, nodeBytes, layout, wk.binary, c)) return fail("malformed BVH: child link outside the node buffer");
        if (++consumed > numBinary) return fail("malformed BVH: the node links form a cycle");
        // open the inner child with the largest surface area until four children are held (or only leaves are left)
        while (n < 4) {
            int best = -1;
            float bestArea = -1.0f;
            for (int i = 0; i < n; i++)
                if (c[i].link >= 0) { const float a = cand_area(c[i]); if (a > bestArea || best < 0) { bestArea = a; best = i; } }
            if (best < 0) break;
            Cand two[2];
            if (!read_binary_children(nodes, nodeBytes, layout, c[best].link, two)) return fail("malformed BVH: child link outside the node buffer");
            if (++consumed > numBinary) return fail("malformed BVH: the node links form a cycle");
            c[best] = two[0];
            c[n++] = two[1];
        }
        uint32_t w[16];
        float p[3], sp[3];
        uint8_t qlo[3][4], qhi[3][4];
        for (int a = 0; a < 3; a++) quantise_axis(c, n, a, p[a], sp[a], qlo[a], qhi[a]);
        auto pack = [](const uint8_t b[4]) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); };
        for (int a = 0; a < 3; a++) { std::memcpy(&w[a], &p[a], 4); std::memcpy(&w[3 + a], &sp[a], 4); }
        w[6] = pack(qlo[0]); w[7] = pack(qlo[1]); w[8] = pack(qlo[2]);
        w[9] = pack(qhi[0]); w[10] = pack(qhi[1]); w[11] = pack(qhi[2]);
        for (int i = 0; i < 4; i++) {
            int link = emptyLink;
            if (i < n) {
                if (c[i].link < 0) {
                    if ((size_t)(uint32_t)~c[i].link >= woopRows) return fail("malformed BVH: leaf link outside the triangle buffer");
                    link = c[i].link;
                } else {
                    const size_t idx = out.size() / 16;
                    if (idx >= (size_t)kEntrypointSentinel) return fail("BVH too large for the Wide4 form");
                    link = (int)idx;
                    out.resize(out.size() + 16);
                    work.push_back({link, c[i].link, wk.depth + 1});
                }
            }
            w[12 + i] = (uint32_t)link;
        }
        std::memcpy(&out[(size_t)wk.wide * 16], w, 64);
    }
    if (outMaxDepth) *outMaxDepth = maxDepth;
    if (maxDepth > kWideMaxDepth) return fail("BVH too deep for the Wide4 kernel's traversal stack (use a binary kernel)");
    return 0;
}

// ================================================================================================================
// device: Compact / Compact2 -> Wide4 (what the library itself uses; the host routine above stays as the reference statement of the
// format for the oracle's emulation and for tests).  One persistent kernel; a thread takes wide-node indices from a ticket counter and
// waits until the node's parent has published which binary node it folds (the root is published by the launcher).  Parents always have
// smaller indices than their children and tickets are handed out in index order, so every wait is for a thread that already runs.
// Each node is built with the host routine's arithmetic (explicitly rounded double operations: same plane bytes, bit for bit) and the
// same child-slot order; only the numbering of the nodes differs (slots are handed out by an atomic counter), which no result depends on.
// ================================================================================================================
namespace {

struct WideConv {
    const int* nodes; unsigned long long nodeBytes; int layout; int emptyLink; unsigned long long woopRows;
    unsigned* out; int capacity;            // wide nodes that fit `out`
    unsigned long long* pubOf;              // per wide node, one word written at once: depth << 32 | (binary link + 1); 0 = not yet published
    int* ctr;                               // [0] ticket [1] allocated [2] processed [3] max depth [4] error [5] binary nodes consumed
};
enum : int { kConvBadLink = 1, kConvCycle = 2, kConvBadLeaf = 3, kConvTooLarge = 4 };

struct DCand { float lo[3], hi[3]; int link; };

__device__ __forceinline__ float dcand_area(const DCand& c)
{
    const float dx = __fsub_rn(c.hi[0], c.lo[0]), dy = __fsub_rn(c.hi[1], c.lo[1]), dz = __fsub_rn(c.hi[2], c.lo[2]);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dy), __fmul_rn(dy, dz)), __fmul_rn(dz, dx));
}

__device__ __forceinline__ bool dread_children(const WideConv& a, int addr, DCand out[2])
{
    const unsigned long long byteOfs = (a.layout == Layout_Compact2) ? (unsigned long long)(unsigned)addr * 16ull : (unsigned long long)(unsigned)addr;
    if (byteOfs % 64ull || byteOfs + 64ull > a.nodeBytes) return false;
    const int4* w = reinterpret_cast<const int4*>(a.nodes + byteOfs / 4);
    const int4 w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
    out[0].lo[0] = __int_as_float(w0.x); out[0].hi[0] = __int_as_float(w0.y); out[0].lo[1] = __int_as_float(w0.z); out[0].hi[1] = __int_as_float(w0.w);
    out[1].lo[0] = __int_as_float(w1.x); out[1].hi[0] = __int_as_float(w1.y); out[1].lo[1] = __int_as_float(w1.z); out[1].hi[1] = __int_as_float(w1.w);
    out[0].lo[2] = __int_as_float(w2.x); out[0].hi[2] = __int_as_float(w2.y); out[1].lo[2] = __int_as_float(w2.z); out[1].hi[2] = __int_as_float(w2.w);
    out[0].link = w3.x; out[1].link = w3.y;
    return true;
}

// quantise_axis() of the host routine, operation for operation
__device__ __forceinline__ void dquantise_axis(const DCand* c, int n, int axis, float& p, float& sPrime, unsigned& qloPacked, unsigned& qhiPacked)
{
    float mn = c[0].lo[axis], mx = c[0].hi[axis];
    for (int i = 1; i < n; i++) { mn = (c[i].lo[axis] < mn) ? c[i].lo[axis] : mn; mx = (mx < c[i].hi[axis]) ? c[i].hi[axis] : mx; }   // std::min / std::max
    p = mn;
    const double ext = __dsub_rn((double)mx, (double)mn);
    float s = __double2float_rn(__ddiv_rn(ext, 253.0));
    if (!(s > 1.0e-30f)) s = 1.0e-30f;
    while (__dmul_rn((double)s, 253.0) < ext) s = nextafterf(s, __int_as_float(0x7f800000));
    const double dp = (double)p, ds = (double)s;
    qloPacked = 0; qhiPacked = 0;
    for (int i = 0; i < 4; i++) {
        int qa = 255, qb = 0;                                                // unused slot: inverted box
        if (i < n) {
            const double lo = (double)c[i].lo[axis], hi = (double)c[i].hi[axis];
            qa = (int)floor(__dsub_rn(__ddiv_rn(__dsub_rn(lo, dp), ds), 0.125));
            qa = min(max(qa, 0), 255);
            while (qa > 0 && __dadd_rn(dp, __dmul_rn((double)qa, ds)) > lo) qa--;
            qb = (int)ceil(__dadd_rn(__ddiv_rn(__dsub_rn(hi, dp), ds), 0.125));
            qb = min(max(qb, 0), 255);
            while (qb < 255 && __dadd_rn(dp, __dmul_rn((double)qb, ds)) < hi) qb++;
        }
        qloPacked |= (unsigned)qa << (8 * i); qhiPacked |= (unsigned)qb << (8 * i);
    }
    sPrime = __fmul_rn(s, 32768.0f);
}

__global__ void __launch_bounds__(128) wide4_convert_kernel(WideConv a)
{
    volatile int* ctr = a.ctr;
    volatile unsigned long long* pubOf = a.pubOf;
    for (;;) {
        const int w = atomicAdd(a.ctr + 0, 1);
        unsigned long long pubw = 0;
        for (;;) {
            if (ctr[4]) return;
            if (w < a.capacity) pubw = pubOf[w];
            if (pubw) break;
            // processed BEFORE allocated (the fence keeps the two loads in that order): a node is allocated before its parent counts as
            // processed, so done == alloc read in this order means that nothing is in flight and nothing more will be allocated
            const int done = ctr[2];
            __threadfence();
            const int alloc = ctr[1];
            if (done == alloc && w >= alloc) return;
        }
        const int pub = (int)(unsigned)pubw, depth = (int)(pubw >> 32);
        atomicMax(a.ctr + 3, depth);
        DCand c[4];
        int n = 2, consumed = 1;
        int err = 0;
        if (!dread_children(a, pub - 1, c)) err = kConvBadLink;
        // open the inner child with the largest surface area until four children are held (or only leaves are left)
        while (!err && n < 4) {
            int best = -1;
            float bestArea = -1.0f;
            for (int i = 0; i < n; i++)
                if (c[i].link >= 0) { const float ar = dcand_area(c[i]); if (ar > bestArea || best < 0) { bestArea = ar; best = i; } }
            if (best < 0) break;
            DCand two[2];
            if (!dread_children(a, c[best].link, two)) { err = kConvBadLink; break; }
            consumed++;
            c[best] = two[0];
            c[n++] = two[1];
        }
        if (!err && (unsigned long long)(atomicAdd(a.ctr + 5, consumed) + consumed) > a.nodeBytes / 64ull) err = kConvCycle;
        int inner = 0;
        if (!err)
            for (int i = 0; i < n; i++) {
                if (c[i].link >= 0) inner++;
                else if ((unsigned long long)(unsigned)~c[i].link >= a.woopRows) err = kConvBadLeaf;
            }
        int base = 0;
        if (!err && inner) {
            base = atomicAdd(a.ctr + 1, inner);
            if (base + inner > a.capacity) err = (a.capacity >= (int)kEntrypointSentinel - 4) ? kConvTooLarge : kConvCycle;
        }
        if (err) { atomicCAS(a.ctr + 4, 0, err); return; }
        unsigned wd[16];
        float p, sp;
        unsigned qlo[3], qhi[3];
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            dquantise_axis(c, n, ax, p, sp, qlo[ax], qhi[ax]);
            wd[ax] = __float_as_uint(p); wd[3 + ax] = __float_as_uint(sp);
        }
        wd[6] = qlo[0]; wd[7] = qlo[1]; wd[8] = qlo[2]; wd[9] = qhi[0]; wd[10] = qhi[1]; wd[11] = qhi[2];
        int next = base;
        for (int i = 0; i < 4; i++) {
            int link = a.emptyLink;
            if (i < n) {
                if (c[i].link < 0) link = c[i].link;
                else {
                    link = next++;
                    pubOf[link] = ((unsigned long long)(depth + 1) << 32) | (unsigned)(c[i].link + 1);
                }
            }
            wd[12 + i] = (unsigned)link;
        }
        uint4* o = reinterpret_cast<uint4*>(a.out + (size_t)w * 16);
        o[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]); o[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
        o[2] = make_uint4(wd[8], wd[9], wd[10], wd[11]); o[3] = make_uint4(wd[12], wd[13], wd[14], wd[15]);
        __threadfence();
        atomicAdd(a.ctr + 2, 1);
    }
}

} // namespace

// `wide` is grown to hold one wide node per binary node (the upper bound); `scratch` holds the publication / depth arrays and counters.
cudaError_t convert_compact_to_wide4_device(const void* dNodes, size_t nodeBytes, int layout, size_t woopRows, DevBuf& wide, DevBuf& scratch,
                                            size_t* outWideBytes, int* outMaxDepth, int numSMs, cudaStream_t stream, int* launches, std::string* err)
{
    auto fail = [&](const char* m) { if (err) *err = m; return cudaErrorInvalidValue; };
    if (layout != Layout_Compact && layout != Layout_Compact2) return fail("Wide4 is derived from BVHLayout_Compact / Compact2");
    if (!dNodes || nodeBytes < 64 || nodeBytes % 64 || woopRows < 1) return fail("malformed BVH: empty node or triangle buffer");
    const size_t numBinary = nodeBytes / 64;
    if (numBinary >= (size_t)kEntrypointSentinel) return fail("BVH too large for the Wide4 form");
    cudaError_t e;
    if ((e = wide.reserve(numBinary * 64)) != cudaSuccess) return e;
    const size_t ctrBytes = 64, arr = numBinary * 8;
    if ((e = scratch.reserve(ctrBytes + arr)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(scratch.p, 0, ctrBytes + arr, stream)) != cudaSuccess) return e;      // counters + publication words
    WideConv a;
    a.nodes = static_cast<const int*>(dNodes); a.nodeBytes = nodeBytes; a.layout = layout;
    a.emptyLink = ~(int)(woopRows - 1); a.woopRows = woopRows;
    a.out = wide.as<unsigned>(); a.capacity = (int)numBinary;
    a.ctr = scratch.as<int>();
    a.pubOf = reinterpret_cast<unsigned long long*>(scratch.as<char>() + ctrBytes);
    const unsigned long long rootPub = (1ull << 32) | 1ull;                    // root: wide node 0 folds the binary node at link 0, depth 1
    const int one = 1;                                                         // ... and is the one node allocated so far
    if ((e = cudaMemcpyAsync(a.pubOf, &rootPub, 8, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(a.ctr + 1, &one, 4, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
    int grid = (int)((numBinary / 2 + 127) / 128);
    if (grid > numSMs * 4) grid = numSMs * 4;
    if (grid < 1) grid = 1;
    wide4_convert_kernel<<<grid, 128, 0, stream>>>(a);
    if (launches) *launches += 1;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    int h[8];
    if ((e = cudaMemcpyAsync(h, a.ctr, 32, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    switch (h[4]) {
    case 0: break;
    case kConvBadLink: return fail("malformed BVH: child link outside the node buffer");
    case kConvCycle: return fail("malformed BVH: the node links form a cycle");
    case kConvBadLeaf: return fail("malformed BVH: leaf link outside the triangle buffer");
    default: return fail("BVH too large for the Wide4 form");
    }
    if (outWideBytes) *outWideBytes = (size_t)h[1] * 64;
    if (outMaxDepth) *outMaxDepth = h[3];
    if (h[3] > kWideMaxDepth) return fail("BVH too deep for the Wide4 kernel's traversal stack (use a binary kernel)");
    return cudaSuccess;
}

// ================================================================================================================
// device: traversal
// ================================================================================================================
namespace {

constexpr int kWideStackSize = 128;            // 1 + 3 * kWideMaxDepth entries at most (three pushes per level); the conversion refuses deeper trees
constexpr int kWideBlock = 128;

__device__ __forceinline__ float wfmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float wfmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ void wld256_nc(const float4* p, float4& a, float4& b)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void wld256_cs(const float4* p, float4& a, float4& b)
{
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
// quantised plane byte I of `q` as the float 1 + q * 2^-15: bytes (0x00, q.I, 0x80, 0x3F).  `one` is 0x3F800000 passed as a kernel
// argument: a value ptxas cannot fold, so the SELECTOR becomes the immediate operand of PRMT (it takes one immediate) instead of a
// register that would have to be materialised for every plane
template <int I> __device__ __forceinline__ float qplane(unsigned q, unsigned one)
{
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(q), "r"(one), "n"(0x7604 | (I << 4)));
    return __uint_as_float(r);
}
__device__ __forceinline__ void cas(int& a, int& b) { const int lo = min(a, b), hi = max(a, b); a = lo; b = hi; }

template <int BLOCK, int SMEM_N, bool FAST, bool WIDE_RAYS>
__global__ void __launch_bounds__(BLOCK)
trace_wide4_kernel(int numRays, int anyHit, int fetchThreshold, int nodeExit, unsigned oneBits,
                   const float4* __restrict__ rays, int4* __restrict__ results,
                   const float4* __restrict__ wnodes, const float4* __restrict__ woop,
                   const int* __restrict__ triIndices, int* __restrict__ warpCounter, int* __restrict__ errorFlag,
                   const BatchTable* __restrict__ batches)
{
    __shared__ int s_stack[(SMEM_N > 0 ? SMEM_N : 1) * BLOCK];
    int l_stack[(kWideStackSize > SMEM_N) ? (kWideStackSize - SMEM_N) : 1];

    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;
    int* const sbase = s_stack + tid;

#define NT_PUSH(v)  do { ++sp; if (sp < SMEM_N) sbase[sp * BLOCK] = (v); else if (sp < kWideStackSize) l_stack[sp - SMEM_N] = (v); else { --sp; *(volatile int*)errorFlag = 1; } } while (0)
#define NT_POP(dst) do { (dst) = (sp < SMEM_N) ? sbase[sp * BLOCK] : l_stack[sp - SMEM_N]; --sp; } while (0)

    int   rayidx = -1;
    float origx = 0, origy = 0, origz = 0, dirx = 0, diry = 0, dirz = 0, tmin = 0;
    float idirx = 0, idiry = 0, idirz = 0, oodx = 0, oody = 0, oodz = 0;
    int   sp = 0;
    int   leafAddr = 0;
    int   nodeAddr = kEntrypointSentinel;
    int   hitIndex = -1;
    float hitT = 0, hitU = 0, hitV = 0;
    bool  alive = true;
    const unsigned one = oneBits;           // 0x3F800000 as a kernel argument: see qplane()

    for (;;) {
        // ---------------- ray fetch: one atomicAdd per warp refill (as nt_trace.cu) ----------------
        const bool need = alive && (nodeAddr == kEntrypointSentinel);
        const unsigned needMask = __ballot_sync(0xffffffffu, need);
        if (needMask) {
            const int n = __popc(needMask);
            const int rank = __popc(needMask & ((1u << lane) - 1u));
            const int leader = __ffs(needMask) - 1;
            int base = 0;
            if ((int)lane == leader) base = atomicAdd(warpCounter, n);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (need) rayidx = base + rank;
        }
        if (need) {
            if (rayidx >= numRays) { alive = false; rayidx = -1; }
            else {
                float4 o, d;
                const float4* rp = rays + rayidx * 2;
                if (batches) { const int b = batch_of(batches, rayidx); rp = batches->rays[b] + (size_t)(rayidx - __ldg(&batches->start[b])) * 2; }
                if (WIDE_RAYS) wld256_cs(rp, o, d);
                else { o = __ldcs(rp + 0); d = __ldcs(rp + 1); }
                origx = o.x; origy = o.y; origz = o.z; tmin = o.w;
                dirx = d.x; diry = d.y; dirz = d.z; hitT = d.w;
                const float ooeps = exp2f(-80.0f);                      // fermi_speculative_while_while.cu:94-98
                idirx = 1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
                idiry = 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
                idirz = 1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
                oodx = origx * idirx; oody = origy * idiry; oodz = origz * idirz;
                sp = 0;
                if (SMEM_N > 0) sbase[0] = kEntrypointSentinel; else l_stack[0] = kEntrypointSentinel;
                leafAddr = 0;
                nodeAddr = 0;
                hitIndex = -1;
                hitU = 0.0f; hitV = 0.0f;
            }
        }
        if (!__any_sync(0xffffffffu, alive)) break;

        // ---------------- traversal ----------------
        while (nodeAddr != kEntrypointSentinel) {
            while ((unsigned)nodeAddr < (unsigned)kEntrypointSentinel) {
                const float4* ptr = wnodes + (size_t)nodeAddr * 4;
                float4 A0, A1, B0, B1;
                wld256_nc(ptr, A0, A1);            // (p.x p.y p.z s'.x) (s'.y s'.z qlo.x qlo.y)
                wld256_nc(ptr + 2, B0, B1);        // (qlo.z qhi.x qhi.y qhi.z) (link0..3)

                const float ax = A0.w * idirx, ay = A1.x * idiry, az = A1.y * idirz;
                const float bx = fmaf(A0.x, idirx, -oodx) - ax;
                const float by = fmaf(A0.y, idiry, -oody) - ay;
                const float bz = fmaf(A0.z, idirz, -oodz) - az;
                // near / far plane bytes by the sign of the ray direction
                const bool ngx = idirx < 0.0f, ngy = idiry < 0.0f, ngz = idirz < 0.0f;
                const unsigned qlx = __float_as_uint(A1.z), qly = __float_as_uint(A1.w), qlz = __float_as_uint(B0.x);
                const unsigned qhx = __float_as_uint(B0.y), qhy = __float_as_uint(B0.z), qhz = __float_as_uint(B0.w);
                const unsigned nx = ngx ? qhx : qlx, fx = ngx ? qlx : qhx;
                const unsigned ny = ngy ? qhy : qly, fy = ngy ? qly : qhy;
                const unsigned nz = ngz ? qhz : qlz, fz = ngz ? qlz : qhz;

                int k0, k1, k2, k3;
#define NT_CHILD(I, K)                                                                                                    \
                {                                                                                                         \
                    const float tn = wfmax3(fmaf(qplane<I>(nx, one), ax, bx), fmaf(qplane<I>(ny, one), ay, by),           \
                                            fmaxf(fmaf(qplane<I>(nz, one), az, bz), tmin));                               \
                    const float tf = wfmin3(fmaf(qplane<I>(fx, one), ax, bx), fmaf(qplane<I>(fy, one), ay, by),           \
                                            fminf(fmaf(qplane<I>(fz, one), az, bz), hitT));                               \
                    K = (tn <= tf) ? ((__float_as_int(tn) & ~3) | I) : 0x7fffffff;                                        \
                }
                NT_CHILD(0, k0) NT_CHILD(1, k1) NT_CHILD(2, k2) NT_CHILD(3, k3)
#undef NT_CHILD
                // children in order of entry distance (the child slot rides in the two low mantissa bits of the key)
                cas(k0, k1); cas(k2, k3); cas(k0, k2); cas(k1, k3); cas(k1, k2);
                const int l0 = __float_as_int(B1.x), l1 = __float_as_int(B1.y), l2 = __float_as_int(B1.z), l3 = __float_as_int(B1.w);
#define NT_LINK(K) (((K) & 2) ? (((K) & 1) ? l3 : l2) : (((K) & 1) ? l1 : l0))
                // Measured and dropped (profiles/r2_summary.md): the three pushes without branches while they fit the
                // shared-memory part of the stack (-9 %: computing all three links costs more than the divergent ifs save) and
                // __launch_bounds__(128, 12) for 12 instead of 10 CTAs per SM (40 registers, 42 bytes of spills: -1..-6 %).
                if (k0 == 0x7fffffff) {
                    NT_POP(nodeAddr);
                } else {
                    nodeAddr = NT_LINK(k0);
                    if (k1 != 0x7fffffff) {
                        if (k2 != 0x7fffffff) {
                            if (k3 != 0x7fffffff) NT_PUSH(NT_LINK(k3));
                            NT_PUSH(NT_LINK(k2));
                        }
                        NT_PUSH(NT_LINK(k1));
                    }
                }
#undef NT_LINK
                // first leaf => postpone and continue traversal (speculative while-while, as nt_trace.cu)
                if (nodeAddr < 0 && leafAddr >= 0) {
                    leafAddr = nodeAddr;
                    NT_POP(nodeAddr);
                }
                // all lanes hold a leaf => process them.  Also when fewer than nodeExit lanes are left in this loop: a Wide4 node step costs
                // ~150 instructions, too many to issue for a handful of lanes; the ones still searching sit out the leaf phase and come back
                // (measured on the bench frame, profiles/r2_summary.md: nodeExit 8 = +10 % AO, +6 % diffuse over 0)
                {
                    const unsigned am = __activemask();
                    if (!__any_sync(am, leafAddr >= 0) || __popc(am) < nodeExit) break;
                }
            }

            // postponed leaves: the reference's Woop test (Util.cpp:99-127), identical to nt_trace.cu
            while (leafAddr < 0) {
                int triAddr = ~leafAddr;
                float4 v00 = __ldg(woop + triAddr);
                for (;;) {
                    if (__float_as_int(v00.x) == (int)0x80000000) break;
                    float t;
                    if (FAST) {
                        const float Oz = v00.w - origx * v00.x - origy * v00.y - origz * v00.z;
                        t = Oz * __fdividef(1.0f, dirx * v00.x + diry * v00.y + dirz * v00.z);
                    } else {
                        const float Oz = __fsub_rn(__fsub_rn(__fsub_rn(v00.w, __fmul_rn(origx, v00.x)), __fmul_rn(origy, v00.y)), __fmul_rn(origz, v00.z));
                        const float dd = __fadd_rn(__fadd_rn(__fmul_rn(dirx, v00.x), __fmul_rn(diry, v00.y)), __fmul_rn(dirz, v00.z));
                        t = __fmul_rn(Oz, __frcp_rn(dd));
                    }
                    if (t > tmin && t < hitT) {
                        const float4 v11 = __ldg(woop + triAddr + 1);
                        float u;
                        if (FAST) u = (v11.w + origx * v11.x + origy * v11.y + origz * v11.z) + t * (dirx * v11.x + diry * v11.y + dirz * v11.z);
                        else {
                            const float Ou = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v11.x, origx), __fmul_rn(v11.y, origy)), __fmul_rn(v11.z, origz)), v11.w);
                            const float Du = __fadd_rn(__fadd_rn(__fmul_rn(v11.x, dirx), __fmul_rn(v11.y, diry)), __fmul_rn(v11.z, dirz));
                            u = __fadd_rn(Ou, __fmul_rn(t, Du));
                        }
                        if (u >= 0.0f) {
                            const float4 v22 = __ldg(woop + triAddr + 2);
                            float v;
                            if (FAST) v = (v22.w + origx * v22.x + origy * v22.y + origz * v22.z) + t * (dirx * v22.x + diry * v22.y + dirz * v22.z);
                            else {
                                const float Ov = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v22.x, origx), __fmul_rn(v22.y, origy)), __fmul_rn(v22.z, origz)), v22.w);
                                const float Dv = __fadd_rn(__fadd_rn(__fmul_rn(v22.x, dirx), __fmul_rn(v22.y, diry)), __fmul_rn(v22.z, dirz));
                                v = __fadd_rn(Ov, __fmul_rn(t, Dv));
                            }
                            if (v >= 0.0f && (FAST ? (u + v) : __fadd_rn(u, v)) <= 1.0f) {
                                hitT = t; hitU = u; hitV = v;
                                hitIndex = triAddr;
                                if (anyHit) { nodeAddr = kEntrypointSentinel; break; }
                            }
                        }
                    }
                    triAddr += 3;
                    v00 = __ldg(woop + triAddr);
                }
                leafAddr = nodeAddr;
                if (nodeAddr < 0) NT_POP(nodeAddr);
            }
            if (__popc(__activemask()) < fetchThreshold) break;
        }

        // ---------------- result ----------------
        if (rayidx >= 0 && nodeAddr == kEntrypointSentinel) {
            int id = hitIndex;
            if (id != -1) id = __ldg(triIndices + id);
            int4* op = results + rayidx;
            if (batches) { const int b = batch_of(batches, rayidx); op = batches->results[b] + (rayidx - __ldg(&batches->start[b])); }
            __stcs(op, make_int4(id, __float_as_int(hitT), __float_as_int(hitU), __float_as_int(hitV)));
            rayidx = -1;
        }
    }
#undef NT_PUSH
#undef NT_POP
}

// ================================================================================================================
// Multi-ray lanes ("mr" kernels).  profiles/r2a_binding_*.json: both the binary and the Wide4 kernel execute their instructions
// with 11-12 of 32 lanes on secondary rays — a lane whose ray waits for the other phase (holds a leaf during the node phase, sits at
// a node during the leaf phase, or has finished) idles, and issue slots / the ALU pipe are what bind.  Here every lane owns TWO
// rays whose state lives in shared memory (64 bytes + a short stack each); a warp step is ONE phase chosen by vote:
//   NODE   every lane steps whichever of its two rays sits at an inner node,
//   LEAF   every lane runs the Woop tests of the leaf whichever of its rays sits at,
//   FETCH  every lane with a free slot takes a new ray from the global counter (one atomicAdd per warp, as before).
// A lane idles only when neither of its rays fits the phase.  Per-ray traversal order is the depth-first order of the tree, exactly
// as in the one-ray kernels (which only postpone leaves, never reorder them), so results are bit-identical to theirs, any-hit ids
// included.  Ray state: f0 = (orig.xyz, tmin)  f1 = (dir.xyz, hitIndex)  f2 = (idir.xyz, hitT)  f3 = (hitU, hitV, rayidx, -).
// ================================================================================================================
enum : int { kFmtCompact = 0, kFmtCompact2 = 1, kFmtWide4 = 2 };
constexpr int kMrLocal = 124;                  // stack entries beyond the shared-memory part, per ray (local memory, rarely touched)

template <int BLOCK, int SMEM_N, bool FAST, int FMT>
__global__ void __launch_bounds__(BLOCK)
trace_mr_kernel(int numRays, int anyHit, int fetchThreshold, int leafThreshold, unsigned oneBits,
                const float4* __restrict__ rays, int4* __restrict__ results,
                const float4* __restrict__ nodes, const float4* __restrict__ woop,
                const int* __restrict__ triIndices, int* __restrict__ warpCounter, int* __restrict__ errorFlag)
{
    extern __shared__ float4 s_dyn[];
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;
    float4* const F = s_dyn + tid;                                              // field f of slot k: F[(f * 2 + k) * BLOCK]
    int* const S = reinterpret_cast<int*>(s_dyn + 8 * BLOCK) + tid;             // stack entry e of slot k: S[(k * SMEM_N + e) * BLOCK]
    int l_stack[2][kMrLocal];
    const unsigned one = oneBits;

    int cur0 = kEntrypointSentinel, cur1 = kEntrypointSentinel;                 // sentinel = free slot
    int sp0 = 0, sp1 = 0;
    bool more = true;                                                           // warp-uniform: the global counter still has rays

#define MR_PUSH(v)  do { if (sp < SMEM_N) Sk[sp * BLOCK] = (v); else if (sp < SMEM_N + kMrLocal) l_stack[k][sp - SMEM_N] = (v); else { --sp; *(volatile int*)errorFlag = 1; } ++sp; } while (0)
#define MR_POP(dst) do { if (sp == 0) (dst) = kEntrypointSentinel; else { --sp; (dst) = (sp < SMEM_N) ? Sk[sp * BLOCK] : l_stack[k][sp - SMEM_N]; } } while (0)
    // ray in slot k has finished: one int4 result store (id remapped through triIndex; a miss keeps t = tmax)
#define MR_FINISH()                                                                                                          \
    do {                                                                                                                     \
        const float4 r3 = Fk[6 * BLOCK];                                                                                     \
        const int hi = __float_as_int(Fk[2 * BLOCK].w);                                                                      \
        const float ht = Fk[4 * BLOCK].w;                                                                                    \
        int id = hi;                                                                                                         \
        if (id != -1) id = __ldg(triIndices + id);                                                                           \
        __stcs(results + __float_as_int(r3.z), make_int4(id, __float_as_int(ht), __float_as_int(r3.x), __float_as_int(r3.y))); \
    } while (0)

    for (;;) {
        const bool e0 = cur0 == kEntrypointSentinel, e1 = cur1 == kEntrypointSentinel;
        const bool n0 = (unsigned)cur0 < (unsigned)kEntrypointSentinel, n1 = (unsigned)cur1 < (unsigned)kEntrypointSentinel;
        const unsigned mE = more ? __ballot_sync(0xffffffffu, e0 || e1) : 0u;
        const unsigned mN = __ballot_sync(0xffffffffu, n0 || n1);
        const unsigned mL = __ballot_sync(0xffffffffu, cur0 < 0 || cur1 < 0);
        const int cN = __popc(mN), cL = __popc(mL);

        if (mE && (__popc(mE) >= fetchThreshold || (mN | mL) == 0u)) {
            // ---------------- FETCH ----------------
            const bool need = e0 || e1;
            const int n = __popc(mE);
            const int rank = __popc(mE & ((1u << lane) - 1u));
            const int leader = __ffs(mE) - 1;
            int base = 0;
            if ((int)lane == leader) base = atomicAdd(warpCounter, n);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + n >= numRays) more = false;
            const int rayidx = base + rank;
            if (need && rayidx < numRays) {
                const int k = e0 ? 0 : 1;
                float4* const Fk = F + k * BLOCK;
                float4 o, d;
                wld256_cs(rays + rayidx * 2, o, d);
                const float ooeps = exp2f(-80.0f);                      // fermi_speculative_while_while.cu:94-98
                const float ix = 1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
                const float iy = 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
                const float iz = 1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
                Fk[0] = o;                                                              // (orig, tmin)
                Fk[2 * BLOCK] = make_float4(d.x, d.y, d.z, __int_as_float(-1));         // (dir, hitIndex)
                Fk[4 * BLOCK] = make_float4(ix, iy, iz, d.w);                           // (idir, hitT = tmax)
                Fk[6 * BLOCK] = make_float4(0.0f, 0.0f, __int_as_float(rayidx), 0.0f);  // (hitU, hitV, rayidx, -)
                if (k) { cur1 = 0; sp1 = 0; } else { cur0 = 0; sp0 = 0; }
            }
        } else if ((mN | mL) == 0u) {
            break;
        } else if (mL && (cN == 0 || cL >= leafThreshold || cL >= cN)) {
            // ---------------- LEAF: Woop tests of one leaf (Util.cpp:99-127, identical to nt_trace.cu) ----------------
            if (cur0 < 0 || cur1 < 0) {
                const int k = (cur0 < 0) ? 0 : 1;
                float4* const Fk = F + k * BLOCK;
                int* const Sk = S + k * SMEM_N * BLOCK;
                int sp = k ? sp1 : sp0;
                int nodeAddr;
                const float4 f0 = Fk[0], f1 = Fk[2 * BLOCK];
                const float origx = f0.x, origy = f0.y, origz = f0.z, tmin = f0.w, dirx = f1.x, diry = f1.y, dirz = f1.z;
                float hitT = Fk[4 * BLOCK].w, hitU = 0.0f, hitV = 0.0f;
                int hitIndex = -1;
                bool done = false;
                int triAddr = ~(k ? cur1 : cur0);
                float4 v00 = __ldg(woop + triAddr);
                for (;;) {
                    if (__float_as_int(v00.x) == (int)0x80000000) break;
                    float t;
                    if (FAST) {
                        const float Oz = v00.w - origx * v00.x - origy * v00.y - origz * v00.z;
                        t = Oz * __fdividef(1.0f, dirx * v00.x + diry * v00.y + dirz * v00.z);
                    } else {
                        const float Oz = __fsub_rn(__fsub_rn(__fsub_rn(v00.w, __fmul_rn(origx, v00.x)), __fmul_rn(origy, v00.y)), __fmul_rn(origz, v00.z));
                        const float dd = __fadd_rn(__fadd_rn(__fmul_rn(dirx, v00.x), __fmul_rn(diry, v00.y)), __fmul_rn(dirz, v00.z));
                        t = __fmul_rn(Oz, __frcp_rn(dd));
                    }
                    if (t > tmin && t < hitT) {
                        const float4 v11 = __ldg(woop + triAddr + 1);
                        float u;
                        if (FAST) u = (v11.w + origx * v11.x + origy * v11.y + origz * v11.z) + t * (dirx * v11.x + diry * v11.y + dirz * v11.z);
                        else {
                            const float Ou = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v11.x, origx), __fmul_rn(v11.y, origy)), __fmul_rn(v11.z, origz)), v11.w);
                            const float Du = __fadd_rn(__fadd_rn(__fmul_rn(v11.x, dirx), __fmul_rn(v11.y, diry)), __fmul_rn(v11.z, dirz));
                            u = __fadd_rn(Ou, __fmul_rn(t, Du));
                        }
                        if (u >= 0.0f) {
                            const float4 v22 = __ldg(woop + triAddr + 2);
                            float v;
                            if (FAST) v = (v22.w + origx * v22.x + origy * v22.y + origz * v22.z) + t * (dirx * v22.x + diry * v22.y + dirz * v22.z);
                            else {
                                const float Ov = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v22.x, origx), __fmul_rn(v22.y, origy)), __fmul_rn(v22.z, origz)), v22.w);
                                const float Dv = __fadd_rn(__fadd_rn(__fmul_rn(v22.x, dirx), __fmul_rn(v22.y, diry)), __fmul_rn(v22.z, dirz));
                                v = __fadd_rn(Ov, __fmul_rn(t, Dv));
                            }
                            if (v >= 0.0f && (FAST ? (u + v) : __fadd_rn(u, v)) <= 1.0f) {
                                hitT = t; hitU = u; hitV = v;
                                hitIndex = triAddr;
                                if (anyHit) { done = true; break; }
                            }
                        }
                    }
                    triAddr += 3;
                    v00 = __ldg(woop + triAddr);
                }
                if (hitIndex != -1) {                        // a closer hit in this leaf: publish it to the ray's state
                    reinterpret_cast<float*>(Fk + 2 * BLOCK)[3] = __int_as_float(hitIndex);
                    reinterpret_cast<float*>(Fk + 4 * BLOCK)[3] = hitT;
                    reinterpret_cast<float2*>(Fk + 6 * BLOCK)[0] = make_float2(hitU, hitV);
                }
                if (done) nodeAddr = kEntrypointSentinel; else MR_POP(nodeAddr);
                if (nodeAddr == kEntrypointSentinel) MR_FINISH();
                if (k) { cur1 = nodeAddr; sp1 = sp; } else { cur0 = nodeAddr; sp0 = sp; }
            }
        } else {
            // ---------------- NODE: one inner-node step ----------------
            if (n0 || n1) {
                const int k = n0 ? 0 : 1;
                float4* const Fk = F + k * BLOCK;
                int* const Sk = S + k * SMEM_N * BLOCK;
                int sp = k ? sp1 : sp0;
                int nodeAddr = k ? cur1 : cur0;
                const float4 f0 = Fk[0], f2 = Fk[4 * BLOCK];
                const float tmin = f0.w, hitT = f2.w;
                const float idirx = f2.x, idiry = f2.y, idirz = f2.z;
                const float oodx = f0.x * idirx, oody = f0.y * idiry, oodz = f0.z * idirz;
                if (FMT == kFmtWide4) {
                    const float4* ptr = nodes + (size_t)nodeAddr * 4;
                    float4 A0, A1, B0, B1;
                    wld256_nc(ptr, A0, A1);
                    wld256_nc(ptr + 2, B0, B1);
                    const float ax = A0.w * idirx, ay = A1.x * idiry, az = A1.y * idirz;
                    const float bx = fmaf(A0.x, idirx, -oodx) - ax;
                    const float by = fmaf(A0.y, idiry, -oody) - ay;
                    const float bz = fmaf(A0.z, idirz, -oodz) - az;
                    const bool ngx = idirx < 0.0f, ngy = idiry < 0.0f, ngz = idirz < 0.0f;
                    const unsigned qlx = __float_as_uint(A1.z), qly = __float_as_uint(A1.w), qlz = __float_as_uint(B0.x);
                    const unsigned qhx = __float_as_uint(B0.y), qhy = __float_as_uint(B0.z), qhz = __float_as_uint(B0.w);
                    const unsigned nx = ngx ? qhx : qlx, fx = ngx ? qlx : qhx;
                    const unsigned ny = ngy ? qhy : qly, fy = ngy ? qly : qhy;
                    const unsigned nz = ngz ? qhz : qlz, fz = ngz ? qlz : qhz;
                    int k0, k1, k2, k3;
#define NT_CHILD(I, K)                                                                                                    \
                    {                                                                                                     \
                        const float tn = wfmax3(fmaf(qplane<I>(nx, one), ax, bx), fmaf(qplane<I>(ny, one), ay, by),       \
                                                fmaxf(fmaf(qplane<I>(nz, one), az, bz), tmin));                           \
                        const float tf = wfmin3(fmaf(qplane<I>(fx, one), ax, bx), fmaf(qplane<I>(fy, one), ay, by),       \
                                                fminf(fmaf(qplane<I>(fz, one), az, bz), hitT));                           \
                        K = (tn <= tf) ? ((__float_as_int(tn) & ~3) | I) : 0x7fffffff;                                    \
                    }
                    NT_CHILD(0, k0) NT_CHILD(1, k1) NT_CHILD(2, k2) NT_CHILD(3, k3)
#undef NT_CHILD
                    cas(k0, k1); cas(k2, k3); cas(k0, k2); cas(k1, k3); cas(k1, k2);
                    const int l0 = __float_as_int(B1.x), l1 = __float_as_int(B1.y), l2 = __float_as_int(B1.z), l3 = __float_as_int(B1.w);
#define NT_LINK(K) (((K) & 2) ? (((K) & 1) ? l3 : l2) : (((K) & 1) ? l1 : l0))
                    if (k0 == 0x7fffffff) {
                        MR_POP(nodeAddr);
                    } else {
                        nodeAddr = NT_LINK(k0);
                        if (k1 != 0x7fffffff) {
                            if (k2 != 0x7fffffff) {
                                if (k3 != 0x7fffffff) MR_PUSH(NT_LINK(k3));
                                MR_PUSH(NT_LINK(k2));
                            }
                            MR_PUSH(NT_LINK(k1));
                        }
                    }
#undef NT_LINK
                } else {
                    // binary Compact / Compact2 node (CudaBVH.hpp:43-47), slab test as nt_trace.cu
                    const float4* ptr = (FMT == kFmtCompact2) ? nodes + nodeAddr
                                                              : reinterpret_cast<const float4*>(reinterpret_cast<const char*>(nodes) + nodeAddr);
                    float4 n0xy, n1xy, nz, cn;
                    wld256_nc(ptr, n0xy, n1xy);
                    wld256_nc(ptr + 2, nz, cn);
                    int c0idx = __float_as_int(cn.x), c1idx = __float_as_int(cn.y);
                    const float c0lox = n0xy.x * idirx - oodx, c0hix = n0xy.y * idirx - oodx;
                    const float c0loy = n0xy.z * idiry - oody, c0hiy = n0xy.w * idiry - oody;
                    const float c0loz = nz.x * idirz - oodz,   c0hiz = nz.y * idirz - oodz;
                    const float c1loz = nz.z * idirz - oodz,   c1hiz = nz.w * idirz - oodz;
                    const float c0min = wfmax3(fminf(c0lox, c0hix), fminf(c0loy, c0hiy), fmaxf(fminf(c0loz, c0hiz), tmin));
                    const float c0max = wfmin3(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy), fminf(fmaxf(c0loz, c0hiz), hitT));
                    const float c1lox = n1xy.x * idirx - oodx, c1hix = n1xy.y * idirx - oodx;
                    const float c1loy = n1xy.z * idiry - oody, c1hiy = n1xy.w * idiry - oody;
                    const float c1min = wfmax3(fminf(c1lox, c1hix), fminf(c1loy, c1hiy), fmaxf(fminf(c1loz, c1hiz), tmin));
                    const float c1max = wfmin3(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy), fminf(fmaxf(c1loz, c1hiz), hitT));
                    const bool trav0 = (c0max >= c0min), trav1 = (c1max >= c1min);
                    if (!trav0 && !trav1) {
                        MR_POP(nodeAddr);
                    } else {
                        nodeAddr = trav0 ? c0idx : c1idx;
                        if (trav0 && trav1) {
                            if (c1min < c0min) { const int t = nodeAddr; nodeAddr = c1idx; c1idx = t; }
                            MR_PUSH(c1idx);
                        }
                    }
                }
                if (nodeAddr == kEntrypointSentinel) MR_FINISH();
                if (k) { cur1 = nodeAddr; sp1 = sp; } else { cur0 = nodeAddr; sp0 = sp; }
            }
        }
    }
#undef MR_PUSH
#undef MR_POP
#undef MR_FINISH
}

// ================================================================================================================
// Parked-ray kernels ("sw").  The one-ray kernels run their node loop with 15 and their leaf loop with 5-7 of 32 lanes on secondary
// rays: a lane whose ray holds a leaf idles through the node phase, a lane whose ray sits at a node idles through the leaf phase.
// The mr kernels above fix that by keeping both rays of a lane in shared memory and paying a state load / store on EVERY step (+46 %
// thread instructions, slower overall).  Here the lane's active ray lives in registers exactly as in the one-ray kernels and a second,
// PARKED ray waits in shared memory (80 bytes + its own short stack); the two are exchanged only at the two phase boundaries of the
// while-while loop, in convergent code, and only by lanes that would otherwise sit the coming phase out:
//   before the node loop:  active ray cannot take a node step, parked ray can        -> exchange
//   before the leaf loop:  active ray holds no leaf,           parked ray holds one   -> exchange
// Per-ray traversal is untouched (same node order, same postponed-leaf rule), so results are bit-identical to the one-ray kernel of the
// same node format.  MEASURED AND NOT ADOPTED (scripts/sw_sweep.sh, profiles/r2g_sw_kernels.json): 0.85-0.88x the one-ray kernels on the
// bench frame (binary AO 3 306 vs 3 908, Wide4 diffuse 2 197 vs 2 499 Mrays/s); 55 registers + 14 KB of shared memory leave 9 CTAs per SM
// with 93 KB of L1, and fewer CTAs are worse still (6: -10 %, 4: -28 %) -- the kernel lives on resident warps hiding L1/L2 latency, and a
// second ray per lane does not replace them.  Kept under the b200_sw* names as the record of the experiment.  Parked state: f0 (orig.xyz, tmin)  f1 (dir.xyz, hitT)  f2 (hitIndex, hitU, hitV, rayidx)  f3 (nodeAddr, leafAddr, sp, -)
// f4 (idir.xyz, -).
// ================================================================================================================
constexpr int kSwLocal = 120;                  // stack entries beyond the shared-memory part, per ray (local memory, rarely touched)

template <int BLOCK, int SMEM_N, bool FAST, int FMT>
__global__ void __launch_bounds__(BLOCK)
trace_sw_kernel(int numRays, int anyHit, int fetchThreshold, int nodeExit, unsigned oneBits,
                const float4* __restrict__ rays, int4* __restrict__ results,
                const float4* __restrict__ nodes, const float4* __restrict__ woop,
                const int* __restrict__ triIndices, int* __restrict__ warpCounter, int* __restrict__ errorFlag)
{
    extern __shared__ float4 s_dyn[];
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;
    float4* const P = s_dyn + tid;                                              // parked field f: P[f * BLOCK]
    int* const S = reinterpret_cast<int*>(s_dyn + 5 * BLOCK) + tid;             // stack entry e of slot k: S[(k * SMEM_N + e) * BLOCK]
    int l_stack[2][kSwLocal];
    const unsigned one = oneBits;

    int   rayidx = -1;
    float origx = 0, origy = 0, origz = 0, dirx = 0, diry = 0, dirz = 0, tmin = 0;
    float idirx = 0, idiry = 0, idirz = 0, oodx = 0, oody = 0, oodz = 0;
    int   sp = 0;
    int   leafAddr = 0;
    int   nodeAddr = kEntrypointSentinel;
    int   hitIndex = -1;
    float hitT = 0, hitU = 0, hitV = 0;
    int   slot = 0;                          // which of the lane's two stacks the active ray uses
    int*  sbase = S;
    unsigned pflags = 0;                     // parked ray: 1 = present, 2 = can take a node step, 4 = holds a leaf
    bool  more = true;                       // warp-uniform: the global counter still has rays

#define SW_PUSH(v)  do { ++sp; if (sp < SMEM_N) sbase[sp * BLOCK] = (v); else if (sp < SMEM_N + kSwLocal) l_stack[slot][sp - SMEM_N] = (v); else { --sp; *(volatile int*)errorFlag = 1; } } while (0)
#define SW_POP(dst) do { (dst) = (sp < SMEM_N) ? sbase[sp * BLOCK] : l_stack[slot][sp - SMEM_N]; --sp; } while (0)
    // active <-> parked, one float4 at a time (four temporaries); an absent parked ray leaves the active slot empty
#define SW_EXCHANGE()                                                                                                        \
    do {                                                                                                                     \
        const bool outValid = rayidx >= 0, inValid = (pflags & 1u) != 0u;                                                    \
        const unsigned nf = outValid ? (1u | (((unsigned)nodeAddr < (unsigned)kEntrypointSentinel) ? 2u : 0u) | ((leafAddr < 0) ? 4u : 0u)) : 0u; \
        float4 t;                                                                                                            \
        t = P[0];         if (outValid) P[0] = make_float4(origx, origy, origz, tmin);                                        \
        origx = t.x; origy = t.y; origz = t.z; tmin = t.w;                                                                    \
        t = P[BLOCK];     if (outValid) P[BLOCK] = make_float4(dirx, diry, dirz, hitT);                                       \
        dirx = t.x; diry = t.y; dirz = t.z; hitT = t.w;                                                                       \
        t = P[2 * BLOCK]; if (outValid) P[2 * BLOCK] = make_float4(__int_as_float(hitIndex), hitU, hitV, __int_as_float(rayidx)); \
        hitIndex = __float_as_int(t.x); hitU = t.y; hitV = t.z; rayidx = inValid ? __float_as_int(t.w) : -1;                  \
        t = P[3 * BLOCK]; if (outValid) P[3 * BLOCK] = make_float4(__int_as_float(nodeAddr), __int_as_float(leafAddr), __int_as_float(sp), 0.0f); \
        nodeAddr = inValid ? __float_as_int(t.x) : kEntrypointSentinel; leafAddr = inValid ? __float_as_int(t.y) : 0; sp = __float_as_int(t.z); \
        t = P[4 * BLOCK]; if (outValid) P[4 * BLOCK] = make_float4(idirx, idiry, idirz, 0.0f);                                \
        idirx = t.x; idiry = t.y; idirz = t.z;                                                                                \
        oodx = origx * idirx; oody = origy * idiry; oodz = origz * idirz;                                                     \
        pflags = nf; slot ^= 1; sbase = S + slot * (SMEM_N * BLOCK);                                                          \
    } while (0)
#define SW_FINISH()                                                                                                          \
    do {                                                                                                                     \
        if (rayidx >= 0 && nodeAddr == kEntrypointSentinel && leafAddr >= 0) {                                               \
            int id = hitIndex;                                                                                               \
            if (id != -1) id = __ldg(triIndices + id);                                                                       \
            __stcs(results + rayidx, make_int4(id, __float_as_int(hitT), __float_as_int(hitU), __float_as_int(hitV)));       \
            rayidx = -1;                                                                                                     \
        }                                                                                                                    \
    } while (0)

    for (;;) {
        // ---------------- fetch: a lane takes a new ray when one of its two slots is free ----------------
        {
            const bool want = more && (rayidx < 0 || !(pflags & 1u));
            const unsigned wantMask = __ballot_sync(0xffffffffu, want);
            const unsigned workMask = __ballot_sync(0xffffffffu, rayidx >= 0 || (pflags & 1u));
            if (wantMask && (__popc(wantMask) >= fetchThreshold || workMask == 0u)) {
                if (want && rayidx >= 0) SW_EXCHANGE();            // park the active ray (the parked slot is free): the new ray goes to the registers
                const int n = __popc(wantMask);
                const int rank = __popc(wantMask & ((1u << lane) - 1u));
                const int leader = __ffs(wantMask) - 1;
                int base = 0;
                if ((int)lane == leader) base = atomicAdd(warpCounter, n);
                base = __shfl_sync(0xffffffffu, base, leader);
                if (base + n >= numRays) more = false;
                const int r = base + rank;
                if (want && r < numRays) {
                    float4 o, d;
                    wld256_cs(rays + r * 2, o, d);
                    rayidx = r;
                    origx = o.x; origy = o.y; origz = o.z; tmin = o.w;
                    dirx = d.x; diry = d.y; dirz = d.z; hitT = d.w;
                    const float ooeps = exp2f(-80.0f);                      // fermi_speculative_while_while.cu:94-98
                    idirx = 1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
                    idiry = 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
                    idirz = 1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
                    oodx = origx * idirx; oody = origy * idiry; oodz = origz * idirz;
                    sp = 0;
                    if (SMEM_N > 0) sbase[0] = kEntrypointSentinel; else l_stack[slot][0] = kEntrypointSentinel;
                    leafAddr = 0;
                    nodeAddr = 0;
                    hitIndex = -1;
                    hitU = 0.0f; hitV = 0.0f;
                }
            } else if (workMask == 0u) {
                break;
            }
        }

        // ---------------- node phase ----------------
        {
            const bool sw = !((unsigned)nodeAddr < (unsigned)kEntrypointSentinel && rayidx >= 0) && (pflags & 2u);
            if (__any_sync(0xffffffffu, sw)) { if (sw) SW_EXCHANGE(); }
        }
        while ((unsigned)nodeAddr < (unsigned)kEntrypointSentinel) {
            if (FMT == kFmtWide4) {
                const float4* ptr = nodes + (size_t)nodeAddr * 4;
                float4 A0, A1, B0, B1;
                wld256_nc(ptr, A0, A1);
                wld256_nc(ptr + 2, B0, B1);
                const float ax = A0.w * idirx, ay = A1.x * idiry, az = A1.y * idirz;
                const float bx = fmaf(A0.x, idirx, -oodx) - ax;
                const float by = fmaf(A0.y, idiry, -oody) - ay;
                const float bz = fmaf(A0.z, idirz, -oodz) - az;
                const bool ngx = idirx < 0.0f, ngy = idiry < 0.0f, ngz = idirz < 0.0f;
                const unsigned qlx = __float_as_uint(A1.z), qly = __float_as_uint(A1.w), qlz = __float_as_uint(B0.x);
                const unsigned qhx = __float_as_uint(B0.y), qhy = __float_as_uint(B0.z), qhz = __float_as_uint(B0.w);
                const unsigned nx = ngx ? qhx : qlx, fx = ngx ? qlx : qhx;
                const unsigned ny = ngy ? qhy : qly, fy = ngy ? qly : qhy;
                const unsigned nz = ngz ? qhz : qlz, fz = ngz ? qlz : qhz;
                int k0, k1, k2, k3;
#define NT_CHILD(I, K)                                                                                                    \
                {                                                                                                         \
                    const float tn = wfmax3(fmaf(qplane<I>(nx, one), ax, bx), fmaf(qplane<I>(ny, one), ay, by),           \
                                            fmaxf(fmaf(qplane<I>(nz, one), az, bz), tmin));                               \
                    const float tf = wfmin3(fmaf(qplane<I>(fx, one), ax, bx), fmaf(qplane<I>(fy, one), ay, by),           \
                                            fminf(fmaf(qplane<I>(fz, one), az, bz), hitT));                               \
                    K = (tn <= tf) ? ((__float_as_int(tn) & ~3) | I) : 0x7fffffff;                                        \
                }
                NT_CHILD(0, k0) NT_CHILD(1, k1) NT_CHILD(2, k2) NT_CHILD(3, k3)
#undef NT_CHILD
                cas(k0, k1); cas(k2, k3); cas(k0, k2); cas(k1, k3); cas(k1, k2);
                const int l0 = __float_as_int(B1.x), l1 = __float_as_int(B1.y), l2 = __float_as_int(B1.z), l3 = __float_as_int(B1.w);
#define NT_LINK(K) (((K) & 2) ? (((K) & 1) ? l3 : l2) : (((K) & 1) ? l1 : l0))
                if (k0 == 0x7fffffff) {
                    SW_POP(nodeAddr);
                } else {
                    nodeAddr = NT_LINK(k0);
                    if (k1 != 0x7fffffff) {
                        if (k2 != 0x7fffffff) {
                            if (k3 != 0x7fffffff) SW_PUSH(NT_LINK(k3));
                            SW_PUSH(NT_LINK(k2));
                        }
                        SW_PUSH(NT_LINK(k1));
                    }
                }
#undef NT_LINK
            } else {
                // binary Compact / Compact2 node (CudaBVH.hpp:43-47), slab test as nt_trace.cu
                const float4* ptr = (FMT == kFmtCompact2) ? nodes + nodeAddr
                                                          : reinterpret_cast<const float4*>(reinterpret_cast<const char*>(nodes) + nodeAddr);
                float4 n0xy, n1xy, nz, cn;
                wld256_nc(ptr, n0xy, n1xy);
                wld256_nc(ptr + 2, nz, cn);
                int c0idx = __float_as_int(cn.x), c1idx = __float_as_int(cn.y);
                const float c0lox = n0xy.x * idirx - oodx, c0hix = n0xy.y * idirx - oodx;
                const float c0loy = n0xy.z * idiry - oody, c0hiy = n0xy.w * idiry - oody;
                const float c0loz = nz.x * idirz - oodz,   c0hiz = nz.y * idirz - oodz;
                const float c1loz = nz.z * idirz - oodz,   c1hiz = nz.w * idirz - oodz;
                const float c0min = wfmax3(fminf(c0lox, c0hix), fminf(c0loy, c0hiy), fmaxf(fminf(c0loz, c0hiz), tmin));
                const float c0max = wfmin3(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy), fminf(fmaxf(c0loz, c0hiz), hitT));
                const float c1lox = n1xy.x * idirx - oodx, c1hix = n1xy.y * idirx - oodx;
                const float c1loy = n1xy.z * idiry - oody, c1hiy = n1xy.w * idiry - oody;
                const float c1min = wfmax3(fminf(c1lox, c1hix), fminf(c1loy, c1hiy), fmaxf(fminf(c1loz, c1hiz), tmin));
                const float c1max = wfmin3(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy), fminf(fmaxf(c1loz, c1hiz), hitT));
                const bool trav0 = (c0max >= c0min), trav1 = (c1max >= c1min);
                if (!trav0 && !trav1) {
                    SW_POP(nodeAddr);
                } else {
                    nodeAddr = trav0 ? c0idx : c1idx;
                    if (trav0 && trav1) {
                        if (c1min < c0min) { const int t = nodeAddr; nodeAddr = c1idx; c1idx = t; }
                        SW_PUSH(c1idx);
                    }
                }
            }
            // first leaf => postpone and continue traversal (speculative while-while, as nt_trace.cu)
            if (nodeAddr < 0 && leafAddr >= 0) {
                leafAddr = nodeAddr;
                SW_POP(nodeAddr);
            }
            {
                const unsigned am = __activemask();
                if (!__any_sync(am, leafAddr >= 0) || __popc(am) < nodeExit) break;
            }
        }
        SW_FINISH();

        // ---------------- leaf phase: the reference's Woop test (Util.cpp:99-127), identical to nt_trace.cu ----------------
        {
            const bool sw = !(leafAddr < 0 && rayidx >= 0) && (pflags & 4u);
            if (__any_sync(0xffffffffu, sw)) { if (sw) SW_EXCHANGE(); }
        }
        while (leafAddr < 0) {
            int triAddr = ~leafAddr;
            float4 v00 = __ldg(woop + triAddr);
            for (;;) {
                if (__float_as_int(v00.x) == (int)0x80000000) break;
                float t;
                if (FAST) {
                    const float Oz = v00.w - origx * v00.x - origy * v00.y - origz * v00.z;
                    t = Oz * __fdividef(1.0f, dirx * v00.x + diry * v00.y + dirz * v00.z);
                } else {
                    const float Oz = __fsub_rn(__fsub_rn(__fsub_rn(v00.w, __fmul_rn(origx, v00.x)), __fmul_rn(origy, v00.y)), __fmul_rn(origz, v00.z));
                    const float dd = __fadd_rn(__fadd_rn(__fmul_rn(dirx, v00.x), __fmul_rn(diry, v00.y)), __fmul_rn(dirz, v00.z));
                    t = __fmul_rn(Oz, __frcp_rn(dd));
                }
                if (t > tmin && t < hitT) {
                    const float4 v11 = __ldg(woop + triAddr + 1);
                    float u;
                    if (FAST) u = (v11.w + origx * v11.x + origy * v11.y + origz * v11.z) + t * (dirx * v11.x + diry * v11.y + dirz * v11.z);
                    else {
                        const float Ou = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v11.x, origx), __fmul_rn(v11.y, origy)), __fmul_rn(v11.z, origz)), v11.w);
                        const float Du = __fadd_rn(__fadd_rn(__fmul_rn(v11.x, dirx), __fmul_rn(v11.y, diry)), __fmul_rn(v11.z, dirz));
                        u = __fadd_rn(Ou, __fmul_rn(t, Du));
                    }
                    if (u >= 0.0f) {
                        const float4 v22 = __ldg(woop + triAddr + 2);
                        float v;
                        if (FAST) v = (v22.w + origx * v22.x + origy * v22.y + origz * v22.z) + t * (dirx * v22.x + diry * v22.y + dirz * v22.z);
                        else {
                            const float Ov = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v22.x, origx), __fmul_rn(v22.y, origy)), __fmul_rn(v22.z, origz)), v22.w);
                            const float Dv = __fadd_rn(__fadd_rn(__fmul_rn(v22.x, dirx), __fmul_rn(v22.y, diry)), __fmul_rn(v22.z, dirz));
                            v = __fadd_rn(Ov, __fmul_rn(t, Dv));
                        }
                        if (v >= 0.0f && (FAST ? (u + v) : __fadd_rn(u, v)) <= 1.0f) {
                            hitT = t; hitU = u; hitV = v;
                            hitIndex = triAddr;
                            if (anyHit) { nodeAddr = kEntrypointSentinel; break; }
                        }
                    }
                }
                triAddr += 3;
                v00 = __ldg(woop + triAddr);
            }
            leafAddr = nodeAddr;
            if (nodeAddr < 0) SW_POP(nodeAddr);
        }
        SW_FINISH();
    }
#undef SW_PUSH
#undef SW_POP
#undef SW_EXCHANGE
#undef SW_FINISH
}

struct SwTuning { int smemStack; int ctasPerSM; int fetchThreshold; int nodeExit; };
SwTuning sw_tuning()
{
    static SwTuning t = [] {
        SwTuning r{8, 0, 8, 8};
        if (const char* e = getenv("NT_SW_SMEM")) r.smemStack = atoi(e);
        if (const char* e = getenv("NT_SW_CTAS")) r.ctasPerSM = atoi(e);          // 0 = as many as fit
        if (const char* e = getenv("NT_SW_FETCH")) r.fetchThreshold = atoi(e);
        if (const char* e = getenv("NT_SW_NODE_EXIT")) r.nodeExit = atoi(e);
        return r;
    }();
    return t;
}

template <int SMEM_N, bool FAST, int FMT>
cudaError_t launch_sw_variant(const TraceLaunch& a, int* launches)
{
    auto kern = trace_sw_kernel<kWideBlock, SMEM_N, FAST, FMT>;
    constexpr int smemBytes = kWideBlock * (5 * 16 + 2 * SMEM_N * 4);
    static int blocksPerSM = 0, epoch = -1;
    if (epoch != launch_epoch()) {
        epoch = launch_epoch();
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
        if (e != cudaSuccess) return e;
        const int want = sw_tuning().ctasPerSM;
        int carve = 100;
        if (want > 0) { carve = (want * (smemBytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024); if (carve > 100) carve = 100; }
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kern, kWideBlock, smemBytes);
        if (e != cudaSuccess) return e;
        if (want > 0 && blocksPerSM > want) blocksPerSM = want;
        if (blocksPerSM < 1) blocksPerSM = 1;
    }
    int grid = (a.numRays + 2 * kWideBlock - 1) / (2 * kWideBlock);
    if (grid > a.numSMs * blocksPerSM) grid = a.numSMs * blocksPerSM;
    const SwTuning t = sw_tuning();
    kern<<<grid, kWideBlock, smemBytes, a.stream>>>(a.numRays, a.anyHit, t.fetchThreshold, t.nodeExit, 0x3F800000u, a.rays, a.results,
                                                     FMT == kFmtWide4 ? a.wideNodes : a.nodes, a.woop, a.triIndices, a.warpCounter, a.errorFlag);
    if (launches) *launches = 1;
    return cudaGetLastError();
}

template <bool FAST, int FMT>
cudaError_t launch_sw_stack(const TraceLaunch& a, int* launches)
{
    switch (sw_tuning().smemStack) {
    case 4:  return launch_sw_variant<4, FAST, FMT>(a, launches);
    case 12: return launch_sw_variant<12, FAST, FMT>(a, launches);
    default: return launch_sw_variant<8, FAST, FMT>(a, launches);
    }
}

struct MrTuning { int smemStack; int ctasPerSM; int fetchThreshold; int leafThreshold; };
MrTuning mr_tuning()
{
    static MrTuning t = [] {
        MrTuning r{8, 0, 12, 16};
        if (const char* e = getenv("NT_MR_SMEM")) r.smemStack = atoi(e);
        if (const char* e = getenv("NT_MR_CTAS")) r.ctasPerSM = atoi(e);          // 0 = as many as fit
        if (const char* e = getenv("NT_MR_FETCH")) r.fetchThreshold = atoi(e);
        if (const char* e = getenv("NT_MR_LEAF")) r.leafThreshold = atoi(e);
        return r;
    }();
    return t;
}

template <int SMEM_N, bool FAST, int FMT>
cudaError_t launch_mr_variant(const TraceLaunch& a, int* launches)
{
    auto kern = trace_mr_kernel<kWideBlock, SMEM_N, FAST, FMT>;
    constexpr int smemBytes = kWideBlock * (8 * 16 + 2 * SMEM_N * 4);
    static int blocksPerSM = 0, epoch = -1;
    if (epoch != launch_epoch()) {
        epoch = launch_epoch();
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
        if (e != cudaSuccess) return e;
        int want = mr_tuning().ctasPerSM;
        // shared memory the wanted CTAs need (+1 KB per CTA the hardware reserves); what is left of the 228 KB stays L1
        int carve = 100;
        if (want > 0) { carve = (want * (smemBytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024); if (carve > 100) carve = 100; }
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kern, kWideBlock, smemBytes);
        if (e != cudaSuccess) return e;
        if (want > 0 && blocksPerSM > want) blocksPerSM = want;
        if (blocksPerSM < 1) blocksPerSM = 1;
    }
    int grid = (a.numRays + 2 * kWideBlock - 1) / (2 * kWideBlock);
    if (grid > a.numSMs * blocksPerSM) grid = a.numSMs * blocksPerSM;
    const MrTuning t = mr_tuning();
    kern<<<grid, kWideBlock, smemBytes, a.stream>>>(a.numRays, a.anyHit, t.fetchThreshold, t.leafThreshold, 0x3F800000u, a.rays, a.results,
                                                     FMT == kFmtWide4 ? a.wideNodes : a.nodes, a.woop, a.triIndices, a.warpCounter, a.errorFlag);
    if (launches) *launches = 1;
    return cudaGetLastError();
}

template <bool FAST, int FMT>
cudaError_t launch_mr_stack(const TraceLaunch& a, int* launches)
{
    switch (mr_tuning().smemStack) {
    case 4:  return launch_mr_variant<4, FAST, FMT>(a, launches);
    case 12: return launch_mr_variant<12, FAST, FMT>(a, launches);
    default: return launch_mr_variant<8, FAST, FMT>(a, launches);
    }
}

struct WideTuning { int smemStack; int carveout; int fetchThreshold; int nodeExit; };
WideTuning wide_tuning()
{
    static WideTuning t = [] {
        WideTuning r{8, 20, 20, 8};
        if (const char* e = getenv("NT_WIDE_NODE_EXIT")) r.nodeExit = atoi(e);
        if (const char* e = getenv("NT_WIDE_SMEM")) r.smemStack = atoi(e);
        if (const char* e = getenv("NT_WIDE_CARVEOUT")) r.carveout = atoi(e);
        if (const char* e = getenv("NT_WIDE_FETCH")) r.fetchThreshold = atoi(e);
        return r;
    }();
    return t;
}

template <int SMEM_N, bool FAST, bool WIDE_RAYS>
cudaError_t launch_wide_variant(const TraceLaunch& a, int* launches)
{
    auto kern = trace_wide4_kernel<kWideBlock, SMEM_N, FAST, WIDE_RAYS>;
    static int blocksPerSM = 0, epoch = -1;
    if (epoch != launch_epoch()) {
        epoch = launch_epoch();
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, wide_tuning().carveout);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kern, kWideBlock, 0);
        if (e != cudaSuccess) return e;
        if (blocksPerSM < 1) blocksPerSM = 1;
    }
    int grid = (a.numRays + kWideBlock - 1) / kWideBlock;
    if (grid > a.numSMs * blocksPerSM) grid = a.numSMs * blocksPerSM;
    kern<<<grid, kWideBlock, 0, a.stream>>>(a.numRays, a.anyHit, wide_tuning().fetchThreshold, wide_tuning().nodeExit, 0x3F800000u, a.rays, a.results, a.wideNodes, a.woop, a.triIndices, a.warpCounter, a.errorFlag, a.batches);
    if (launches) *launches = 1;
    return cudaGetLastError();
}

template <bool FAST, bool WIDE_RAYS>
cudaError_t launch_wide_stack(const TraceLaunch& a, int* launches)
{
    switch (wide_tuning().smemStack) {
    case 4:  return launch_wide_variant<4, FAST, WIDE_RAYS>(a, launches);
    case 16: return launch_wide_variant<16, FAST, WIDE_RAYS>(a, launches);
    default: return launch_wide_variant<8, FAST, WIDE_RAYS>(a, launches);
    }
}

} // namespace

cudaError_t launch_trace_mr(const TraceLaunch& a, int* launches)
{
    if (a.numRays <= 0) { if (launches) *launches = 0; return cudaSuccess; }
    // 256-bit loads: 32-byte aligned rays, 64-byte aligned nodes (true for everything the library allocates)
    if ((reinterpret_cast<size_t>(a.rays) & 31) || (reinterpret_cast<size_t>(a.nodes) & 63)) return cudaErrorInvalidValue;
    if (a.kernel == Kernel_Wide4Mr) {
        if (!a.wideNodes || (reinterpret_cast<size_t>(a.wideNodes) & 63)) return cudaErrorInvalidValue;
        return a.fast ? launch_mr_stack<true, kFmtWide4>(a, launches) : launch_mr_stack<false, kFmtWide4>(a, launches);
    }
    if (a.layout == Layout_Compact2) return a.fast ? launch_mr_stack<true, kFmtCompact2>(a, launches) : launch_mr_stack<false, kFmtCompact2>(a, launches);
    return a.fast ? launch_mr_stack<true, kFmtCompact>(a, launches) : launch_mr_stack<false, kFmtCompact>(a, launches);
}

cudaError_t launch_trace_sw(const TraceLaunch& a, int* launches)
{
    if (a.numRays <= 0) { if (launches) *launches = 0; return cudaSuccess; }
    if ((reinterpret_cast<size_t>(a.rays) & 31) || (reinterpret_cast<size_t>(a.nodes) & 63)) return cudaErrorInvalidValue;
    if (a.kernel == Kernel_Wide4Sw) {
        if (!a.wideNodes || (reinterpret_cast<size_t>(a.wideNodes) & 63)) return cudaErrorInvalidValue;
        return a.fast ? launch_sw_stack<true, kFmtWide4>(a, launches) : launch_sw_stack<false, kFmtWide4>(a, launches);
    }
    if (a.layout == Layout_Compact2) return a.fast ? launch_sw_stack<true, kFmtCompact2>(a, launches) : launch_sw_stack<false, kFmtCompact2>(a, launches);
    return a.fast ? launch_sw_stack<true, kFmtCompact>(a, launches) : launch_sw_stack<false, kFmtCompact>(a, launches);
}

cudaError_t launch_trace_wide4(const TraceLaunch& a, int* launches)
{
    if (a.numRays <= 0) { if (launches) *launches = 0; return cudaSuccess; }
    if (!a.wideNodes || (reinterpret_cast<size_t>(a.wideNodes) & 63)) return cudaErrorInvalidValue;
    const bool wideRays = a.batches || (reinterpret_cast<size_t>(a.rays) & 31) == 0;       // the buffers of a batch table are checked by nt_trace_batches
    if (a.fast) return wideRays ? launch_wide_stack<true, true>(a, launches) : launch_wide_stack<true, false>(a, launches);
    return wideRays ? launch_wide_stack<false, true>(a, launches) : launch_wide_stack<false, false>(a, launches);
}

} // namespace nt
