// ntrace_b200 — the reference's "basic" CudaBVH layouts (AOS_AOS, AOS_SOA, SOA_AOS, SOA_SOA) on the device.
//
// Reference: src/rt/cuda/CudaBVH.cpp:453-575 (createNodeBasic / createTriWoopBasic / createTriIndexBasic) and
// CudaBVH.hpp:60-85.  In those layouts every BVH node, leaves included, is a 64-byte record
//   int4 x4: (c0.lo.x,c0.hi.x,c0.lo.y,c0.hi.y) (c1...) (c0.lo.z,c0.hi.z,c1.lo.z,c1.hi.z) (c0, c1, splitBits, 0)
// where an inner node's c0/c1 are child node indices (>= 0 inner, ~index leaf) and a leaf node's c0/c1 are the
// [lo,hi) range of its triangles in triIndex order; triangles are 3 Woop float4 in a 64-byte slot.  "SOA" stores the
// four (three) 16-byte planes in separate quarters of the buffer.  Buffers are padded to 4096 bytes, the padding is
// uninitialised, so the node count is only known by walking the tree.
//
// The tesla_* kernels that consume these layouts are not re-implemented; a BVH uploaded in one of them is rewritten
// ON THE GPU into the Compact form (implicit leaves, terminator-delimited triangle lists, CudaBVH.cpp:579-687) that
// the one B200 traversal kernel reads.  Same tree, same boxes, same Woop data: results are those of the Compact BVH.
//   1. level-synchronous walk from node 0: marks inner nodes, counts leaf starts, validates indices / ranges
//   2. two exclusive scans: inner-node numbering (old index order, root stays 0) and leaf rank by range start
//   3. one kernel rewrites the inner nodes, one the triangles (+ terminators, + the -0.0f guard of CudaBVH.cpp:627)
#include "nt_common.cuh"
#include "nt_sort.cuh"

namespace nt {
namespace {

struct BasicView {
    const int4* nodes; int nodeMul, nodePlane;        // plane j of node i at nodes[i * nodeMul + j * nodePlane]
    const float4* woop; int triMul, triPlane;
    const int* triIndex;
    int numSlots, numRefs;
};

enum : unsigned { kErrChildRange = 1u, kErrNodeTwice = 2u, kErrLeafRange = 4u, kErrLeafOverlap = 8u, kErrCoverage = 16u };

struct WalkState { unsigned err; unsigned nextCount; unsigned emptyLeaves; unsigned long long covered; };

__device__ __forceinline__ int4 node_plane(const BasicView& v, int i, int j) { return __ldg(v.nodes + (size_t)i * v.nodeMul + (size_t)j * v.nodePlane); }

__global__ void walk_level_kernel(BasicView v, const int* __restrict__ frontier, int n, int* __restrict__ next,
                                  unsigned* __restrict__ innerFlag, unsigned* __restrict__ leafStart, int* __restrict__ leafEnd, WalkState* st)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int node = frontier[t];
    if (atomicExch(&innerFlag[node], 1u) != 0u) { atomicOr(&st->err, kErrNodeTwice); return; }
    const int4 link = node_plane(v, node, 3);
    const int child[2] = {link.x, link.y};
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int c = child[k];
        if (c >= 0) {
            if (c >= v.numSlots || c == 0) { atomicOr(&st->err, kErrChildRange); continue; }
            next[atomicAdd(&st->nextCount, 1u)] = c;
        } else {
            const int leaf = ~c;
            if (leaf >= v.numSlots) { atomicOr(&st->err, kErrChildRange); continue; }
            const int4 range = node_plane(v, leaf, 3);
            const int lo = range.x, hi = range.y;
            if (lo < 0 || hi < lo || hi > v.numRefs) { atomicOr(&st->err, kErrLeafRange); continue; }
            if (lo == hi) { atomicAdd(&st->emptyLeaves, 1u); continue; }
            if (atomicAdd(&leafStart[lo], 1u) != 0u) atomicOr(&st->err, kErrLeafOverlap);
            leafEnd[lo] = hi;
            atomicAdd(&st->covered, (unsigned long long)(hi - lo));
        }
    }
}

// every leaf must end where another begins (or at numRefs): with distinct starts, a start at 0 and total size numRefs
// this makes the leaves a partition of [0, numRefs)
__global__ void check_partition_kernel(int numRefs, const unsigned* __restrict__ leafStart, const int* __restrict__ leafEnd, WalkState* st)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numRefs) return;
    if (i == 0 && leafStart[0] == 0u) atomicOr(&st->err, kErrCoverage);
    if (leafStart[i] != 0u) {
        const int hi = leafEnd[i];
        if (hi != numRefs && leafStart[hi] == 0u) atomicOr(&st->err, kErrCoverage);
    }
}

__global__ void emit_nodes_kernel(BasicView v, const unsigned* __restrict__ innerFlag, const unsigned* __restrict__ innerRank,
                                  const unsigned* __restrict__ leafRank, int emptyLeafAddr, int ofsDiv, int4* __restrict__ out)
{
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= v.numSlots || innerFlag[node] == 0u) return;
    int4 link = node_plane(v, node, 3);
    int child[2] = {link.x, link.y};
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int c = child[k];
        if (c >= 0) child[k] = (int)innerRank[c] * (64 / ofsDiv);          // Compact: byte offset; Compact2: offset / 16
        else {
            const int4 range = node_plane(v, ~c, 3);
            child[k] = (range.x == range.y) ? ~emptyLeafAddr : ~(range.x * 3 + (int)leafRank[range.x]);
        }
    }
    int4* dst = out + (size_t)innerRank[node] * 4;
    dst[0] = node_plane(v, node, 0);
    dst[1] = node_plane(v, node, 1);
    dst[2] = node_plane(v, node, 2);
    dst[3] = make_int4(child[0], child[1], link.z, 0);
}

__global__ void emit_tris_kernel(BasicView v, const unsigned* __restrict__ leafStart, const unsigned* __restrict__ leafRank,
                                 int spareTerminatorAt, float4* __restrict__ outWoop, int* __restrict__ outIndex)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float4 term = make_float4(__int_as_float(0x80000000), __int_as_float(0x80000000), __int_as_float(0x80000000), __int_as_float(0x80000000));
    if (i == 0 && spareTerminatorAt >= 0) { outWoop[spareTerminatorAt] = term; outIndex[spareTerminatorAt] = 0; }
    if (i >= v.numRefs) return;
    const int rank = (int)(leafRank[i] + leafStart[i]) - 1;                 // leaves that start at or before i, minus one
    const size_t base = (size_t)i * 3 + rank;
    float4 w0 = __ldg(v.woop + (size_t)i * v.triMul);
    if (w0.x == 0.0f) w0.x = 0.0f;                                          // -0.0f would read as the terminator (CudaBVH.cpp:627-628)
    outWoop[base] = w0;
    outWoop[base + 1] = __ldg(v.woop + (size_t)i * v.triMul + v.triPlane);
    outWoop[base + 2] = __ldg(v.woop + (size_t)i * v.triMul + 2 * (size_t)v.triPlane);
    outIndex[base] = __ldg(v.triIndex + i);
    outIndex[base + 1] = 0;
    outIndex[base + 2] = 0;
    if (i == v.numRefs - 1 || leafStart[i + 1] != 0u) { outWoop[base + 3] = term; outIndex[base + 3] = 0; }
}

// Compact <-> Compact2: inner-child links are byte offsets resp. byte offsets / 16 (CudaBVH.cpp:86,614)
__global__ void rescale_links_kernel(int4* nodes, size_t numNodes, int num, int den)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) return;
    int4 link = nodes[i * 4 + 3];
    if (link.x >= 0) link.x = link.x / den * num;
    if (link.y >= 0) link.y = link.y / den * num;
    nodes[i * 4 + 3] = link;
}

// SAH of a flat Compact / Compact2 tree with the formula the reference reports (BVHNode::computeSubtreeProbabilities, BVHNode.cpp:79-94
// through BVH.cpp:67-70; Platform costs Cn = Ct = 1): sum over nodes of P(n) * (Cn * #children + Ct * #triangles) with
// P(child) = P(parent) * A(child) / A(parent), P(root) = 1.  The product telescopes to A(n) / A(root), so the sum is parallel: one thread
// per inner node adds, for each of its two child boxes, 2 * A (inner child) or #triangles * A (leaf child, triangles counted up to the
// terminator); the root contributes 2.  Replaces the reference's single-thread calcSAH kernel (emitTreeKernel.cu:1361-1400), which uses
// inner cost 1 (do not mix the two).  acc = {sum of child terms, A(root), leaves, triangles} in double / u64.
__global__ void __launch_bounds__(256) sah_kernel(const float4* __restrict__ nodes, size_t numNodes, const float4* __restrict__ woop, size_t woopRows,
                                                   double* __restrict__ acc, unsigned long long* __restrict__ counts)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double sum = 0.0;
    unsigned leaves = 0, tris = 0;
    if (i < numNodes) {
        const float4 n0 = __ldg(nodes + i * 4), n1 = __ldg(nodes + i * 4 + 1), nz = __ldg(nodes + i * 4 + 2), cn = __ldg(nodes + i * 4 + 3);
        const float dx[2] = {n0.y - n0.x, n1.y - n1.x}, dy[2] = {n0.w - n0.z, n1.w - n1.z}, dz[2] = {nz.y - nz.x, nz.w - nz.z};
        const int link[2] = {__float_as_int(cn.x), __float_as_int(cn.y)};
        for (int c = 0; c < 2; c++) {
            const double area = 2.0 * ((double)dx[c] * dy[c] + (double)dy[c] * dz[c] + (double)dz[c] * dx[c]);     // AABB::area (Util.hpp:45)
            if (link[c] >= 0) sum += 2.0 * area;
            else {
                unsigned n = 0;
                for (size_t a = (size_t)(unsigned)~link[c]; a < woopRows && __float_as_uint(__ldg(woop + a).x) != 0x80000000u; a += 3) n++;
                sum += (double)n * area;
                leaves++; tris += n;
            }
        }
        if (i == 0) {
            const float lo[3] = {fminf(n0.x, n1.x), fminf(n0.z, n1.z), fminf(nz.x, nz.z)}, hi[3] = {fmaxf(n0.y, n1.y), fmaxf(n0.w, n1.w), fmaxf(nz.y, nz.w)};
            const double ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
            acc[1] = 2.0 * (ex * ey + ey * ez + ez * ex);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        leaves += __shfl_xor_sync(0xffffffffu, leaves, o);
        tris += __shfl_xor_sync(0xffffffffu, tris, o);
    }
    if ((threadIdx.x & 31) == 0 && (sum != 0.0 || leaves)) {
        atomicAdd(acc, sum);
        atomicAdd(counts, (unsigned long long)leaves);
        atomicAdd(counts + 1, (unsigned long long)tris);
    }
}

} // namespace

// out4 (device, 32 bytes, zeroed by the caller): {double childTerms, double rootArea, u64 leaves, u64 triangles}
cudaError_t launch_sah(const float4* dNodes, size_t numNodes, const float4* dWoop, size_t woopRows, void* out4, cudaStream_t stream)
{
    if (numNodes == 0) return cudaErrorInvalidValue;
    sah_kernel<<<(unsigned)((numNodes + 255) / 256), 256, 0, stream>>>(dNodes, numNodes, dWoop, woopRows, (double*)out4, (unsigned long long*)out4 + 2);
    return cudaGetLastError();
}

cudaError_t rescale_compact_links(int4* dNodes, size_t numNodes, int mulNum, int mulDen, cudaStream_t stream)
{
    if (numNodes == 0) return cudaSuccess;
    rescale_links_kernel<<<(unsigned)((numNodes + 255) / 256), 256, 0, stream>>>(dNodes, numNodes, mulNum, mulDen);
    return cudaGetLastError();
}

cudaError_t convert_basic_layout(int layout, const void* dNodes, size_t nodeBytes, const void* dWoop, size_t woopBytes,
                                 const int* dTriIndex, size_t idxBytes, int targetLayout, BuildOutput& out, DevBuf& scratch,
                                 cudaStream_t stream, int* outLaunches, std::string* err)
{
    auto fail = [&](const char* m) { if (err) *err = m; return cudaErrorInvalidValue; };
    if (layout < Layout_AOS_AOS || layout > Layout_SOA_SOA) return fail("not a basic CudaBVH layout");
    if (nodeBytes < 64 || nodeBytes % 64) return fail("node buffer must be a non-empty multiple of 64 bytes");
    const long long S = (long long)(nodeBytes / 64), R = (long long)(idxBytes / 4);
    if (R <= 0 || idxBytes % 4) return fail("empty triangle index buffer");
    if (woopBytes % 64 || (long long)(woopBytes / 64) < R) return fail("Woop buffer smaller than 64 bytes per triangle reference");
    if (S > 0x3fffffff || R > 0x1fffffff) return fail("BVH too large for 32-bit Compact offsets");
    const bool nodeSOA = (layout == Layout_SOA_AOS || layout == Layout_SOA_SOA);
    const bool triSOA = (layout == Layout_AOS_SOA || layout == Layout_SOA_SOA);
    BasicView v;
    v.nodes = (const int4*)dNodes; v.nodeMul = nodeSOA ? 1 : 4; v.nodePlane = nodeSOA ? (int)(nodeBytes / 64) : 1;
    v.woop = (const float4*)dWoop; v.triMul = triSOA ? 1 : 4; v.triPlane = triSOA ? (int)(woopBytes / 64) : 1;
    v.triIndex = dTriIndex; v.numSlots = (int)S; v.numRefs = (int)R;

    // scratch: innerFlag[S] innerRank[S] leafStart[R+1] leafRank[R+1] leafEnd[R] frontier A/B [S] blockSums state
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oFlag = 0, oRank = oFlag + align(S * 4), oLs = oRank + align(S * 4), oLr = oLs + align((R + 1) * 4),
                 oLe = oLr + align((R + 1) * 4), oFa = oLe + align(R * 4), oFb = oFa + align(S * 8), oSums = oFb + align(S * 8),      // frontiers: up to 2 pushes per node
                
                 oState = oSums + align(scan_block_sums_bytes(S > R + 1 ? S : R + 1)), total = oState + 256;
    cudaError_t e = scratch.reserve(total);
    if (e != cudaSuccess) return e;
    char* base = (char*)scratch.p;
    unsigned* innerFlag = (unsigned*)(base + oFlag); unsigned* innerRank = (unsigned*)(base + oRank);
    unsigned* leafStart = (unsigned*)(base + oLs); unsigned* leafRank = (unsigned*)(base + oLr);
    int* leafEnd = (int*)(base + oLe); int* fa = (int*)(base + oFa); int* fb = (int*)(base + oFb);
    unsigned* sums = (unsigned*)(base + oSums); WalkState* st = (WalkState*)(base + oState);
    unsigned* totals = (unsigned*)(base + oState + 64);

    int launches = 0;
    if ((e = cudaMemsetAsync(base, 0, oLe, stream)) != cudaSuccess) return e;               // flags, ranks, leaf starts
    if ((e = cudaMemsetAsync(st, 0, 256, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(fa, 0, 4, stream)) != cudaSuccess) return e;                   // frontier = {root 0}
    int n = 1;
    long long visited = 0;
    WalkState hs;
    while (n > 0) {
        visited += n;
        if (visited > S) return fail("malformed BVH: more reachable inner nodes than node slots");
        walk_level_kernel<<<(n + 255) / 256, 256, 0, stream>>>(v, fa, n, fb, innerFlag, leafStart, leafEnd, st);
        launches++;
        if ((e = cudaMemcpyAsync(&hs, st, sizeof(hs), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
        if (hs.err) break;
        n = (int)hs.nextCount;
        if (n > S) return fail("malformed BVH: child list longer than the node buffer");
        if ((e = cudaMemsetAsync(&st->nextCount, 0, 4, stream)) != cudaSuccess) return e;
        int* t = fa; fa = fb; fb = t;
    }
    if (!hs.err) {
        check_partition_kernel<<<(int)((R + 255) / 256), 256, 0, stream>>>((int)R, leafStart, leafEnd, st);
        launches++;
        if ((e = cudaMemcpyAsync(&hs, st, sizeof(hs), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
        if (!hs.err && hs.covered != (unsigned long long)R) hs.err |= kErrCoverage;
    }
    if (hs.err) {
        if (hs.err & kErrChildRange) return fail("malformed BVH: child index outside the node buffer");
        if (hs.err & kErrNodeTwice) return fail("malformed BVH: node referenced by two parents");
        if (hs.err & kErrLeafRange) return fail("malformed BVH: leaf triangle range outside the index buffer");
        if (hs.err & kErrLeafOverlap) return fail("malformed BVH: two leaves share a triangle range start");
        return fail("malformed BVH: leaf ranges do not partition the triangle index buffer");
    }
    if ((e = exclusive_scan<unsigned>(innerFlag, innerRank, S, sums, totals, stream, &launches)) != cudaSuccess) return e;
    if ((e = exclusive_scan<unsigned>(leafStart, leafRank, R + 1, sums, totals + 1, stream, &launches)) != cudaSuccess) return e;
    unsigned ht[2];
    if ((e = cudaMemcpyAsync(ht, totals, 8, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    const long long numInner = ht[0], numLeaves = ht[1];
    const bool spare = hs.emptyLeaves != 0;
    const long long woopVec = R * 3 + numLeaves + (spare ? 1 : 0);
    if (woopVec > 0x7fffffffLL || numInner * 64 > 0x7fffffffLL) return fail("BVH too large for 32-bit Compact offsets");
    out.nodeBytes = (size_t)numInner * 64; out.woopBytes = (size_t)woopVec * 16; out.idxBytes = (size_t)woopVec * 4;
    if ((e = out.nodes->reserve(out.nodeBytes)) != cudaSuccess) return e;
    if ((e = out.woop->reserve(out.woopBytes)) != cudaSuccess) return e;
    if ((e = out.triIndex->reserve(out.idxBytes)) != cudaSuccess) return e;
    const int spareAt = spare ? (int)(R * 3 + numLeaves) : -1;
    emit_nodes_kernel<<<(int)((S + 255) / 256), 256, 0, stream>>>(v, innerFlag, innerRank, leafRank, spareAt,
                                                                  targetLayout == Layout_Compact2 ? 16 : 1, out.nodes->as<int4>());
    emit_tris_kernel<<<(int)((R + 255) / 256), 256, 0, stream>>>(v, leafStart, leafRank, spareAt, out.woop->as<float4>(), out.triIndex->as<int>());
    launches += 2;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (outLaunches) *outLaunches += launches;
    return cudaStreamSynchronize(stream);
}

} // namespace nt
