// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp / orc_lbvh.hpp headers).
#include "orc_lbvh.hpp"
#include <chrono>
#include <climits>
#include <numeric>

namespace orc {

// ---- device-semantics helpers -------------------------------------------------------------
// CUDA float->int conversion (cvt.rzi.s32.f32) saturates and maps NaN to 0.
static inline int cvt_sat(float f)
{
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT_MAX;
    if (f <= -2147483648.0f) return INT_MIN;
    return (int)f;
}
// emitTreeKernel.cu:70-72 — the int clamp goes through the float overload
static inline int clamp_via_float(int v, int lo, int hi)
{
    float f = std::fmax((float)lo, std::fmin((float)v, (float)hi));
    return (int)f;
}
// emitTreeKernel.cu:647-653
static inline uint32_t spread(uint32_t n)
{
    n &= 0x3ff;
    n = (n ^ (n << 16)) & 0xff0000ff;
    n = (n ^ (n << 8)) & 0x0300f00f;
    n = (n ^ (n << 4)) & 0x030c30c3;
    return (n ^ (n << 2)) & 0x09249249;
}
static inline V3 fmin3(V3 a, V3 b) { return V3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)); }
static inline V3 fmax3(V3 a, V3 b) { return V3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)); }

// emitTreeKernel.cu:655-691 + HLBVHBuilder.cpp:67-83 (step = (hi - lo) / 1024.0f on the host)
void morton_codes(const Scene& sc, V3 lo, V3 hi, uint32_t* codes)
{
    V3 step = V3((hi.x - lo.x) / 1024.0f, (hi.y - lo.y) / 1024.0f, (hi.z - lo.z) / 1024.0f);
    for (int t = 0; t < sc.numTris; t++) {
        V3 a = sc.v(t, 0), b = sc.v(t, 1), c = sc.v(t, 2);
        V3 tlo = fmin3(a, fmin3(b, c));
        V3 thi = fmax3(a, fmax3(b, c));
        V3 half = V3((thi.x - tlo.x) / 2.0f, (thi.y - tlo.y) / 2.0f, (thi.z - tlo.z) / 2.0f);
        V3 mid = tlo + half;
        V3 q = (mid - lo) / step;
        int x = clamp_via_float(cvt_sat(std::floor(q.x)), 0, 1023);
        int y = clamp_via_float(cvt_sat(std::floor(q.y)), 0, 1023);
        int z = clamp_via_float(cvt_sat(std::floor(q.z)), 0, 1023);
        codes[t] = spread((uint32_t)x) | (spread((uint32_t)y) << 1) | (spread((uint32_t)z) << 2);
    }
}

// radixSort.cu:22-46 — thrust::sort_by_key on (u32, s32): ascending, stable (radix).
void sort_pairs_stable(uint32_t* keys, int32_t* idx, int n)
{
    std::vector<uint32_t> k2(n);
    std::vector<int32_t> i2(n);
    uint32_t* ks = keys; uint32_t* kd = k2.data();
    int32_t* is = idx; int32_t* id = i2.data();
    for (int pass = 0; pass < 4; pass++) {
        size_t hist[257] = {0};
        int shift = pass * 8;
        for (int i = 0; i < n; i++) hist[((ks[i] >> shift) & 255) + 1]++;
        for (int d = 0; d < 256; d++) hist[d + 1] += hist[d];
        for (int i = 0; i < n; i++) { size_t p = hist[(ks[i] >> shift) & 255]++; kd[p] = ks[i]; id[p] = is[i]; }
        std::swap(ks, kd); std::swap(is, id);
    }
    // 4 passes -> data is back in the caller's arrays
}

// emitTreeKernel.cu:574-635.  `1.0/(...)` is a double division rounded to float.
void calc_woop_gpu(V3 v0, V3 v1, V3 v2, float out[12])
{
    V3 c0 = v0 - v2, c1 = v1 - v2, c2 = cross(c0, c1);
    float dexp = c0.x * (c2.z * c1.y - c1.z * c2.y) - c0.y * (c2.z * c1.x - c1.z * c2.x) + c0.z * (c2.y * c1.x - c1.y * c2.x);
    float det = (float)(1.0 / (double)dexp);
    V3 i0, i1, i2;
    i0.x =  (c2.z * c1.y - c1.z * c2.y) * det;
    i0.y = -(c2.z * c1.x - c1.z * c2.x) * det;
    i0.z =  (c2.y * c1.x - c1.y * c2.x) * det;
    i1.x = -(c2.z * c0.y - c0.z * c2.y) * det;
    i1.y =  (c2.z * c0.x - c0.z * c2.x) * det;
    i1.z = -(c2.y * c0.x - c0.y * c2.x) * det;
    i2.x =  (c1.z * c0.y - c0.z * c1.y) * det;
    i2.y = -(c1.z * c0.x - c0.z * c1.x) * det;
    i2.z =  (c1.y * c0.x - c0.y * c1.x) * det;
    auto fdot = [](V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; };
    out[0] = i2.x; out[1] = i2.y; out[2] = i2.z; out[3] = -fdot(-i2, v2);
    out[4] = i0.x; out[5] = i0.y; out[6] = i0.z; out[7] = fdot(-i0, v2);
    out[8] = i1.x; out[9] = i1.y; out[10] = i1.z; out[11] = fdot(-i1, v2);
    if (out[0] == 0.0f) out[0] = 0.0f;
}

namespace {

struct QEntry { int node, start, end; };

struct Ctx {
    const Scene* sc;
    HLBVHParams p;
    const uint32_t* keys;
    const int32_t* idx;
    std::vector<float> inWoop;         // 12 floats per original triangle
    std::vector<int32_t>* nodes;
    std::vector<int32_t>* woop;
    std::vector<int32_t>* triIndex;
    uint32_t allTris = 0, numLeafs = 0;

    void ensureNode(int i) { if ((size_t)(i + 1) * 16 > nodes->size()) nodes->resize((size_t)(i + 1) * 16, 0); }

    // emitTreeKernel.cu:170-231
    int createLeaf(int start, int end)
    {
        uint32_t numTris = (uint32_t)(end - start);
        uint32_t at = allTris, nl = numLeafs;
        allTris += numTris; numLeafs += 1;
        size_t outW = (size_t)at * 3 + nl;
        if (woop->size() < (outW + numTris * 3 + 1) * 4) woop->resize((outW + numTris * 3 + 1) * 4, 0);
        if (triIndex->size() < outW + numTris * 3 + 1) triIndex->resize(outW + numTris * 3 + 1, 0);
        for (uint32_t i = 0; i < numTris; i++) {
            int tri = idx[start + i];
            for (int k = 0; k < 12; k++) (*woop)[(outW + i * 3) * 4 + k] = (int32_t)f2u(inWoop[(size_t)tri * 12 + k]);
            (*triIndex)[outW + i * 3 + 0] = tri;
            (*triIndex)[outW + i * 3 + 1] = 0;
            (*triIndex)[outW + i * 3 + 2] = 0;
        }
        for (int k = 0; k < 4; k++) (*woop)[(outW + numTris * 3) * 4 + k] = (int32_t)0x80000000u;
        (*triIndex)[outW + numTris * 3] = 0;
        return ~(int)(at * 3 + nl);
    }

    // emitTreeKernel.cu:233-381, one launch == one call; threads run in queue order.
    int emitLevel(int level, const std::vector<QEntry>& in, std::vector<QEntry>& out, int inOfs)
    {
        out.clear();
        const int L = p.leafSize;
        for (const QEntry& q : in) {
            int nIdx = q.node, nStart = q.start, nEnd = q.end;
            int split = -1;
            int oldLevel = level, lv = level;
            while (lv >= 0 && (((keys[nStart] >> lv) & 1) == ((keys[nEnd - 1] >> lv) & 1))) lv--;
            if (lv >= 0) {
                uint32_t startBit = (keys[nStart] >> lv) & 1;
                int a = nStart, b = nEnd;
                for (;;) {
                    split = (a + b) >> 1;
                    uint32_t splitBit = (keys[split] >> lv) & 1;
                    if (((keys[split - 1] >> lv) & 1) != splitBit) break;
                    if (splitBit == startBit) a = split; else b = split;
                }
            } else
                split = (nStart + nEnd) >> 1;

            bool leftLeaf = (split - nStart) <= L || oldLevel == 0;
            bool rightLeaf = (nEnd - split) <= L || oldLevel == 0;
            int outIdx = inOfs + (int)out.size();
            ensureNode(nIdx);
            int c0, c1;
            if (leftLeaf) {
                c0 = createLeaf(nStart, split);
                (*nodes)[(size_t)nIdx * 16 + 0] = nStart;
                (*nodes)[(size_t)nIdx * 16 + 1] = split;
            } else {
                out.push_back({outIdx, nStart, split});
                c0 = outIdx * 64;
                outIdx++;
            }
            if (rightLeaf) {
                c1 = createLeaf(split, nEnd);
                (*nodes)[(size_t)nIdx * 16 + 4] = split;
                (*nodes)[(size_t)nIdx * 16 + 5] = nEnd;
            } else {
                out.push_back({outIdx, split, nEnd});
                c1 = outIdx * 64;
            }
            (*nodes)[(size_t)nIdx * 16 + 12] = c0;
            (*nodes)[(size_t)nIdx * 16 + 13] = c1;
            (*nodes)[(size_t)nIdx * 16 + 14] = lv % 3;      // C remainder: -1 % 3 == -1
            (*nodes)[(size_t)nIdx * 16 + 15] = 0;
        }
        return (int)out.size();
    }

    // emitTreeKernel.cu:383-408
    void calcLeaf(int start, int end, V3& lo, V3& hi) const
    {
        V3 e(p.epsilon, p.epsilon, p.epsilon);
        for (int i = start; i < end; i++) {
            int t = idx[i];
            V3 a = sc->v(t, 0), b = sc->v(t, 1), c = sc->v(t, 2);
            lo = fmin3(lo, fmin3(a, fmin3(b, c)) - e);
            hi = fmax3(hi, fmax3(a, fmax3(b, c)) + e);
        }
    }

    // emitTreeKernel.cu:417-562, the active (#else) branch
    void calcAABB(int start, int cnt)
    {
        auto F = [&](int node, int w) { return u2f((uint32_t)(*nodes)[(size_t)node * 16 + w]); };
        for (int q = 0; q < cnt; q++) {
            int n = start + q;
            int32_t* w = &(*nodes)[(size_t)n * 16];
            int nl = w[12], nr = w[13];
            float n0[4], n1[4], n2[4];
            if (nl < 0) {
                V3 lo(F32_MAX, F32_MAX, F32_MAX), hi(-F32_MAX, -F32_MAX, -F32_MAX);
                calcLeaf(w[0], w[1], lo, hi);
                n0[0] = lo.x; n0[1] = hi.x; n0[2] = lo.y; n0[3] = hi.y; n2[0] = lo.z; n2[1] = hi.z;
            } else {
                int c = nl / 64;
                n0[0] = std::fmin(F(c, 0), F(c, 4)); n0[1] = std::fmax(F(c, 1), F(c, 5));
                n0[2] = std::fmin(F(c, 2), F(c, 6)); n0[3] = std::fmax(F(c, 3), F(c, 7));
                n2[0] = std::fmin(F(c, 8), F(c, 10)); n2[1] = std::fmax(F(c, 9), F(c, 11));
            }
            if (nr < 0) {
                V3 lo(F32_MAX, F32_MAX, F32_MAX), hi(-F32_MAX, -F32_MAX, -F32_MAX);
                calcLeaf(w[4], w[5], lo, hi);
                n1[0] = lo.x; n1[1] = hi.x; n1[2] = lo.y; n1[3] = hi.y; n2[2] = lo.z; n2[3] = hi.z;
            } else {
                int c = nr / 64;
                n1[0] = std::fmin(F(c, 0), F(c, 4)); n1[1] = std::fmax(F(c, 1), F(c, 5));
                n1[2] = std::fmin(F(c, 2), F(c, 6)); n1[3] = std::fmax(F(c, 3), F(c, 7));
                n2[2] = std::fmin(F(c, 8), F(c, 10)); n2[3] = std::fmax(F(c, 9), F(c, 11));
            }
            for (int k = 0; k < 4; k++) { w[k] = (int32_t)f2u(n0[k]); w[4 + k] = (int32_t)f2u(n1[k]); w[8 + k] = (int32_t)f2u(n2[k]); }
        }
    }
};

// ordered-int encoding, emitTreeKernel.cu:78-85 (only used to define bin-box semantics; min/max
// over the encoding equals float min/max, so plain floats are used below)

float area3(V3 v) { return (v.x * v.y + v.y * v.z + v.z * v.x) * 2.0f; }

} // namespace

void build_lbvh(const Scene& sc, V3 lo, V3 hi, const HLBVHParams& p, LBVHResult& out)
{
    auto t0 = std::chrono::steady_clock::now();
    const int n = sc.numTris;
    out.sortedKeys.resize(n);
    out.sortedIdx.resize(n);
    morton_codes(sc, lo, hi, out.sortedKeys.data());
    std::iota(out.sortedIdx.begin(), out.sortedIdx.end(), 0);
    sort_pairs_stable(out.sortedKeys.data(), out.sortedIdx.data(), n);

    Ctx c;
    c.sc = &sc; c.p = p; c.keys = out.sortedKeys.data(); c.idx = out.sortedIdx.data();
    c.nodes = &out.bvh.nodes; c.woop = &out.bvh.woop; c.triIndex = &out.bvh.triIndex;
    out.bvh.nodes.clear(); out.bvh.woop.clear(); out.bvh.triIndex.clear();
    c.inWoop.resize((size_t)n * 12);
    for (int t = 0; t < n; t++) calc_woop_gpu(sc.v(t, 0), sc.v(t, 1), sc.v(t, 2), &c.inWoop[(size_t)t * 12]);

    std::vector<int>& lvl = out.levelNodes;
    lvl.clear();
    lvl.push_back(1);                                   // HLBVHBuilder.cpp:559 / :705
    const int nbits = 30;
    uint32_t nodeWritten = 1, nodeCreated = 1;
    std::vector<QEntry> q0, q1;
    int bitOfs = 0;
    out.numClusters = 0;
    c.ensureNode(0);

    const bool lbvhOnly = !p.hlbvh || p.hlbvhBits == 10;   // HLBVHBuilder.cpp:44-47
    if (lbvhOnly) {
        q0.push_back({0, 0, n});
    } else {
        // ---------------- createClustersC + buildTopLevel (HLBVHBuilder.cpp:98-317) ------------
        const int m = 10 - p.hlbvhBits;
        const int d = 3 * (10 - m);
        bitOfs = 3 * m;
        std::vector<int> clsStart;
        for (int i = 0; i < n; i++)
            if (i == 0 || (d != 0 && (c.keys[i] >> d) != (c.keys[i - 1] >> d)) || d == 0) clsStart.push_back(i);
        const int clsCnt = (int)clsStart.size();
        clsStart.push_back(n);
        out.numClusters = clsCnt;
        std::vector<V3> clsLo(clsCnt, V3(F32_MAX, F32_MAX, F32_MAX)), clsHi(clsCnt, V3(-F32_MAX, -F32_MAX, -F32_MAX));
        for (int ci = 0; ci < clsCnt; ci++)
            for (int i = clsStart[ci]; i < clsStart[ci + 1]; i++) {
                int t = c.idx[i];
                V3 a = sc.v(t, 0), b = sc.v(t, 1), cc = sc.v(t, 2);
                clsLo[ci] = fmin3(clsLo[ci], fmin3(a, fmin3(b, cc)));
                clsHi[ci] = fmax3(clsHi[ci], fmax3(a, fmax3(b, cc)));
            }
        std::vector<int> clsSplitId(clsCnt, 0);
        std::vector<int> clsBin((size_t)clsCnt * 3, 0);

        struct Task { V3 lo, hi; int cnt; int id; int plane; int child; };
        std::vector<Task> qi, qo;
        qi.push_back({lo, hi, clsCnt, 0, 0, -1});
        uint32_t sahCreated = 1, sahWritten = 1, oldTerminated = 0, oofs = 0;
        const int B = 8;
        struct BinBox { V3 lo, hi; int cnt; };

        while (sahCreated > 0) {
            std::vector<BinBox> bins((size_t)sahCreated * B * 3);
            for (auto& b : bins) { b.lo = V3(F32_MAX, F32_MAX, F32_MAX); b.hi = V3(-F32_MAX, -F32_MAX, -F32_MAX); b.cnt = 0; }
            // fillBins (emitTreeKernel.cu:713-777)
            for (int ci = 0; ci < clsCnt; ci++) {
                int nid = clsSplitId[ci];
                if (nid < 0) continue;
                V3 a0 = clsLo[ci], a1 = clsHi[ci];
                V3 half = V3((a1.x - a0.x) / 2.0f, (a1.y - a0.y) / 2.0f, (a1.z - a0.z) / 2.0f);
                V3 mid = a0 + half;
                const Task& t = qi[nid];
                V3 step = V3((t.hi.x - t.lo.x) / 8.0f, (t.hi.y - t.lo.y) / 8.0f, (t.hi.z - t.lo.z) / 8.0f);
                V3 qv = (mid - t.lo) / step;
                int bid[3];
                for (int k = 0; k < 3; k++) bid[k] = clamp_via_float(cvt_sat(std::floor(qv[k])), 0, B - 1);
                for (int k = 0; k < 3; k++) {
                    clsBin[(size_t)ci * 3 + k] = bid[k];
                    BinBox& bb = bins[((size_t)nid * 3 + k) * B + bid[k]];
                    bb.lo = fmin3(bb.lo, a0); bb.hi = fmax3(bb.hi, a1); bb.cnt++;
                }
            }
            // findSplit (emitTreeKernel.cu:779-953), threads in task order
            qo.clear();
            uint32_t created = 0;
            std::vector<int> taskAxisLeaf(sahCreated), taskPlane(sahCreated), taskChild(sahCreated), taskArrivals(sahCreated, 0);
            for (uint32_t nid = 0; nid < sahCreated; nid++) {
                const BinBox* bb = &bins[(size_t)nid * 3 * B];
                V3 mnL, mxL, mnR, mxR;
                int cntL = 0, cntR = 0, axis = 0, split = -1;
                float sah = F32_MAX;
                for (int a = 0; a < 3; a++) {
                    V3 mn[B - 1], mx[B - 1]; int cnt[B - 1];
                    V3 mnr(F32_MAX, F32_MAX, F32_MAX), mxr(-F32_MAX, -F32_MAX, -F32_MAX);
                    int cc = 0;
                    for (int b = B - 1; b > 0; b--) {
                        mnr = fmin3(mnr, bb[a * B + b].lo); mxr = fmax3(mxr, bb[a * B + b].hi);
                        mn[b - 1] = mnr; mx[b - 1] = mxr;
                        cc += bb[a * B + b].cnt; cnt[b - 1] = cc;
                    }
                    V3 mnl(F32_MAX, F32_MAX, F32_MAX), mxl(-F32_MAX, -F32_MAX, -F32_MAX);
                    cc = 0;
                    for (int b = 0; b < B - 1; b++) {
                        mnl = fmin3(mnl, bb[a * B + b].lo); mxl = fmax3(mxl, bb[a * B + b].hi);
                        cc += bb[a * B + b].cnt;
                        float s = cc * area3(mxl - mnl) + cnt[b] * area3(mx[b] - mn[b]);
                        if (s < sah) {
                            sah = s; split = b; axis = a; cntL = cc; cntR = cnt[b];
                            mnL = mnl; mxL = mxl; mnR = mn[b]; mxR = mx[b];
                        }
                    }
                }
                const Task& t = qi[nid];
                if (split == -1) {
                    for (int i = 0; i < B; i++)
                        if (bb[0 * B + i].cnt != 0) {
                            // reference leaves mx_right unassigned here (emitTreeKernel.cu:844-845);
                            // the restatement gives both children the occupied bin's box.
                            mnL = mnR = bb[i].lo; mxL = mxR = bb[i].hi;
                            break;
                        }
                    cntR = t.cnt / 2;
                    cntL = t.cnt - cntR;
                    split = -cntL;
                    axis = 0;
                }
                int nodesNew = (cntL > 1) + (cntR > 1);
                uint32_t ofs = created;
                int idxN = (int)(sahWritten + ofs);
                created += nodesNew;
                int val = 0, l = 0, r = 0;
                if (cntL > 1) { l = idxN * 64; qo.push_back({mnL, mxL, cntL, idxN, 0, -1}); val++; }
                if (cntR > 1) { r = (idxN + val) * 64; qo.push_back({mnR, mxR, cntR, idxN + val, 0, -1}); }
                taskAxisLeaf[nid] = axis | ((((cntL <= 1) << 1) | (cntR <= 1)) << 2);
                taskPlane[nid] = split;
                taskChild[nid] = (int)ofs;
                c.ensureNode(t.id);
                int32_t* w = &(*c.nodes)[(size_t)t.id * 16];
                w[12] = l; w[13] = r; w[14] = axis; w[15] = 0;
            }
            // distribute (emitTreeKernel.cu:955-1027), threads in cluster order
            for (int ci = 0; ci < clsCnt; ci++) {
                int old = clsSplitId[ci];
                if (old < 0) continue;
                int splitId = taskPlane[old];
                int childId = taskChild[old];
                int qid = qi[old].id;
                int axis = taskAxisLeaf[old] & 0xF;
                int leafs = axis >> 2;
                axis &= 3;
                int cnt = clsStart[ci + 1] - clsStart[ci];
                int binId;
                if (splitId < 0) { splitId = (-splitId) - 1; binId = taskArrivals[old]++; }
                else binId = clsBin[(size_t)ci * 3 + axis];
                if (leafs == 0) { clsSplitId[ci] = childId + (binId <= splitId ? 0 : 1); continue; }
                bool goLeft = binId <= splitId;
                bool single = goLeft ? (leafs & 2) : (leafs & 1);
                if (!single) { clsSplitId[ci] = childId; continue; }
                c.ensureNode(qid);
                int word = goLeft ? 12 : 13;
                int rw = goLeft ? 0 : 4;
                if (cnt <= p.leafSize) {
                    int leaf = c.createLeaf(clsStart[ci], clsStart[ci + 1]);
                    int32_t* w = &(*c.nodes)[(size_t)qid * 16];
                    w[word] = leaf; w[rw] = clsStart[ci]; w[rw + 1] = clsStart[ci + 1];
                } else {
                    uint32_t idxN = created++;
                    oofs++;
                    q0.push_back({(int)(sahWritten + idxN), clsStart[ci], clsStart[ci + 1]});
                    (*c.nodes)[(size_t)qid * 16 + word] = (int)(sahWritten + idxN) * 64;
                }
                clsSplitId[ci] = -1;
            }
            uint32_t terminated = oofs - oldTerminated;
            oldTerminated = oofs;
            sahCreated = created;
            if (sahCreated != 0) lvl.push_back((int)sahCreated);
            sahWritten += sahCreated;
            sahCreated -= terminated;
            qi.swap(qo);
        }
        nodeWritten = sahWritten;
        nodeCreated = oofs;
    }

    // ---------------- buildBottomLevel (HLBVHBuilder.cpp:319-406) ----------------------------
    {
        uint32_t level = 0;
        std::vector<QEntry>* qin = &q0; std::vector<QEntry>* qout = &q1;
        while (level < (uint32_t)(nbits - bitOfs) && nodeCreated > 0) {
            nodeCreated = (uint32_t)c.emitLevel(nbits - (int)(level + 1 + bitOfs), *qin, *qout, (int)nodeWritten);
            lvl.push_back((int)nodeCreated);
            nodeWritten += nodeCreated;
            if (lvl.back() == 0) lvl.pop_back();
            std::swap(qin, qout);
            level++;
        }
    }
    out.bvh.nodes.resize((size_t)nodeWritten * 16, 0);
    out.bvh.woop.resize(((size_t)n * 3 + c.numLeafs) * 4);
    out.bvh.triIndex.resize((size_t)n * 3 + c.numLeafs);

    // ---------------- calcAABB, bottom-up per level (HLBVHBuilder.cpp:408-449) ---------------
    {
        uint32_t w = nodeWritten;
        for (int l = (int)lvl.size() - 1; l >= 0; l--) { w -= lvl[l]; c.calcAABB((int)w, lvl[l]); }
    }
    out.numNodes = (int)nodeWritten;
    out.numLeaves = (int)c.numLeafs;
    out.buildSeconds = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count();
}

// ---- canonical form --------------------------------------------------------------------------
namespace {
int canon_rec(const int32_t* nodes, const int32_t* woop, const int32_t* triIndex, int addr, Canonical& out)
{
    if (addr < 0) {
        int nt = 0;
        for (int a = ~addr; (uint32_t)woop[a * 4] != 0x80000000u; a += 3) {
            out.tris.push_back(triIndex[a]);
            for (int k = 0; k < 12; k++) out.woop.push_back(u2f((uint32_t)woop[a * 4 + k]));
            nt++;
        }
        out.leafSizes.push_back(nt);
        return nt;
    }
    const int32_t* w = nodes + addr / 4;
    size_t slot = out.inner.size();
    out.inner.resize(slot + 3);
    for (int k = 0; k < 12; k++) out.boxes.push_back(u2f((uint32_t)w[k]));
    int l = canon_rec(nodes, woop, triIndex, w[12], out);
    int r = canon_rec(nodes, woop, triIndex, w[13], out);
    out.inner[slot] = l; out.inner[slot + 1] = r; out.inner[slot + 2] = w[14];
    return l + r;
}
}

void canonicalize(const int32_t* nodes, const int32_t* woop, const int32_t* triIndex, Canonical& out)
{
    out = Canonical();
    canon_rec(nodes, woop, triIndex, 0, out);
}

} // namespace orc
