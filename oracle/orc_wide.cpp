// ORACLE — TEST INFRASTRUCTURE ONLY.  See orc_wide.hpp.
#include "orc_wide.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

namespace {

inline float as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int32_t as_int(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
// product: qplane<I>() in nt_wide.cu — bytes (0x00, q.I, 0x80, 0x3F) = 1 + q * 2^-15
inline float qplane(uint32_t word, int i) { return as_float(0x3F800000u | (((word >> (8 * i)) & 0xffu) << 8)); }

void trace_wide4_one(const uint32_t* wnodes, const int32_t* woop, const int32_t* triIndex, Ray& ray, RayResult& res, bool closest, uint32_t* cnt)
{
    // ray setup as the kernels do it (fermi_speculative_while_while.cu:94-107)
    const float ooeps = std::exp2(-80.0f);
    const float dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    const float idirx = 1.0f / (std::fabs(dx) > ooeps ? dx : std::copysign(ooeps, dx));
    const float idiry = 1.0f / (std::fabs(dy) > ooeps ? dy : std::copysign(ooeps, dy));
    const float idirz = 1.0f / (std::fabs(dz) > ooeps ? dz : std::copysign(ooeps, dz));
    const float oodx = ray.o.x * idirx, oody = ray.o.y * idiry, oodz = ray.o.z * idirz;
    const bool ngx = idirx < 0.0f, ngy = idiry < 0.0f, ngz = idirz < 0.0f;

    int stack[160];
    int sp = 0;
    int node = 0;
    for (;;) {
        if (node < 0) {
            if (cnt) cnt[2]++;
            for (int triAddr = ~node;; triAddr += 3) {
                if ((uint32_t)woop[triAddr * 4] == 0x80000000u) break;
                if (cnt) cnt[1]++;
                const float* z = reinterpret_cast<const float*>(woop + triAddr * 4);
                float t = ray_triangle_woop(z, z + 4, z + 8, ray);
                if (t > ray.tmin && t < ray.tmax) {
                    ray.tmax = t;
                    res.t = t;
                    res.id = triIndex[triAddr];
                    if (!closest) return;
                }
            }
        } else {
            if (cnt) cnt[0]++;
            const uint32_t* w = wnodes + (size_t)node * 16;
            const float ax = as_float(w[3]) * idirx, ay = as_float(w[4]) * idiry, az = as_float(w[5]) * idirz;
            const float bx = std::fmaf(as_float(w[0]), idirx, -oodx) - ax;
            const float by = std::fmaf(as_float(w[1]), idiry, -oody) - ay;
            const float bz = std::fmaf(as_float(w[2]), idirz, -oodz) - az;
            const uint32_t nx = ngx ? w[9] : w[6], fx = ngx ? w[6] : w[9];
            const uint32_t ny = ngy ? w[10] : w[7], fy = ngy ? w[7] : w[10];
            const uint32_t nz = ngz ? w[11] : w[8], fz = ngz ? w[8] : w[11];
            int key[4];
            for (int i = 0; i < 4; i++) {
                const float tn = std::fmax(std::fmax(std::fmaf(qplane(nx, i), ax, bx), std::fmaf(qplane(ny, i), ay, by)),
                                           std::fmax(std::fmaf(qplane(nz, i), az, bz), ray.tmin));
                const float tf = std::fmin(std::fmin(std::fmaf(qplane(fx, i), ax, bx), std::fmaf(qplane(fy, i), ay, by)),
                                           std::fmin(std::fmaf(qplane(fz, i), az, bz), ray.tmax));
                key[i] = (tn <= tf) ? ((as_int(tn) & ~3) | i) : 0x7fffffff;
            }
            std::sort(key, key + 4);
            if (key[0] != 0x7fffffff) {
                for (int j = 3; j >= 1; j--)
                    if (key[j] != 0x7fffffff) stack[sp++] = (int)w[12 + (key[j] & 3)];
                node = (int)w[12 + (key[0] & 3)];
                continue;
            }
        }
        if (sp == 0) return;
        node = stack[--sp];
    }
}

} // namespace

void trace_wide4(const uint32_t* wnodes, const int32_t* woop, const int32_t* triIndex,
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
        res.t = ray.tmax;
        if (counters) { counters[3 * i] = counters[3 * i + 1] = counters[3 * i + 2] = 0; }
        trace_wide4_one(wnodes, woop, triIndex, ray, res, closest, counters ? counters + 3 * i : nullptr);
    }
}

namespace {
struct Box { double lo[3], hi[3]; };

// union of the leaf boxes below a binary link, keyed by leaf link: the boxes the reference's tree assigns to each leaf
void collect_leaves(const int32_t* nodes, size_t nodeBytes, int layout, int addr, std::map<int, Box>& leaves, int depth, bool& ok)
{
    if (!ok || depth > 4096) { ok = false; return; }
    const size_t byteOfs = (layout == 5) ? (size_t)(uint32_t)addr * 16 : (size_t)(uint32_t)addr;
    if (byteOfs % 64 || byteOfs + 64 > nodeBytes) { ok = false; return; }
    const int32_t* w = nodes + byteOfs / 4;
    const float* f = reinterpret_cast<const float*>(w);
    Box b[2];
    b[0].lo[0] = f[0]; b[0].hi[0] = f[1]; b[0].lo[1] = f[2]; b[0].hi[1] = f[3]; b[0].lo[2] = f[8]; b[0].hi[2] = f[9];
    b[1].lo[0] = f[4]; b[1].hi[0] = f[5]; b[1].lo[1] = f[6]; b[1].hi[1] = f[7]; b[1].lo[2] = f[10]; b[1].hi[2] = f[11];
    for (int c = 0; c < 2; c++) {
        const int link = w[12 + c];
        if (link < 0) {
            if (leaves.count(link)) { ok = false; return; }
            leaves[link] = b[c];
        } else collect_leaves(nodes, nodeBytes, layout, link, leaves, depth + 1, ok);
    }
}
}

int check_wide4(const uint32_t* wnodes, size_t numWide, const int32_t* nodes, size_t nodeBytes, int layout, double out[4])
{
    std::map<int, Box> leaves;
    bool ok = true;
    collect_leaves(nodes, nodeBytes, layout, 0, leaves, 0, ok);
    if (!ok) return 1;
    // walk the wide tree; every leaf link must be one of the binary tree's, seen once, inside its decoded box; every inner
    // child's decoded box must contain the decoded boxes of its own children (checked through the leaves: a leaf box must be
    // inside the decoded box of EVERY ancestor slot)
    struct Item { uint32_t node; int depth; std::vector<Box> anc; };
    std::vector<Item> st;
    st.push_back({0, 1, {}});
    std::map<int, int> seen;
    size_t visited = 0;
    int maxDepth = 0;
    double worst = 0.0;
    while (!st.empty()) {
        Item it = std::move(st.back());
        st.pop_back();
        if (it.node >= numWide || ++visited > numWide) return 2;
        maxDepth = std::max(maxDepth, it.depth);
        const uint32_t* w = wnodes + (size_t)it.node * 16;
        for (int i = 0; i < 4; i++) {
            Box d;
            bool inverted = false;
            double step[3];
            for (int a = 0; a < 3; a++) {
                const double p = as_float(w[a]), s = (double)as_float(w[3 + a]) / 32768.0;
                const uint32_t ql = (w[6 + a] >> (8 * i)) & 0xff, qh = (w[9 + a] >> (8 * i)) & 0xff;
                if (ql > qh) inverted = true;
                d.lo[a] = p + ql * s; d.hi[a] = p + qh * s; step[a] = s;
            }
            const int link = (int)w[12 + i];
            if (inverted) continue;                    // unused slot
            auto inside = [&](const Box& in, const Box& outb) {
                for (int a = 0; a < 3; a++) {
                    if (in.lo[a] < outb.lo[a]) return false;
                    if (in.hi[a] > outb.hi[a]) return false;
                }
                return true;
            };
            if (link < 0) {
                auto f = leaves.find(link);
                if (f == leaves.end()) return 3;
                if (seen[link]++) return 4;
                if (!inside(f->second, d)) return 5;
                for (const Box& a : it.anc) if (!inside(f->second, a)) return 6;
                for (int a = 0; a < 3; a++) {
                    worst = std::max(worst, (f->second.lo[a] - d.lo[a]) / step[a]);
                    worst = std::max(worst, (d.hi[a] - f->second.hi[a]) / step[a]);
                }
            } else {
                Item c{(uint32_t)link, it.depth + 1, it.anc};
                c.anc.push_back(d);
                st.push_back(std::move(c));
            }
        }
    }
    if (seen.size() != leaves.size()) return 7;
    if (visited != numWide) return 8;
    out[0] = (double)visited; out[1] = (double)leaves.size(); out[2] = maxDepth; out[3] = worst;
    return 0;
}

} // namespace orc
