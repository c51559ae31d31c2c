// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the arithmetic NTrace's tracing/build path uses.  Nothing in the
// product path (ntrace_b200/, include/) may include, link or call this; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// Parity status: the reference ships no golden vectors for this path, so the
// oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/_ref/libref.so is the reference's
// own CPU sources (SAHBVHBuilder, SplitBVHBuilder, BVHNode SAH, BVH::trace, CudaBVH createCompact /
// woopifyTri / trace, Intersect::*, PixelTable) compiled unmodified (oracle/Makefile `ref`), and
// tests/test_reference_pin.py requires bit-identical results from this restatement, live and against
// frozen fixtures (tests/golden/ref_*).  PINNED on the CPU: orc_math, orc_bvh, pixel_table.  orc_lbvh (HLBVH
// kernels) and the ray generators of orc_raygen restate DEVICE code: they are pinned on the GPU box, where the
// reference's own kernels, compiled for sm_100a (oracle/ref_gpu.py), run beside the product kernels
// (tests/test_gpu_reference_kernels.py); in the container they are cross-validated in tests/test_oracle_*.py.
//
// Compile with: -O2 -ffp-contract=off -fno-fast-math  (IEEE fp32, no FMA contraction).
//
// Restates: src/framework/base/Math.hpp (vector ops :114-204, :400; det/inverse :965-1045),
//           src/rt/Util.hpp:35-87 (AABB, Ray, RayResult).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>

namespace orc {

static constexpr float F32_MAX = 3.402823466e+38f;

struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit V3(float a) : x(a), y(a), z(a) {}
    float  operator[](int i) const { return (&x)[i]; }
    float& operator[](int i)       { return (&x)[i]; }
};

static inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(V3 a)       { return V3(-a.x, -a.y, -a.z); }
static inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(float s, V3 a) { return a * s; }
static inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
// Defs.hpp:212-213 — FW::min / FW::max on the host: on equal operands (incl. +0 vs -0) and on NaN they return b,
// unlike std::min / std::max which return a.  Visible in the sign of zero of node boxes, so restated exactly.
static inline float fw_min(float a, float b) { return (a < b) ? a : b; }
static inline float fw_max(float a, float b) { return (a > b) ? a : b; }
static inline V3 vmin(V3 a, V3 b) { return V3(fw_min(a.x, b.x), fw_min(a.y, b.y), fw_min(a.z, b.z)); }
static inline V3 vmax(V3 a, V3 b) { return V3(fw_max(a.x, b.x), fw_max(a.y, b.y), fw_max(a.z, b.z)); }
static inline float hmin(V3 a) { return fw_min(fw_min(a.x, a.y), a.z); }
static inline float hmax(V3 a) { return fw_max(fw_max(a.x, a.y), a.z); }
// Math.hpp:148 — r = v[0]; r += v[1]; r += v[2]
static inline float hsum(V3 a) { float r = a.x; r += a.y; r += a.z; return r; }
// Math.hpp:185 — r = 0; r += a[i]*b[i]
static inline float dot(V3 a, V3 b) { float r = 0.0f; r += a.x * b.x; r += a.y * b.y; r += a.z * b.z; return r; }
// Math.hpp:400
static inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// Math.hpp:114
static inline float rcp(float a) { return (a != 0.0f) ? 1.0f / a : 0.0f; }
// Math.hpp:142-144
static inline float length(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 normalize(V3 a) { return a * (1.0f * rcp(length(a))); }
// Math.hpp:115
static inline V3 lerp(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
static inline float clampf(float v, float lo, float hi) { return fw_min(fw_max(v, lo), hi); }

static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// Util.hpp:35-58
struct AABB {
    V3 mn, mx;
    AABB() : mn(F32_MAX, F32_MAX, F32_MAX), mx(-F32_MAX, -F32_MAX, -F32_MAX) {}
    AABB(V3 a, V3 b) : mn(a), mx(b) {}
    void grow(V3 p) { mn = vmin(mn, p); mx = vmax(mx, p); }
    void grow(const AABB& b) { grow(b.mn); grow(b.mx); }
    void intersect(const AABB& b) { mn = vmax(mn, b.mn); mx = vmin(mx, b.mx); }
    bool valid() const { return mn.x <= mx.x && mn.y <= mx.y && mn.z <= mx.z; }
    float area() const {
        if (!valid()) return 0.0f;
        V3 d = mx - mn;
        return (d.x * d.y + d.y * d.z + d.z * d.x) * 2.0f;
    }
};

// Util.hpp:62-87 — 32-byte ray, 16-byte result
struct Ray { V3 o; float tmin; V3 d; float tmax; };
struct RayResult { int32_t id; float t; int32_t padA; int32_t padB; };
static_assert(sizeof(Ray) == 32 && sizeof(RayResult) == 16, "layout");

// 4x4 matrix, row-major m[r][c]; Math.hpp:1024-1045 (inverse by cofactors, fp32),
// :996-1001 (3x3 determinant term order), :1058-1107 (products).
struct M4 {
    float m[4][4];
    static M4 identity() { M4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = (i == j) ? 1.0f : 0.0f; return r; }
};

static inline float det3(const float v[3][3]) {
    return v[0][0] * v[1][1] * v[2][2] - v[0][0] * v[1][2] * v[2][1] +
           v[1][0] * v[2][1] * v[0][2] - v[1][0] * v[2][2] * v[0][1] +
           v[2][0] * v[0][1] * v[1][2] - v[2][0] * v[0][2] * v[1][1];
}

static inline M4 invert(const M4& a) {
    M4 r;
    float d = 0.0f, si = 1.0f;
    for (int i = 0; i < 4; i++) {
        float sj = si;
        for (int j = 0; j < 4; j++) {
            float sub[3][3];
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < 3; l++)
                    sub[k][l] = a.m[(k < j) ? k : k + 1][(l < i) ? l : l + 1];
            float dd = det3(sub) * sj;
            r.m[i][j] = dd;
            d += dd * a.m[j][i];
            sj = -sj;
        }
        si = -si;
    }
    float rd = rcp(d);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            r.m[i][j] = r.m[i][j] * rd * 4.0f;
    return r;
}

static inline M4 mul(const M4& a, const M4& b) {
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float rr = 0.0f;
            for (int k = 0; k < 4; k++) rr += a.m[i][k] * b.m[k][j];
            r.m[i][j] = rr;
        }
    return r;
}

} // namespace orc
