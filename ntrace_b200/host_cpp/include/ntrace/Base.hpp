// ntrace_b200 C++ host — base types.  Header-only; everything that touches the GPU goes through the C ABI of
// include/ntrace_b200.h (libntrace_b200.so), this layer links no CUDA library.
//
// Mirrors, with the reference's names and argument meaning, the small part of FW:: the tracing path uses:
//   src/framework/base/Defs.hpp:136-151   fail / setError  (here: FW::fail throws FW::Error instead of aborting, so a host
//                                         application can catch what the reference turns into a message box + exit)
//   src/framework/base/Math.hpp           Vec2i / Vec3i / Vec3f / Vec4f / Mat4f (only the members the path calls)
//   src/rt/Util.hpp:35-87                 AABB, Ray, RayResult
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" {
#include "ntrace_b200.h"
}

namespace FW
{
typedef int32_t S32;
typedef uint32_t U32;
typedef int64_t S64;
typedef uint8_t U8;
typedef float F32;
typedef std::string String;

class Error : public std::runtime_error
{
public:
    explicit Error(const std::string& msg) : std::runtime_error(msg) {}
};

inline void fail(const char* fmt, ...)
{
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Error(buf);
}

// every C-ABI call goes through this: non-zero -> the library's message, as an Error
inline void ntCheck(int rc)
{
    if (rc != 0) throw Error(nt_last_error());
}

struct Vec2i { S32 x, y; Vec2i(S32 x_ = 0, S32 y_ = 0) : x(x_), y(y_) {} };
struct Vec3i { S32 x, y, z; Vec3i(S32 x_ = 0, S32 y_ = 0, S32 z_ = 0) : x(x_), y(y_), z(z_) {} };

struct Vec3f
{
    F32 x, y, z;
    Vec3f(F32 a = 0.0f) : x(a), y(a), z(a) {}
    Vec3f(F32 x_, F32 y_, F32 z_) : x(x_), y(y_), z(z_) {}
    Vec3f operator+(const Vec3f& v) const { return Vec3f(x + v.x, y + v.y, z + v.z); }
    Vec3f operator-(const Vec3f& v) const { return Vec3f(x - v.x, y - v.y, z - v.z); }
    Vec3f operator-() const { return Vec3f(-x, -y, -z); }
    Vec3f operator*(F32 s) const { return Vec3f(x * s, y * s, z * s); }
    F32 dot(const Vec3f& v) const { F32 r = 0.0f; r += x * v.x; r += y * v.y; r += z * v.z; return r; }
    Vec3f cross(const Vec3f& v) const { return Vec3f(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x); }
    F32 length() const { return std::sqrt(dot(*this)); }
    Vec3f normalized() const { F32 l = length(); return *this * (l != 0.0f ? 1.0f / l : 0.0f); }      // rcp(0) == 0 (Math.hpp:114)
    Vec3f min(const Vec3f& v) const { return Vec3f(x < v.x ? x : v.x, y < v.y ? y : v.y, z < v.z ? z : v.z); }
    Vec3f max(const Vec3f& v) const { return Vec3f(x > v.x ? x : v.x, y > v.y ? y : v.y, z > v.z ? z : v.z); }
    const F32* getPtr() const { return &x; }
};
inline Vec3f cross(const Vec3f& a, const Vec3f& b) { return a.cross(b); }
inline F32 dot(const Vec3f& a, const Vec3f& b) { return a.dot(b); }

struct Vec4f { F32 x, y, z, w; Vec4f(F32 x_ = 0, F32 y_ = 0, F32 z_ = 0, F32 w_ = 0) : x(x_), y(y_), z(z_), w(w_) {} };

// row-major m[r][c] (the reference stores columns; only products and the inverse are used here, restated to give the
// same fp32 results: products accumulate in column order k = 0..3, the inverse is the cofactor form of Math.hpp:1024-1045)
struct Mat4f
{
    F32 m[4][4];
    Mat4f() { setIdentity(); }
    void setIdentity() { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i][j] = (i == j) ? 1.0f : 0.0f; }
    F32& operator()(int r, int c) { return m[r][c]; }
    F32 operator()(int r, int c) const { return m[r][c]; }
    Mat4f operator*(const Mat4f& b) const
    {
        Mat4f r;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                F32 s = 0.0f;
                for (int k = 0; k < 4; k++) s += m[i][k] * b.m[k][j];
                r.m[i][j] = s;
            }
        return r;
    }
    static F32 det3(const F32 v[3][3])
    {
        return v[0][0] * v[1][1] * v[2][2] - v[0][0] * v[1][2] * v[2][1] + v[1][0] * v[2][1] * v[0][2] -
               v[1][0] * v[2][2] * v[0][1] + v[2][0] * v[0][1] * v[1][2] - v[2][0] * v[0][2] * v[1][1];
    }
    Mat4f inverted() const
    {
        Mat4f r;
        F32 d = 0.0f, si = 1.0f;
        for (int i = 0; i < 4; i++) {
            F32 sj = si;
            for (int j = 0; j < 4; j++) {
                F32 sub[3][3];
                int rr = 0;
                for (int k = 0; k < 4; k++) {
                    if (k == j) continue;
                    int cc = 0;
                    for (int l = 0; l < 4; l++) { if (l == i) continue; sub[rr][cc++] = m[k][l]; }
                    rr++;
                }
                F32 dd = det3(sub) * sj;
                r.m[i][j] = dd;
                d += dd * m[j][i];
                sj = -sj;
            }
            si = -si;
        }
        F32 rd = (d != 0.0f) ? 1.0f / d : 0.0f;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = r.m[i][j] * rd * 4.0f;
        return r;
    }
    // Math.cpp:66-92
    static Mat4f fitToView(int w, int h)
    {
        F32 vx = (F32)w, vy = (F32)h;
        F32 s = (vx / 2.0f < vy / 2.0f) ? vx / 2.0f : vy / 2.0f;
        Mat4f r;
        r.m[0][0] = 2.0f / vx * s;
        r.m[1][1] = 2.0f / vy * s;
        return r;
    }
    static Mat4f perspective(F32 fov, F32 nearDist, F32 farDist)
    {
        const F32 pi = 3.14159265358979323846f;
        F32 f = 1.0f / std::tan(fov * pi / 360.0f);
        F32 d = 1.0f / (nearDist - farDist);
        Mat4f r;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = 0.0f;
        r.m[0][0] = f; r.m[1][1] = f;
        r.m[2][2] = (nearDist + farDist) * d;
        r.m[2][3] = 2.0f * nearDist * farDist * d;
        r.m[3][2] = -1.0f;
        return r;
    }
    const F32* getPtr() const { return &m[0][0]; }
};
inline Mat4f invert(const Mat4f& a) { return a.inverted(); }

// src/rt/Util.hpp:35-87
class AABB
{
public:
    AABB() : m_mn(3.402823466e+38f), m_mx(-3.402823466e+38f) {}
    AABB(const Vec3f& mn, const Vec3f& mx) : m_mn(mn), m_mx(mx) {}
    void grow(const Vec3f& p) { m_mn = m_mn.min(p); m_mx = m_mx.max(p); }
    const Vec3f& min() const { return m_mn; }
    const Vec3f& max() const { return m_mx; }
private:
    Vec3f m_mn, m_mx;
};

struct Ray
{
    Ray() : origin(0.0f), tmin(0.0f), direction(0.0f), tmax(0.0f) {}
    void degenerate() { tmax = tmin - 1.0f; }
    Vec3f origin; F32 tmin; Vec3f direction; F32 tmax;
};
struct RayResult
{
    RayResult(S32 ii = -1, F32 ti = 0.0f) : id(ii), t(ti), padA(0), padB(0) {}
    bool hit() const { return id != -1; }
    void clear() { id = -1; }
    S32 id; F32 t; S32 padA; S32 padB;
};
static_assert(sizeof(Ray) == 32 && sizeof(RayResult) == 16, "Ray / RayResult layout (Util.hpp:62-87)");

// kernels/CudaTracerKernels.hpp:52-63
enum BVHLayout
{
    BVHLayout_AOS_AOS = 0, BVHLayout_AOS_SOA, BVHLayout_SOA_AOS, BVHLayout_SOA_SOA,
    BVHLayout_Compact, BVHLayout_Compact2, BVHLayout_CPU, BVHLayout_Max
};
}
