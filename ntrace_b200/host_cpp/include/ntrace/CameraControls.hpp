// FW::CameraControls — camera state, the reference's signature codec and the matrices Renderer needs.
// Reference: src/framework/3d/CameraControls.cpp:250-284 (orientation, worldToCamera), :362-395 (decodeSignature),
// :491-541 (6-bit float / direction codec), Renderer.cpp:473-477 (nscreenToWorld).  No GUI, no input handling.
#pragma once
#include "ntrace/Base.hpp"

namespace FW
{
class CameraControls
{
public:
    CameraControls() : m_position(0.0f, 0.0f, 1.5f), m_forward(0.0f, 0.0f, -1.0f), m_up(0.0f, 1.0f, 0.0f),
                       m_keepAligned(false), m_speed(1.0f), m_fov(70.0f), m_near(0.001f), m_far(3.0f) {}

    const Vec3f& getPosition() const { return m_position; }
    const Vec3f& getForward() const { return m_forward; }
    const Vec3f& getUp() const { return m_up; }
    F32 getFOV() const { return m_fov; }
    F32 getNear() const { return m_near; }
    F32 getFar() const { return m_far; }
    void setPosition(const Vec3f& v) { m_position = v; }
    void setForward(const Vec3f& v) { m_forward = v; }
    void setUp(const Vec3f& v) { m_up = v; }
    void setFOV(F32 v) { m_fov = v; }
    void setNear(F32 v) { m_near = v; }
    void setFar(F32 v) { m_far = v; }

    // CameraControls.cpp:250-284
    Mat4f getWorldToCamera() const
    {
        Vec3f c2 = -m_forward.normalized();
        Vec3f c0 = cross(m_up, c2).normalized();
        Vec3f c1 = cross(c2, c0).normalized();
        const Vec3f col[3] = {c0, c1, c2};
        Mat4f r;
        for (int i = 0; i < 3; i++) {
            r(i, 0) = col[i].x; r(i, 1) = col[i].y; r(i, 2) = col[i].z;
            r(i, 3) = -dot(col[i], m_position);                 // -(orientation^T * position)[i]
        }
        return r;
    }
    Mat4f getCameraToClip() const { return Mat4f::perspective(m_fov, m_near, m_far); }
    Mat4f getWorldToClip() const { return getCameraToClip() * getWorldToCamera(); }
    // Renderer.cpp:473-477: invert(gl->xformFitToView(-1, 2) * worldToClip)
    Mat4f getNScreenToWorld(int w, int h) const { return invert(Mat4f::fitToView(w, h) * getWorldToClip()); }

    // CameraControls.cpp:362-395
    void decodeSignature(const String& sig)
    {
        String s = sig;
        while (!s.empty() && (s[0] == '"' || s[0] == ' ')) s.erase(0, 1);
        while (!s.empty() && (s[s.size() - 1] == '"' || s[s.size() - 1] == ',' || s[s.size() - 1] == ' ')) s.erase(s.size() - 1);
        const char* src = s.c_str();
        const char* end = src + s.size();
        F32 px = decodeFloat(src, end), py = decodeFloat(src, end), pz = decodeFloat(src, end);
        Vec3f forward = decodeDirection(src, end);
        Vec3f up = decodeDirection(src, end);
        F32 speed = decodeFloat(src, end), fov = decodeFloat(src, end), znear = decodeFloat(src, end), zfar = decodeFloat(src, end);
        bool keepAligned = decodeBits(src, end) != 0;
        m_position = Vec3f(px, py, pz); m_forward = forward; m_up = up;
        m_speed = speed; m_fov = fov; m_near = znear; m_far = zfar; m_keepAligned = keepAligned;
    }

private:
    // CameraControls.cpp:491-541
    static U32 decodeBits(const char*& src, const char* end)
    {
        if (src >= end) fail("CameraControls: Invalid signature!");
        char c = *src++;
        if (c >= '/' && c <= ':') return (U32)(c - '/');
        if (c >= 'A' && c <= 'Z') return (U32)(c - 'A' + 12);
        if (c >= 'a' && c <= 'z') return (U32)(c - 'a' + 38);
        fail("CameraControls: Invalid signature!");
        return 0;
    }
    static F32 decodeFloat(const char*& src, const char* end)
    {
        U32 bits = 0;
        for (int i = 0; i < 32; i += 6) bits |= decodeBits(src, end) << i;
        F32 f; memcpy(&f, &bits, 4);
        return f;
    }
    static Vec3f decodeDirection(const char*& src, const char* end)
    {
        U32 face = decodeBits(src, end);
        F32 t = ((face & 4) == 0) ? 1.0f : -1.0f, u = 0.0f, v = 0.0f;
        if ((face & 8) == 0) { u = decodeFloat(src, end); v = decodeFloat(src, end); }
        Vec3f tuv = Vec3f(t, u, v).normalized();
        switch (face & 3) {
        case 0: return tuv;
        case 1: return Vec3f(tuv.z, tuv.x, tuv.y);
        default: return Vec3f(tuv.y, tuv.z, tuv.x);
        }
    }

    Vec3f m_position, m_forward, m_up;
    bool m_keepAligned;
    F32 m_speed, m_fov, m_near, m_far;
};
}
