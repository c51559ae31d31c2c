// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  pixel_table PINNED against the compiled reference (oracle/_ref, CPU); ray generators PINNED on the GPU box (see orc_raygen.hpp).
//
// Ray generation restated on the CPU from:
//   src/rt/ray/PixelTable.cpp:57-141          index <-> pixel tables
//   src/rt/ray/RayGen.cpp:76-112              primaryCPU (the reference's own CPU twin)
//   src/rt/ray/RayGenKernels.cu:34-45,77-125  jenkinsMix, primary kernel (jitter branch)
//   src/rt/ray/RayGenKernels.cu:129-236       rayGenAOKernel (AO and diffuse)
//   src/rt/cuda/RendererKernels.cu:174-224    countHitsKernel semantics
#include "orc_raygen.hpp"

namespace orc {

void pixel_table(int w, int h, int32_t* indexToPixel, int32_t* pixelToIndex)
{
    int idx = 0;
    int bheight = h & ~7, bwidth = w & ~7;
    int maxdim = (bwidth > bheight) ? bwidth : bheight;
    maxdim |= maxdim >> 1; maxdim |= maxdim >> 2; maxdim |= maxdim >> 4; maxdim |= maxdim >> 8; maxdim |= maxdim >> 16;
    maxdim = (maxdim + 1) >> 1;
    int width8 = bwidth >> 3, height8 = bheight >> 3;
    for (int i = 0; i < maxdim * maxdim; i++) {
        int tx = 0, ty = 0, val = i, bit = 1;
        while (val) {
            if (val & 1) tx |= bit;
            if (val & 2) ty |= bit;
            bit += bit;
            val >>= 2;
        }
        if (tx < width8 && ty < height8) {
            for (int inner = 0; inner < 64; inner++) {
                int ix = ((inner & 1) >> 0) | ((inner & 4) >> 1) | ((inner & 16) >> 2);
                int iy = ((inner & 2) >> 1) | ((inner & 8) >> 2) | ((inner & 32) >> 3);
                int pos = (ty * 8 + iy) * w + (tx * 8 + ix);
                if (pixelToIndex) pixelToIndex[pos] = idx;
                indexToPixel[idx++] = pos;
            }
        }
    }
    for (int px = 0; px < bwidth; px++)
        for (int py = bheight; py < h; py++) {
            int pos = px + py * w;
            if (pixelToIndex) pixelToIndex[pos] = idx;
            indexToPixel[idx++] = pos;
        }
    for (int py = 0; py < h; py++)
        for (int px = bwidth; px < w; px++) {
            int pos = px + py * w;
            if (pixelToIndex) pixelToIndex[pos] = idx;
            indexToPixel[idx++] = pos;
        }
}

static inline void jenkins_mix(uint32_t& a, uint32_t& b, uint32_t& c)
{
    a -= b; a -= c; a ^= (c >> 13);
    b -= c; b -= a; b ^= (a << 8);
    c -= a; c -= b; c ^= (b >> 13);
    a -= b; a -= c; a ^= (c >> 12);
    b -= c; b -= a; b ^= (a << 16);
    c -= a; c -= b; c ^= (b >> 5);
    a -= b; a -= c; a ^= (c >> 3);
    b -= c; b -= a; b ^= (a << 10);
    c -= a; c -= b; c ^= (b >> 15);
}

void raygen_primary(Ray* rays, int32_t* idToSlot, int32_t* slotToID, V3 origin, const M4& n2w,
                    int w, int h, float maxDist, uint32_t randomSeed)
{
    std::vector<int32_t> i2p((size_t)w * h);
    pixel_table(w, h, i2p.data(), nullptr);
    for (int i = 0; i < w * h; i++) {
        int pixel = i2p[i];
        float sx = 2.0f * ((float)(pixel % w) + 0.5f) / (float)w - 1.0f;
        float sy = 2.0f * ((float)(pixel / w) + 0.5f) / (float)h - 1.0f;
        if (randomSeed != 0) {                      // RayGenKernels.cu:99-110
            uint32_t a = randomSeed + (uint32_t)i, b = 0x9e3779b9u, c = 0x9e3779b9u;
            jenkins_mix(a, b, c);
            jenkins_mix(a, b, c);
            float ox = (float)((double)(float)a * 2.3283064365386963e-10);
            float oy = (float)((double)(float)b * 2.3283064365386963e-10);
            sx = sx + ox * 0.005f;
            sy = sy + oy * 0.005f;
        }
        float v[4] = {sx, sy, 0.0f, 1.0f}, r[4];
        for (int a = 0; a < 4; a++) { float rr = 0.0f; for (int b = 0; b < 4; b++) rr += n2w.m[a][b] * v[b]; r[a] = rr; }
        V3 world = V3(r[0] / r[3], r[1] / r[3], r[2] / r[3]);
        Ray& ray = rays[i];
        if (slotToID) slotToID[i] = pixel;
        if (idToSlot) idToSlot[pixel] = i;
        ray.o = origin;
        ray.d = normalize(world - origin);
        ray.tmin = 0.0f;
        ray.tmax = maxDist;
    }
}

void raygen_ao(Ray* outRays, int32_t* outIDToSlot, int32_t* outSlotToID,
               const Ray* inRays, const RayResult* inResults, const V3* normals,
               int firstInputSlot, int numInputRays, int numSamples, float maxDist, uint32_t randomSeed)
{
    const float PI = 3.14159265358979323846f;
    for (int task = 0; task < numInputRays; task++) {
        int inSlot = task + firstInputSlot;
        const Ray& inRay = inRays[inSlot];
        const RayResult& inRes = inResults[inSlot];
        int outSlot = task * numSamples;

        float epsilon = 1.0e-4f;
        V3 origin = inRay.o + inRay.d * std::fmax(inRes.t - epsilon, 0.0f);

        int tri = inRes.id;
        V3 normal(1.0f, 0.0f, 0.0f);
        if (tri != -1) normal = normals[tri];
        if (dot(normal, inRay.d) > 0.0f) normal = -normal;

        V3 na(std::fabs(normal.x), std::fabs(normal.y), std::fabs(normal.z));
        float nm = std::fmax(std::fmax(na.x, na.y), na.z);
        V3 perp(normal.y, -normal.x, 0.0f);
        if (nm == na.z) perp = V3(0.0f, normal.z, -normal.y);
        else if (nm == na.x) perp = V3(-normal.z, 0.0f, normal.x);
        perp = normalize(perp);
        V3 biperp = cross(normal, perp);

        uint32_t a = randomSeed + (uint32_t)task, b = 0x9e3779b9u, c = 0x9e3779b9u;
        jenkins_mix(a, b, c);
        jenkins_mix(a, b, c);
        float angle = (float)((double)(2.0f * PI * (float)c) * 2.3283064365386963e-10);

        V3 t0 = perp * std::cos(angle) + biperp * std::sin(angle);
        V3 t1 = perp * -std::sin(angle) + biperp * std::cos(angle);

        for (int i = 0; i < numSamples; i++) {
            float x = 0.0f, xadd = 1.0f;
            unsigned hc2 = (unsigned)i + 1;
            while (hc2 != 0) { xadd *= 0.5f; if (hc2 & 1) x += xadd; hc2 >>= 1; }
            float y = 0.0f, yadd = 1.0f;
            int hc3 = i + 1;
            while (hc3 != 0) { yadd *= 1.0f / 3.0f; y += (float)(hc3 % 3) * yadd; hc3 /= 3; }

            float ang = 2.0f * PI * y;
            float r = std::sqrt(x);
            x = r * std::cos(ang);
            y = r * std::sin(ang);
            float z = std::sqrt(1.0f - x * x - y * y);

            Ray& o = outRays[outSlot + i];
            o.o = origin;
            o.d = normalize(t0 * x + t1 * y + normal * z);
            o.tmin = 0.0f;
            o.tmax = (tri == -1) ? -1.0f : maxDist;
            if (outIDToSlot) outIDToSlot[outSlot + i] = i + outSlot;
            if (outSlotToID) outSlotToID[outSlot + i] = i + outSlot;
        }
    }
}

// rayGenShadowKernel — RayGenKernels.cu:47-73 (hammersley, sobol2D), 240-302
void raygen_shadow(Ray* outRays, int32_t* outIDToSlot, int32_t* outSlotToID, const Ray* inRays, const RayResult* inResults,
                   int firstInputSlot, int numInputRays, int numSamples, V3 lightPos, float lightRadius, uint32_t randomSeed)
{
    const float k32 = 2.3283064365386963e-10f;     // exp2(-32): scaling by a power of two is exact, so float == the kernel's double product
    for (int task = 0; task < numInputRays; task++) {
        int inSlot = task + firstInputSlot;
        const Ray& inRay = inRays[inSlot];
        const RayResult& inRes = inResults[inSlot];
        int outSlot = task * numSamples;
        float epsilon = 1.0e-2f;
        V3 origin = inRay.o + inRay.d * std::fmax(inRes.t - epsilon, 0.0f);
        uint32_t a = randomSeed + (uint32_t)task, b = 0x9e3779b9u, c = 0x9e3779b9u;
        jenkins_mix(a, b, c);
        jenkins_mix(a, b, c);
        V3 offset((float)a * k32, (float)b * k32, (float)c * k32);
        int tri = inRes.id;
        for (int i = 0; i < numSamples; i++) {
            uint32_t r1 = 0, r2 = 0;
            {
                uint32_t v1 = 1u << 31, v2 = 3u << 30;
                for (int j = i; j; j >>= 1) {
                    if (j & 1) { r1 ^= v1; r2 ^= v2 << 1; }
                    v1 |= v1 >> 1;
                    v2 ^= v2 >> 1;
                }
            }
            float p[3] = {(float)r1 * k32, (float)r2 * k32, ((float)i + 0.5f) / (float)numSamples};
            float o[3] = {offset.x, offset.y, offset.z};
            for (int k = 0; k < 3; k++) {
                p[k] = p[k] + o[k];
                if (p[k] >= 1.0f) p[k] -= 1.0f;
                p[k] = p[k] * 2.0f - 1.0f;
            }
            V3 target = lightPos + V3(p[0], p[1], p[2]) * lightRadius;
            V3 direction = target - origin;
            Ray& o2 = outRays[outSlot + i];
            o2.o = origin;
            o2.d = normalize(direction);
            o2.tmin = 0.0f;
            o2.tmax = (tri == -1) ? -1.0f : length(direction);
            if (outIDToSlot) outIDToSlot[outSlot + i] = i + outSlot;
            if (outSlotToID) outSlotToID[outSlot + i] = i + outSlot;
        }
    }
}

int count_hits(const RayResult* results, int n)
{
    int c = 0;
    for (int i = 0; i < n; i++) c += (results[i].id >= 0) ? 1 : 0;
    return c;
}

void tri_normals(const Scene& sc, V3* out)
{
    for (int t = 0; t < sc.numTris; t++) {
        V3 v0 = sc.v(t, 0), v1 = sc.v(t, 1), v2 = sc.v(t, 2);
        out[t] = normalize(cross(v1 - v0, v2 - v0));   // Scene.cpp:112
    }
}

} // namespace orc

// ---- RayBuffer::mortonSort restated (RayBuffer.cpp:88-163, RayBufferKernels.cu:62-196) ---------------------------
namespace orc {

static inline uint32_t f2u_sat(float f)
{
    if (!(f > 0.0f)) return 0u;                    // NaN and negatives -> 0 (device conversion semantics)
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}

void ray_morton_keys(const Ray* rays, int n, uint32_t* keys6 /* n*6 */, float aabb[6])
{
    V3 lo(F32_MAX, F32_MAX, F32_MAX), hi(-F32_MAX, -F32_MAX, -F32_MAX);
    for (int i = 0; i < n; i++) {                  // findAABBKernel: origins and end points
        V3 o = rays[i].o;
        lo = vmin(lo, o); hi = vmax(hi, o);
        V3 e = o + rays[i].d * rays[i].tmax;
        lo = vmin(lo, e); hi = vmax(hi, e);
    }
    aabb[0] = lo.x; aabb[1] = lo.y; aabb[2] = lo.z; aabb[3] = hi.x; aabb[4] = hi.y; aabb[5] = hi.z;
    for (int i = 0; i < n; i++) {                  // genMortonKeysKernel
        V3 a = (rays[i].o - lo) / (hi - lo);
        V3 nd = normalize(rays[i].d);
        V3 b = V3((nd.x + 1.0f) * 0.5f, (nd.y + 1.0f) * 0.5f, (nd.z + 1.0f) * 0.5f);
        uint32_t comp[6] = {f2u_sat(a.x * 256.0f * 65536.0f), f2u_sat(a.y * 256.0f * 65536.0f), f2u_sat(a.z * 256.0f * 65536.0f),
                            f2u_sat(b.x * 32.0f * 65536.0f), f2u_sat(b.y * 32.0f * 65536.0f), f2u_sat(b.z * 32.0f * 65536.0f)};
        uint32_t* h = keys6 + (size_t)i * 6;
        for (int k = 0; k < 6; k++) h[k] = 0;
        for (int k = 0; k < 6; k++)               // collectBits
            for (int bit = 0; bit < 32; bit++) {
                int pos = k + bit * 6;
                h[pos >> 5] |= ((comp[k] >> bit) & 1u) << (pos & 31);
            }
    }
}

// order[new] = old slot.  truncated == 0: full 192-bit comparator of compareMortonKey (hash[5] most significant), ties by
// old slot (the reference's quicksort leaves ties unspecified); truncated != 0: only key bits [83,147) compared.
void ray_morton_order(const Ray* rays, int n, int truncated, int32_t* order, uint64_t* key64Out)
{
    std::vector<uint32_t> keys((size_t)n * 6);
    float aabb[6];
    ray_morton_keys(rays, n, keys.data(), aabb);
    std::vector<uint64_t> k64(n);
    for (int i = 0; i < n; i++) {
        const uint32_t* h = &keys[(size_t)i * 6];
        uint64_t v = 0;
        for (int pos = 83; pos < 147; pos++) v |= (uint64_t)((h[pos >> 5] >> (pos & 31)) & 1u) << (pos - 83);
        k64[i] = v;
    }
    for (int i = 0; i < n; i++) order[i] = i;
    if (truncated)
        std::stable_sort(order, order + n, [&](int a, int b) { return k64[a] < k64[b]; });
    else
        std::stable_sort(order, order + n, [&](int a, int b) {
            const uint32_t* x = &keys[(size_t)a * 6]; const uint32_t* y = &keys[(size_t)b * 6];
            for (int w = 5; w >= 0; w--) if (x[w] != y[w]) return x[w] < y[w];
            return false;
        });
    if (key64Out) for (int i = 0; i < n; i++) key64Out[i] = k64[i];
}

} // namespace orc
