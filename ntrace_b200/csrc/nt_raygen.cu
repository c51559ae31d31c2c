// ntrace_b200 — ray generation and hit accounting kernels for sm_100a.
//
// Replaces (same outputs, same slot order):
//   src/rt/ray/RayGenKernels.cu:77-125   rayGenPrimaryKernel  (+ PixelTable lookup, PixelTable.cpp:57-141)
//   src/rt/ray/RayGenKernels.cu:129-236  rayGenAOKernel       (AO and, with maxDist = camera far, diffuse)
//   src/rt/cuda/RendererKernels.cu:174-224 countHitsKernel
//   src/rt/Scene.cpp:112                 per-triangle normals
// One thread per *output* ray so that the 32-byte ray stores of a warp are contiguous.
#include "nt_common.cuh"

namespace nt {

namespace {

__device__ __forceinline__ void jenkins_mix(unsigned& a, unsigned& b, unsigned& c)
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

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 mk(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ float dot3(F3 a, F3 b) { float r = 0.0f; r = __fadd_rn(r, __fmul_rn(a.x, b.x)); r = __fadd_rn(r, __fmul_rn(a.y, b.y)); r = __fadd_rn(r, __fmul_rn(a.z, b.z)); return r; }
__device__ __forceinline__ F3 cross3(F3 a, F3 b)
{
    return mk(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)),
              __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
              __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
// v * rcp(length(v)) with the reference's rcp(0) == 0 convention (Math.hpp:114,142-144)
__device__ __forceinline__ F3 normalize3(F3 v)
{
    const float len = __fsqrt_rn(dot3(v, v));
    const float s = (len != 0.0f) ? __fdiv_rn(1.0f, len) : 0.0f;
    return mk(__fmul_rn(v.x, s), __fmul_rn(v.y, s), __fmul_rn(v.z, s));
}

struct PrimaryParams {
    float4* rays; int* idToSlot; int* slotToID; const int* indexToPixel;
    float ox, oy, oz; float m[16]; int w, h; float maxDist; unsigned seed;
};

__global__ void __launch_bounds__(256) raygen_primary_kernel(PrimaryParams p)
{
    const int task = blockIdx.x * blockDim.x + threadIdx.x;
    if (task >= p.w * p.h) return;
    const int pixel = __ldg(p.indexToPixel + task);

    // IEEE, non-contracted arithmetic: bit-identical to RayGen::primaryCPU (RayGen.cpp:76-112).
    float sx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)(pixel % p.w), 0.5f)), (float)p.w), 1.0f);
    float sy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)(pixel / p.w), 0.5f)), (float)p.h), 1.0f);
    if (p.seed != 0) {
        unsigned a = p.seed + (unsigned)task, b = 0x9e3779b9u, c = 0x9e3779b9u;
        jenkins_mix(a, b, c);
        jenkins_mix(a, b, c);
        const float jx = __fmul_rn((float)a, 2.3283064365386963e-10f);
        const float jy = __fmul_rn((float)b, 2.3283064365386963e-10f);
        sx = __fadd_rn(sx, __fmul_rn(jx, 0.005f));
        sy = __fadd_rn(sy, __fmul_rn(jy, 0.005f));
    }
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float rr = 0.0f;
        rr = __fadd_rn(rr, __fmul_rn(p.m[i * 4 + 0], sx));
        rr = __fadd_rn(rr, __fmul_rn(p.m[i * 4 + 1], sy));
        rr = __fadd_rn(rr, __fmul_rn(p.m[i * 4 + 2], 0.0f));
        rr = __fadd_rn(rr, __fmul_rn(p.m[i * 4 + 3], 1.0f));
        r[i] = rr;
    }
    const F3 world = mk(__fdiv_rn(r[0], r[3]), __fdiv_rn(r[1], r[3]), __fdiv_rn(r[2], r[3]));
    const F3 d = normalize3(mk(__fsub_rn(world.x, p.ox), __fsub_rn(world.y, p.oy), __fsub_rn(world.z, p.oz)));

    p.rays[task * 2 + 0] = make_float4(p.ox, p.oy, p.oz, 0.0f);
    p.rays[task * 2 + 1] = make_float4(d.x, d.y, d.z, p.maxDist);
    if (p.slotToID) p.slotToID[task] = pixel;
    if (p.idToSlot) p.idToSlot[pixel] = task;
}

__global__ void __launch_bounds__(256) raygen_ao_kernel(AOArgs a)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = a.numInputRays * a.numSamples;
    if (o >= total) return;
    const int task = o / a.numSamples;
    const int i = o - task * a.numSamples;
    const int inSlot = task + a.firstInputSlot;

    const float4 ro = __ldg(a.inRays + inSlot * 2 + 0);
    const float4 rd = __ldg(a.inRays + inSlot * 2 + 1);
    const int4 res = __ldg(a.inResults + inSlot);
    const float hitT = __int_as_float(res.y);
    const int tri = res.x;

    // origin, backed off along the ray (RayGenKernels.cu:151-152)
    const float back = fmaxf(__fsub_rn(hitT, 1.0e-4f), 0.0f);
    const F3 origin = mk(__fadd_rn(ro.x, __fmul_rn(rd.x, back)), __fadd_rn(ro.y, __fmul_rn(rd.y, back)), __fadd_rn(ro.z, __fmul_rn(rd.z, back)));

    F3 normal = mk(1.0f, 0.0f, 0.0f);
    if (tri != -1) normal = mk(__ldg(a.normals + tri * 3 + 0), __ldg(a.normals + tri * 3 + 1), __ldg(a.normals + tri * 3 + 2));
    if (dot3(normal, mk(rd.x, rd.y, rd.z)) > 0.0f) normal = mk(-normal.x, -normal.y, -normal.z);

    const F3 na = mk(fabsf(normal.x), fabsf(normal.y), fabsf(normal.z));
    const float nm = fmaxf(fmaxf(na.x, na.y), na.z);
    F3 perp = mk(normal.y, -normal.x, 0.0f);
    if (nm == na.z) perp = mk(0.0f, normal.z, -normal.y);
    else if (nm == na.x) perp = mk(-normal.z, 0.0f, normal.x);
    perp = normalize3(perp);
    const F3 biperp = cross3(normal, perp);

    unsigned ha = a.seed + (unsigned)task, hb = 0x9e3779b9u, hc = 0x9e3779b9u;
    jenkins_mix(ha, hb, hc);
    jenkins_mix(ha, hb, hc);
    const float PI = 3.14159265358979323846f;
    const float angle = __fmul_rn(__fmul_rn(__fmul_rn(2.0f, PI), (float)hc), 2.3283064365386963e-10f);
    float sa, ca;
    sincosf(angle, &sa, &ca);
    const F3 t0 = mk(__fadd_rn(__fmul_rn(perp.x, ca), __fmul_rn(biperp.x, sa)),
                     __fadd_rn(__fmul_rn(perp.y, ca), __fmul_rn(biperp.y, sa)),
                     __fadd_rn(__fmul_rn(perp.z, ca), __fmul_rn(biperp.z, sa)));
    const F3 t1 = mk(__fadd_rn(__fmul_rn(perp.x, -sa), __fmul_rn(biperp.x, ca)),
                     __fadd_rn(__fmul_rn(perp.y, -sa), __fmul_rn(biperp.y, ca)),
                     __fadd_rn(__fmul_rn(perp.z, -sa), __fmul_rn(biperp.z, ca)));

    // Halton(2,3) sample i (RayGenKernels.cu:194-215)
    float x = 0.0f, xadd = 1.0f;
    for (unsigned hc2 = (unsigned)i + 1; hc2 != 0; hc2 >>= 1) { xadd = __fmul_rn(xadd, 0.5f); if (hc2 & 1) x = __fadd_rn(x, xadd); }
    float y = 0.0f, yadd = 1.0f;
    for (int hc3 = i + 1; hc3 != 0; hc3 /= 3) { yadd = __fmul_rn(yadd, 1.0f / 3.0f); y = __fadd_rn(y, __fmul_rn((float)(hc3 % 3), yadd)); }

    const float ang = __fmul_rn(__fmul_rn(2.0f, PI), y);
    const float r = __fsqrt_rn(x);
    float s2, c2;
    sincosf(ang, &s2, &c2);
    x = __fmul_rn(r, c2);
    y = __fmul_rn(r, s2);
    const float z = __fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, __fmul_rn(x, x)), __fmul_rn(y, y)));

    const F3 dir = normalize3(mk(
        __fadd_rn(__fadd_rn(__fmul_rn(t0.x, x), __fmul_rn(t1.x, y)), __fmul_rn(normal.x, z)),
        __fadd_rn(__fadd_rn(__fmul_rn(t0.y, x), __fmul_rn(t1.y, y)), __fmul_rn(normal.y, z)),
        __fadd_rn(__fadd_rn(__fmul_rn(t0.z, x), __fmul_rn(t1.z, y)), __fmul_rn(normal.z, z))));

    a.outRays[o * 2 + 0] = make_float4(origin.x, origin.y, origin.z, 0.0f);
    a.outRays[o * 2 + 1] = make_float4(dir.x, dir.y, dir.z, (tri == -1) ? -1.0f : a.maxDist);
    if (a.outIDToSlot) a.outIDToSlot[o] = o;
    if (a.outSlotToID) a.outSlotToID[o] = o;
}

// rayGenShadowKernel (RayGenKernels.cu:240-302): numSamples rays from each hit point towards a spherical light,
// sample i = Cranley-Patterson rotated (Sobol2D(i), Hammersley(i)) point in the light's bounding cube.
__global__ void __launch_bounds__(256) raygen_shadow_kernel(ShadowArgs a)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = a.numInputRays * a.numSamples;
    if (o >= total) return;
    const int task = o / a.numSamples;
    const int i = o - task * a.numSamples;
    const int inSlot = task + a.firstInputSlot;

    const float4 ro = __ldg(a.inRays + inSlot * 2 + 0);
    const float4 rd = __ldg(a.inRays + inSlot * 2 + 1);
    const int4 res = __ldg(a.inResults + inSlot);
    const int tri = res.x;
    const float back = fmaxf(__fsub_rn(__int_as_float(res.y), 1.0e-2f), 0.0f);
    const F3 origin = mk(__fadd_rn(ro.x, __fmul_rn(rd.x, back)), __fadd_rn(ro.y, __fmul_rn(rd.y, back)), __fadd_rn(ro.z, __fmul_rn(rd.z, back)));

    unsigned ha = a.seed + (unsigned)task, hb = 0x9e3779b9u, hc = 0x9e3779b9u;
    jenkins_mix(ha, hb, hc);
    jenkins_mix(ha, hb, hc);
    const float k32 = 2.3283064365386963e-10f;                                  // 2^-32: exact scaling
    const float off[3] = {__fmul_rn((float)ha, k32), __fmul_rn((float)hb, k32), __fmul_rn((float)hc, k32)};

    unsigned r1 = 0, r2 = 0;
    {
        unsigned v1 = 1u << 31, v2 = 3u << 30;
        for (int j = i; j; j >>= 1) {
            if (j & 1) { r1 ^= v1; r2 ^= v2 << 1; }
            v1 |= v1 >> 1;
            v2 ^= v2 >> 1;
        }
    }
    float pos[3] = {__fmul_rn((float)r1, k32), __fmul_rn((float)r2, k32), __fdiv_rn(__fadd_rn((float)i, 0.5f), (float)a.numSamples)};
    const float light[3] = {a.lightPos[0], a.lightPos[1], a.lightPos[2]};
    float dirv[3];
    const float org[3] = {origin.x, origin.y, origin.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float p = __fadd_rn(pos[k], off[k]);
        if (p >= 1.0f) p = __fsub_rn(p, 1.0f);
        p = __fsub_rn(__fmul_rn(p, 2.0f), 1.0f);
        const float target = __fadd_rn(light[k], __fmul_rn(a.lightRadius, p));
        dirv[k] = __fsub_rn(target, org[k]);
    }
    const F3 d = mk(dirv[0], dirv[1], dirv[2]);
    const F3 dir = normalize3(d);
    a.outRays[o * 2 + 0] = make_float4(origin.x, origin.y, origin.z, 0.0f);
    a.outRays[o * 2 + 1] = make_float4(dir.x, dir.y, dir.z, (tri == -1) ? -1.0f : __fsqrt_rn(dot3(d, d)));
    if (a.outIDToSlot) a.outIDToSlot[o] = o;
    if (a.outSlotToID) a.outSlotToID[o] = o;
}

__global__ void __launch_bounds__(256) count_hits_kernel(const int4* __restrict__ results, int numRays, int* __restrict__ counter)
{
    int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < numRays; i += gridDim.x * blockDim.x)
        c += (__ldg(&results[i].x) >= 0) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(counter, c);
}

__global__ void __launch_bounds__(256) tri_normals_kernel(const float* __restrict__ verts, const int* __restrict__ tris, int numTris, float* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numTris) return;
    const int i0 = tris[t * 3 + 0], i1 = tris[t * 3 + 1], i2 = tris[t * 3 + 2];
    const F3 v0 = mk(verts[i0 * 3], verts[i0 * 3 + 1], verts[i0 * 3 + 2]);
    const F3 v1 = mk(verts[i1 * 3], verts[i1 * 3 + 1], verts[i1 * 3 + 2]);
    const F3 v2 = mk(verts[i2 * 3], verts[i2 * 3 + 1], verts[i2 * 3 + 2]);
    const F3 e1 = mk(__fsub_rn(v1.x, v0.x), __fsub_rn(v1.y, v0.y), __fsub_rn(v1.z, v0.z));
    const F3 e2 = mk(__fsub_rn(v2.x, v0.x), __fsub_rn(v2.y, v0.y), __fsub_rn(v2.z, v0.z));
    const F3 n = normalize3(cross3(e1, e2));
    out[t * 3 + 0] = n.x; out[t * 3 + 1] = n.y; out[t * 3 + 2] = n.z;
}

} // namespace

cudaError_t launch_raygen_primary(const PrimaryArgs& a, const int* indexToPixel, cudaStream_t s)
{
    PrimaryParams p;
    p.rays = a.rays; p.idToSlot = a.idToSlot; p.slotToID = a.slotToID; p.indexToPixel = indexToPixel;
    p.ox = a.origin[0]; p.oy = a.origin[1]; p.oz = a.origin[2];
    for (int i = 0; i < 16; i++) p.m[i] = a.n2w[i];
    p.w = a.w; p.h = a.h; p.maxDist = a.maxDist; p.seed = a.seed;
    const int n = a.w * a.h;
    if (n <= 0) return cudaSuccess;
    raygen_primary_kernel<<<(n + 255) / 256, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_raygen_ao(const AOArgs& a, cudaStream_t s)
{
    const long long n = (long long)a.numInputRays * a.numSamples;
    if (n <= 0) return cudaSuccess;
    raygen_ao_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_raygen_shadow(const ShadowArgs& a, cudaStream_t s)
{
    const long long n = (long long)a.numInputRays * a.numSamples;
    if (n <= 0) return cudaSuccess;
    raygen_shadow_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_count_hits(const int4* results, int numRays, int* dCounter, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(dCounter, 0, sizeof(int), s);
    if (e != cudaSuccess || numRays <= 0) return e;
    int grid = (numRays + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    count_hits_kernel<<<grid, 256, 0, s>>>(results, numRays, dCounter);
    return cudaGetLastError();
}

cudaError_t launch_tri_normals(const float* verts, const int* tris, int numTris, float* out, cudaStream_t s)
{
    if (numTris <= 0) return cudaSuccess;
    tri_normals_kernel<<<(numTris + 255) / 256, 256, 0, s>>>(verts, tris, numTris, out);
    return cudaGetLastError();
}

} // namespace nt
