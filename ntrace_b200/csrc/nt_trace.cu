// ntrace_b200 — BVH traversal kernels for sm_100a (B200).
//
// Replaces the reference's runtime-compiled trace_bvh kernels
//   src/rt/kernels/fermi_speculative_while_while.cu:54-263   (semantics: speculative while-while, Compact)
//   src/rt/kernels/kepler_dynamic_fetch.cu:61-321            (semantics: persistent warps, dynamic fetch)
// behind the kernel ABI of src/rt/kernels/CudaTracerKernels.hpp:99-112 (TRACE_FUNC_BVH).
//
// Design (B200-first, no texture references, no video-instruction min/max, no local-memory stack):
//  * persistent grid = SMs x resident CTAs; each warp pulls rays from one global counter with a
//    single atomicAdd per refill (ballot + popc rank, leader broadcast with shfl);
//  * nodes are 64 B (two child AABBs + two child links): fetched as 2 x 256-bit ld.global.nc per lane (4 x 128-bit
//    for a buffer that is not 64-byte aligned); Woop triangles as up to 3 x ld.global.nc.v4 with the reference's early-outs;
//  * traversal stack: top entries in shared memory laid out [entry][thread] (bank = lane, conflict
//    free for any mix of depths), remainder spills to a per-thread local array (rarely touched);
//  * slab test on FMNMX3 (3-input min/max exists on sm_100a);
//  * speculative traversal: the first leaf a lane finds is postponed until no lane of the warp is
//    still searching (reference: fermi...cu:176-187);
//  * results are written once per ray when it terminates.
//
// Numerics: the ray/triangle test reproduces the host reference's Intersect::RayTriangleWoop
// (src/rt/Util.cpp:99-127) operation for operation in IEEE fp32 (no FMA contraction, IEEE reciprocal), so
// per-triangle hit decisions and the reported t/u/v are bit-identical to the CPU path traversing the same
// buffers.  The slab test keeps the reference GPU form n*idir - ood (fermi...cu:120-145) with FMA: it only
// decides which nodes are visited, never what a hit is.  FAST selects the other arithmetic the reference has: what
// nvcc -use_fast_math makes of its GPU kernels (contracted FMAs, approximate reciprocal), bit-identical to those kernels.
#include "nt_common.cuh"
#include <cstdlib>

namespace nt {

namespace {

constexpr int kStackSize = 96;                 // reference: STACK_SIZE 64 (fermi...cu:41); SAH/Split trees go 64 deep, keep headroom
constexpr int kDynamicFetchThreshold = 20;     // reference: kepler_dynamic_fetch.cu:43

__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// 256-bit global loads (new on sm_100: LDG.E.256): one request per 32-byte ray / half node instead of two
__device__ __forceinline__ void ld256_nc(const float4* p, float4& a, float4& b)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void ld256_cs(const float4* p, float4& a, float4& b)
{
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

// span begin = max(min(a0,a1), min(b0,b1), min(c0,c1), d); span end = min(max.., d)
__device__ __forceinline__ float span_begin(float a0, float a1, float b0, float b1, float c0, float c1, float d)
{
    return fmax3(fminf(a0, a1), fminf(b0, b1), fmaxf(fminf(c0, c1), d));
}
__device__ __forceinline__ float span_end(float a0, float a1, float b0, float b1, float c0, float c1, float d)
{
    return fmin3(fmaxf(a0, a1), fmaxf(b0, b1), fminf(fmaxf(c0, c1), d));
}

template <int LAYOUT>
__device__ __forceinline__ const float4* node_ptr(const float4* nodes, int addr)
{
    // Compact: byte offset; Compact2: float4 index (CudaBVH.cpp:86,614)
    if (LAYOUT == Layout_Compact2) return nodes + addr;
    return reinterpret_cast<const float4*>(reinterpret_cast<const char*>(nodes) + addr);
}

// BULKRAYS: TMA experiment on the warp ray fetch (north_star: "TMA/shared-memory staging where it measurably helps").  The n <= 32 rays a
// warp reserves are contiguous (n * 32 bytes): the reserving lane starts ONE bulk asynchronous copy (cp.async.bulk global -> shared,
// completion on a per-warp mbarrier) into a 1 KB per-warp staging area and every lane then reads its ray from shared memory, instead
// of one 256-bit streaming load per lane.  NT_TRACE_BULKRAYS=1 selects it.  Measured on the bench frame (scripts/tma_experiments.sh,
// profiles/r2i_tma_experiments.txt): primary 4 713 vs 4 759, AO 3 455 vs 3 874, diffuse 2 103 vs 2 364 Mrays/s, results identical -- the per-lane
// loads of contiguous rays already coalesce into full 128-byte lines, the copy adds an mbarrier round trip on the critical path of every
// refill and its 4 KB per CTA come out of the L1 that serves the BVH -- so the default stays the direct loads.
template <int LAYOUT, int BLOCK, int SMEM_N, bool PERSISTENT, bool FAST, bool WIDE, bool BULKRAYS = false>
__global__ void __launch_bounds__(BLOCK)
trace_kernel(int numRays, int anyHit, int fetchThreshold,
             const float4* __restrict__ rays, int4* __restrict__ results,
             const float4* __restrict__ nodes, const float4* __restrict__ woop,
             const int* __restrict__ triIndices, int* __restrict__ warpCounter, int* __restrict__ errorFlag,
             const BatchTable* __restrict__ batches)
{
    static_assert(SMEM_N >= 0 && SMEM_N <= kStackSize, "stack split");
    __shared__ int s_stack[(SMEM_N > 0 ? SMEM_N : 1) * BLOCK];
    int l_stack[(kStackSize > SMEM_N) ? (kStackSize - SMEM_N) : 1];

    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;
    int* const sbase = s_stack + tid;
    __shared__ __align__(128) float4 s_rays[BULKRAYS ? (BLOCK / 32) * 64 : 1];
    __shared__ __align__(8) unsigned long long s_bar[BULKRAYS ? BLOCK / 32 : 1];
    unsigned bulkPhase = 0;
    if (BULKRAYS) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((unsigned)__cvta_generic_to_shared(&s_bar[tid >> 5])) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }

    // the bound check sits on the spill path only (taken on ~1e-4 of the pushes): a tree deeper than the stack drops the entry and
    // raises the error flag instead of writing past l_stack
#define NT_PUSH(v)  do { ++sp; if (sp < SMEM_N) sbase[sp * BLOCK] = (v); else if (sp < kStackSize) l_stack[sp - SMEM_N] = (v); else { --sp; *(volatile int*)errorFlag = 1; } } while (0)
#define NT_POP(dst) do { (dst) = (sp < SMEM_N) ? sbase[sp * BLOCK] : l_stack[sp - SMEM_N]; --sp; } while (0)

    // Live state (registers).
    int   rayidx = -1;
    float origx = 0, origy = 0, origz = 0, dirx = 0, diry = 0, dirz = 0, tmin = 0;
    float idirx = 0, idiry = 0, idirz = 0, oodx = 0, oody = 0, oodz = 0;
    int   sp = 0;
    int   leafAddr = 0;
    int   nodeAddr = kEntrypointSentinel;
    int   hitIndex = -1;
    float hitT = 0, hitU = 0, hitV = 0;
    bool  alive = true;

    for (;;) {
        // ---------------- ray fetch ----------------
        const bool need = alive && (nodeAddr == kEntrypointSentinel);
        int fetchRank = 0;
        if (PERSISTENT) {
            // One atomicAdd per warp refill: rank the lanes that need a ray, the first of them
            // reserves a contiguous range and broadcasts its base (reference: kepler...cu:96-115).
            const unsigned needMask = __ballot_sync(0xffffffffu, need);
            if (needMask) {
                const int n = __popc(needMask);
                const int rank = __popc(needMask & ((1u << lane) - 1u));
                const int leader = __ffs(needMask) - 1;
                int base = 0;
                if ((int)lane == leader) base = atomicAdd(warpCounter, n);
                base = __shfl_sync(0xffffffffu, base, leader);
                if (need) rayidx = base + rank;
                fetchRank = rank;
                if (BULKRAYS) {
                    const int cnt = min(n, numRays - base);                 // warp-uniform
                    if (cnt > 0) {
                        const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar[tid >> 5]);
                        if ((int)lane == leader) {
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic reads of the staging area before the async write
                            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(cnt * 32) : "memory");
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         :: "r"((unsigned)__cvta_generic_to_shared(&s_rays[(tid >> 5) * 64])), "l"(rays + (size_t)base * 2), "r"(cnt * 32), "r"(bar) : "memory");
                        }
                        unsigned ok;
                        do {
                            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                                         : "=r"(ok) : "r"(bar), "r"(bulkPhase) : "memory");
                        } while (!ok);
                        bulkPhase ^= 1u;
                    }
                }
            }
        } else {
            if (need) rayidx = (rayidx == -1) ? (int)(blockIdx.x * BLOCK + tid) : numRays;
        }
        if (need) {
            if (rayidx >= numRays) { alive = false; rayidx = -1; }
            else {
                float4 o, d;
                if (BULKRAYS) {
                    o = s_rays[(tid >> 5) * 64 + fetchRank * 2]; d = s_rays[(tid >> 5) * 64 + fetchRank * 2 + 1];
                }
                else {
                    const float4* rp = rays + rayidx * 2;
                    if (batches) { const int b = batch_of(batches, rayidx); rp = batches->rays[b] + (size_t)(rayidx - __ldg(&batches->start[b])) * 2; }
                    if (WIDE) ld256_cs(rp, o, d);                        // 32-byte aligned ray buffers (checked by the launcher)
                    else { o = __ldcs(rp + 0); d = __ldcs(rp + 1); }
                }
                origx = o.x; origy = o.y; origz = o.z; tmin = o.w;
                dirx = d.x; diry = d.y; dirz = d.z; hitT = d.w;
                const float ooeps = exp2f(-80.0f);                      // fermi...cu:94-98
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
            // Inner nodes, until every lane still in this loop holds a postponed leaf.
            while ((unsigned)nodeAddr < (unsigned)kEntrypointSentinel) {
                const float4* ptr = node_ptr<LAYOUT>(nodes, nodeAddr);
                float4 n0xy, n1xy, nz, cn;
                if (WIDE) {
                    ld256_nc(ptr, n0xy, n1xy);        // (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y), (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
                    ld256_nc(ptr + 2, nz, cn);        // (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z), (c0, c1, splitInfo, 0) as ints
                } else {
                    n0xy = __ldg(ptr + 0); n1xy = __ldg(ptr + 1); nz = __ldg(ptr + 2); cn = __ldg(ptr + 3);
                }
                int c0idx = __float_as_int(cn.x), c1idx = __float_as_int(cn.y);

                const float c0lox = n0xy.x * idirx - oodx, c0hix = n0xy.y * idirx - oodx;
                const float c0loy = n0xy.z * idiry - oody, c0hiy = n0xy.w * idiry - oody;
                const float c0loz = nz.x * idirz - oodz,   c0hiz = nz.y * idirz - oodz;
                const float c1loz = nz.z * idirz - oodz,   c1hiz = nz.w * idirz - oodz;
                const float c0min = span_begin(c0lox, c0hix, c0loy, c0hiy, c0loz, c0hiz, tmin);
                const float c0max = span_end  (c0lox, c0hix, c0loy, c0hiy, c0loz, c0hiz, hitT);
                const float c1lox = n1xy.x * idirx - oodx, c1hix = n1xy.y * idirx - oodx;
                const float c1loy = n1xy.z * idiry - oody, c1hiy = n1xy.w * idiry - oody;
                const float c1min = span_begin(c1lox, c1hix, c1loy, c1hiy, c1loz, c1hiz, tmin);
                const float c1max = span_end  (c1lox, c1hix, c1loy, c1hiy, c1loz, c1hiz, hitT);

                const bool trav0 = (c0max >= c0min);
                const bool trav1 = (c1max >= c1min);

                if (!trav0 && !trav1) {
                    NT_POP(nodeAddr);
                } else {
                    nodeAddr = trav0 ? c0idx : c1idx;
                    if (trav0 && trav1) {
                        if (c1min < c0min) { const int t = nodeAddr; nodeAddr = c1idx; c1idx = t; }
                        NT_PUSH(c1idx);
                    }
                }

                // First leaf => postpone and continue traversal.
                if (nodeAddr < 0 && leafAddr >= 0) {
                    leafAddr = nodeAddr;
                    NT_POP(nodeAddr);
                }

                // All lanes have found a leaf => process them.
                if (!__any_sync(__activemask(), leafAddr >= 0)) break;
            }

            // Postponed leaves: Woop test against each triangle until the terminator.
            while (leafAddr < 0) {
                // Woop test in exactly the operation order of Intersect::RayTriangleWoop (Util.cpp:99-127), every
                // product and sum rounded separately (no FMA contraction) and an IEEE reciprocal, so that each
                // accept/reject decision is bit-identical to the host reference's.  Rows 1 and 2 of a triangle are fetched
                // only when needed (the reference's early-outs: least L1 traffic, fewest registers).
                int triAddr = ~leafAddr;
                float4 v00 = __ldg(woop + triAddr);
                for (;;) {
                    if (__float_as_int(v00.x) == (int)0x80000000) break;

                    float t;
                    if (FAST) {
                        // the arithmetic of the reference's GPU kernels as nvcc -use_fast_math compiles them: contracted FMAs, approximate 1/x
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
                // Another leaf was postponed => process it as well.
                leafAddr = nodeAddr;
                if (nodeAddr < 0) NT_POP(nodeAddr);
            }

            // Too few lanes left busy => leave and refill the idle ones (kepler...cu:310-311).
            if (PERSISTENT && __popc(__activemask()) < fetchThreshold) break;
        }

        // ---------------- result ----------------
        if (rayidx >= 0 && nodeAddr == kEntrypointSentinel) {
            int id = hitIndex;
            if (id != -1) id = __ldg(triIndices + id);          // fermi...cu:260-261
            int4* op = results + rayidx;
            if (batches) { const int b = batch_of(batches, rayidx); op = batches->results[b] + (rayidx - __ldg(&batches->start[b])); }
            __stcs(op, make_int4(id, __float_as_int(hitT), __float_as_int(hitU), __float_as_int(hitV)));
            if (PERSISTENT) rayidx = -1;
        }
        if (!PERSISTENT) break;
    }
#undef NT_PUSH
#undef NT_POP
}

constexpr int kBlock = 128;

// Tuning knobs (defaults chosen from the ncu captures in profiles/ and the sweep in scripts/tune_trace.sh; the NT_TRACE_*
// environment variables exist for experiments only): entries of the traversal stack kept in shared memory, the
// shared-memory carveout that decides how many CTAs fit next to the L1, and the dynamic-fetch threshold.
struct Tuning { int smemStack; int carveout; int fetchThreshold; };
Tuning tuning()
{
    static Tuning t = [] {
        Tuning r{8, 20, kDynamicFetchThreshold};
        if (const char* e = getenv("NT_TRACE_SMEM")) r.smemStack = atoi(e);
        if (const char* e = getenv("NT_TRACE_CARVEOUT")) r.carveout = atoi(e);
        if (const char* e = getenv("NT_TRACE_FETCH")) r.fetchThreshold = atoi(e);
        return r;
    }();
    return t;
}

template <int LAYOUT, int SMEM_N, bool PERSISTENT, bool FAST, bool WIDE, bool BULKRAYS = false>
cudaError_t launch_variant(const TraceLaunch& a, int* launches)
{
    auto kern = trace_kernel<LAYOUT, kBlock, SMEM_N, PERSISTENT, FAST, WIDE, BULKRAYS>;
    static int blocksPerSM = 0, epoch = -1;
    if (epoch != launch_epoch()) {
        epoch = launch_epoch();
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, tuning().carveout);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kern, kBlock, 0);
        if (e != cudaSuccess) return e;
        if (blocksPerSM < 1) blocksPerSM = 1;
    }
    int grid = (a.numRays + kBlock - 1) / kBlock;
    if (PERSISTENT && grid > a.numSMs * blocksPerSM) grid = a.numSMs * blocksPerSM;   // one resident wave
    if (a.batches && !PERSISTENT) return cudaErrorInvalidValue;                         // a batch table needs the global ray counter
    kern<<<grid, kBlock, 0, a.stream>>>(a.numRays, a.anyHit, tuning().fetchThreshold, a.rays, a.results, a.nodes, a.woop, a.triIndices, a.warpCounter, a.errorFlag, a.batches);
    if (launches) *launches = 1;
    return cudaGetLastError();
}

template <int LAYOUT, bool PERSISTENT, bool FAST, bool WIDE>
cudaError_t launch_stack(const TraceLaunch& a, int* launches)
{
    switch (tuning().smemStack) {            // 8 unless an experiment asks otherwise
    case 0:  return launch_variant<LAYOUT, 0, PERSISTENT, FAST, WIDE>(a, launches);
    case 16: return launch_variant<LAYOUT, 16, PERSISTENT, FAST, WIDE>(a, launches);
    default:
        if (PERSISTENT && WIDE && !FAST && LAYOUT == Layout_Compact) {
            static const bool bulkRays = [] { const char* e = getenv("NT_TRACE_BULKRAYS"); return e && atoi(e) != 0; }();
            if (bulkRays) return launch_variant<LAYOUT, 8, PERSISTENT, FAST, WIDE, (PERSISTENT && WIDE && !FAST && LAYOUT == Layout_Compact)>(a, launches);
        }
        return launch_variant<LAYOUT, 8, PERSISTENT, FAST, WIDE>(a, launches);
    }
}

template <int LAYOUT, bool PERSISTENT>
cudaError_t launch_one(const TraceLaunch& a, int* launches)
{
    // 256-bit loads need a 32-byte aligned ray buffer and 64-byte aligned nodes (true for everything nt_mem_alloc returns)
    // (the ray buffers of a batch table are checked by nt_trace_batches)
    const bool wide = (a.batches || (reinterpret_cast<size_t>(a.rays) & 31) == 0) && (reinterpret_cast<size_t>(a.nodes) & 63) == 0;
    if (a.fast) return wide ? launch_stack<LAYOUT, PERSISTENT, true, true>(a, launches) : launch_stack<LAYOUT, PERSISTENT, true, false>(a, launches);
    return wide ? launch_stack<LAYOUT, PERSISTENT, false, true>(a, launches) : launch_stack<LAYOUT, PERSISTENT, false, false>(a, launches);
}

} // namespace

static int g_launchEpoch = 0;
int launch_epoch() { return g_launchEpoch; }
void reset_launch_caches() { ++g_launchEpoch; }

KernelConfig trace_kernel_config(int kernel, int layout)
{
    // reference: queryConfig() in every kernel file (e.g. fermi...cu:45-50)
    KernelConfig c;
    c.bvhLayout = layout;
    c.blockWidth = 32;
    c.blockHeight = kBlock / 32;
    c.usePersistentThreads = (kernel == Kernel_PlainSpeculative) ? 0 : 1;        // Kernel_Auto: both of its kernels are persistent
    return c;
}

cudaError_t launch_trace(const TraceLaunch& a, int* launches)
{
    if (a.numRays <= 0) { if (launches) *launches = 0; return cudaSuccess; }
    const bool c2 = (a.layout == Layout_Compact2);
    switch (a.kernel) {
    case Kernel_PersistentSpeculative:
        return c2 ? launch_one<Layout_Compact2, true>(a, launches) : launch_one<Layout_Compact, true>(a, launches);
    case Kernel_PlainSpeculative:
        return c2 ? launch_one<Layout_Compact2, false>(a, launches) : launch_one<Layout_Compact, false>(a, launches);
    case Kernel_Wide4Persistent:
        return launch_trace_wide4(a, launches);
    case Kernel_BinaryMr:
    case Kernel_Wide4Mr:
        if (a.batches) return cudaErrorInvalidValue;
        return launch_trace_mr(a, launches);
    case Kernel_BinarySw:
    case Kernel_Wide4Sw:
        if (a.batches) return cudaErrorInvalidValue;
        return launch_trace_sw(a, launches);
    default:
        return cudaErrorInvalidValue;
    }
}

} // namespace nt
