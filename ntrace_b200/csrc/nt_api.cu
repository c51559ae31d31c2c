// ntrace_b200 — implementation of the C ABI declared in include/ntrace_b200.h.
//
// Host-side behaviour mirrors the reference tracer object:
//   CudaBVHTracer::setKernel / getDesiredBVHLayout / setBVH / traceBatch
//     (src/rt/cuda/CudaBVHTracer.cpp:52-168), errors as in :92-101 ("No BVH!", "Incorrect BVH layout!"),
//   FW::Buffer's lazy host<->device migration (src/framework/gpu/Buffer.hpp:107-113) becomes
//   explicit staging of host pointers through grow-only device buffers.
#include "../../include/ntrace_b200.h"
#include "nt_common.cuh"

#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <mutex>
#include <vector>

namespace nt {

static thread_local std::string t_error;

void set_error(const std::string& msg) { t_error = msg; }

bool check_cuda(cudaError_t e, const char* what, const char* file, int line)
{
    if (e == cudaSuccess) return true;
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    set_error(buf);
    cudaGetLastError();   // clear the sticky launch error so later calls can proceed
    return false;
}

cudaError_t DevBuf::reserve(size_t bytes)
{
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; cap = 0; return e; }
    cap = want;
    return cudaSuccess;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

namespace {

struct Context {
    bool inited = false;
    int device = 0, numSMs = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evA = nullptr, evB = nullptr;
    cudaEvent_t userEv[8] = {};
    // host-buffer pipeline: copy-in / copy-out streams and per-chunk events
    static constexpr int kMaxChunks = 16;
    cudaStream_t sIn = nullptr, sOut = nullptr;
    cudaEvent_t evIn[kMaxChunks] = {}, evKa[kMaxChunks] = {}, evKb[kMaxChunks] = {};
    bool deferred = false;
    // deferred mode 2: consecutive launches alternate between two kernel streams, so the CTAs of the next batch move in as the
    // persistent CTAs of the current one run out of rays (the tail of one launch overlaps the head of the next)
    bool overlap = false, overlapPending = false;
    int overlapNext = 0;
    static constexpr int kMaxKernelStreams = 4;
    int numKernelStreams = 2;            // deferred mode 2: launches alternate over this many kernel streams (NT_OVERLAP_STREAMS, experiment knob)
    cudaStream_t kStream[kMaxKernelStreams] = {};
    cudaEvent_t kDone[kMaxKernelStreams] = {}, kFork = nullptr;
    // what the overlapped launches still in flight read and write, so that a later call waits only for the launches whose buffers it
    // touches (a ray generator filling buffer B does not wait for the trace of buffer A)
    struct InFlight { const char* rLo = nullptr; const char* rHi = nullptr; const char* wLo = nullptr; const char* wHi = nullptr; cudaEvent_t ev = nullptr; bool live = false; };
    static constexpr int kRing = 8;
    InFlight ring[kRing];
    int ringNext = 0;
    // ray buffers written by a queued nt_raygen_ao (overlapped mode): the trace of such a buffer waits for ITS generator only, not for
    // everything the main stream has queued since (the generator of the batch after it, which sits behind the previous trace's tail)
    struct Producer { const char* lo = nullptr; const char* hi = nullptr; cudaEvent_t ev = nullptr; bool live = false; };
    Producer prod[kRing];
    int prodNext = 0;
    int64_t launches = 0;
    // asynchronous submission (nt_trace_batch_async / nt_trace_wait): per-slot staging + events
    static constexpr int kAsyncSlots = 4;
    struct AsyncSlot {
        DevBuf rays, results;
        cudaEvent_t evIn = nullptr, evKa = nullptr, evKb = nullptr, evOut = nullptr;
        bool busy = false, copyOut = false;
    } async[kAsyncSlots];

    int kernel = Kernel_PersistentSpeculative;
    int kernelLayout = Layout_Compact;
    bool fastMath = false;

    // resident BVH
    DevBuf nodes, woop, triIndex;
    size_t nodeBytes = 0, woopBytes = 0, idxBytes = 0;
    int bvhLayout = Layout_Max;
    bool haveBVH = false;
    uint64_t generation = 0;       // bumped whenever the resident BVH is replaced or re-laid-out (nt_bvh_generation)
    // a BVH uploaded in one of the basic layouts (AOS/SOA) is kept as given in src* (what download / device_ptrs /
    // broadcast see) and rewritten on the device into nodes/woop/triIndex (Compact form) before the first trace
    DevBuf srcNodes, srcWoop, srcIdx, layoutScratch;
    size_t srcNodeBytes = 0, srcWoopBytes = 0, srcIdxBytes = 0;
    bool basic = false, converted = false;
    // Wide4 form of the resident (Compact / Compact2) node buffer, derived on demand for the b200_wide4* kernels (nt_wide.cu)
    DevBuf wideNodes, wideScratch;
    int raygenOrder = 0;                 // nt_raygen_set_order
    // nt_trace_batches: device tables of the launches in flight, a ring per stream (main stream: index kMaxKernelStreams); a table is
    // written by a copy on the stream of the launch that reads it, so a slot is only reused behind the launch that used it before
    static constexpr int kBatchTableSlots = 8;
    DevBuf batchTables; int batchTableNext[kMaxKernelStreams + 1] = {};
    // b200_auto: the device buffer the library's own primary-ray generator wrote last.  Camera rays share one origin and stay together
    // far down the tree, where the cheaper binary node step wins over Wide4 (4.6 vs 4.1 Grays/s on the bench frame); a stale range can
    // only cost speed, never change a result.
    const char* primLo = nullptr; const char* primHi = nullptr;
    size_t wideBytes = 0;
    bool wideValid = false;
    int wideDepth = 0;
    DevBuf sortedKeys, sortedIdx;
    int builtTris = 0;
    int collapseMode = 0, collapseMaxLeaf = 0;
    int buildLayout = Layout_Compact;

    // staging
    DevBuf stRays, stResults, stA, stB, stC, stD, stE;
    DevBuf counters;          // [0] warp counter, [1] hit counter
    int* errHost = nullptr;   // mapped pinned word the trace kernels OR 1 into on a traversal-stack overflow (errDev: its device address)
    int* errDev = nullptr;
    DevBuf pixelTable; int ptW = 0, ptH = 0;
    DevBuf sceneVerts, sceneTris;
};

Context g;
std::mutex g_mutex;

bool is_device_ptr(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// address the device can dereference for a caller buffer: device/managed memory as is, pinned host memory through its
// UVA mapping; nullptr for pageable host memory
template <class T> T* mapped_device_ptr(T* p)
{
    if (!p) return nullptr;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, (const void*)p);
    if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return p;
    if (at.type == cudaMemoryTypeHost && at.devicePointer) return (T*)at.devicePointer;
    return nullptr;
}

// device view of a caller buffer that the call only reads
int stage_in(const void* src, size_t bytes, DevBuf& st, const void** out)
{
    if (bytes == 0 || !src) { *out = nullptr; return 0; }
    if (is_device_ptr(src)) { *out = src; return 0; }
    NT_CUDA(st.reserve(bytes));
    NT_CUDA(cudaMemcpyAsync(st.p, src, bytes, cudaMemcpyHostToDevice, g.stream));
    *out = st.p;
    return 0;
}
// device view of a caller buffer that the call writes; `*hostDst` is set when a copy-back is due
int stage_out(void* dst, size_t bytes, DevBuf& st, void** out, void** hostDst)
{
    *hostDst = nullptr;
    if (bytes == 0 || !dst) { *out = nullptr; return 0; }
    if (is_device_ptr(dst)) { *out = dst; return 0; }
    NT_CUDA(st.reserve(bytes));
    *out = st.p;
    *hostDst = dst;
    return 0;
}
int copy_back(void* hostDst, const void* dev, size_t bytes)
{
    if (hostDst && bytes) NT_CUDA(cudaMemcpyAsync(hostDst, dev, bytes, cudaMemcpyDeviceToHost, g.stream));
    return 0;
}

bool is_basic_layout(int layout) { return layout >= Layout_AOS_AOS && layout <= Layout_SOA_SOA; }

const char* link_range_error(int layout, size_t nodeBytes, size_t woopBytes)
{
    if (layout == Layout_Compact && nodeBytes >= (size_t)kEntrypointSentinel)
        return "ntrace_b200: BVHLayout_Compact stores child links as 32-bit byte offsets below 0x76543210: node buffers of 1.98 GB or more need BVHLayout_Compact2";
    if (layout == Layout_Compact2 && nodeBytes / 16 >= (size_t)kEntrypointSentinel)
        return "ntrace_b200: node buffer too large for 32-bit BVHLayout_Compact2 links";
    if (woopBytes / 16 >= 0x7fffffffull)
        return "ntrace_b200: triangle buffer too large for 32-bit leaf links";
    return nullptr;
}

// (re)build the traversal form of a basic-layout BVH; caller holds the mutex
int ensure_traversal_form()
{
    if (!g.basic || g.converted) return 0;
    BuildOutput out;
    out.nodes = &g.nodes; out.woop = &g.woop; out.triIndex = &g.triIndex;
    out.sortedKeys = out.sortedIdx = nullptr;
    out.nodeBytes = out.woopBytes = out.idxBytes = 0;
    int launches = 0;
    std::string err;
    cudaError_t e = convert_basic_layout(g.bvhLayout, g.srcNodes.p, g.srcNodeBytes, g.srcWoop.p, g.srcWoopBytes, g.srcIdx.as<int>(), g.srcIdxBytes,
                                         Layout_Compact, out, g.layoutScratch, g.stream, &launches, &err);
    g.launches += launches;
    if (e != cudaSuccess) {
        if (!err.empty()) { set_error("ntrace_b200: " + err); cudaGetLastError(); return 1; }
        NT_CUDA(e);
    }
    g.nodeBytes = out.nodeBytes; g.woopBytes = out.woopBytes; g.idxBytes = out.idxBytes;
    g.converted = true;
    return 0;
}

// derive the Wide4 node array from the resident Compact / Compact2 nodes, on the device (nt_wide.cu); caller holds the mutex
int ensure_wide_form()
{
    if ((g.kernel != Kernel_Wide4Persistent && g.kernel != Kernel_Wide4Mr && g.kernel != Kernel_Wide4Sw && g.kernel != Kernel_Auto) || g.wideValid) return 0;
    std::string err;
    const int srcLayout = g.basic ? (int)Layout_Compact : g.bvhLayout;
    int launches = 0;
    cudaError_t e = convert_compact_to_wide4_device(g.nodes.p, g.nodeBytes, srcLayout, g.woopBytes / 16, g.wideNodes, g.wideScratch, &g.wideBytes,
                                                    &g.wideDepth, g.numSMs, g.stream, &launches, &err);
    g.launches += launches;
    if (e != cudaSuccess) {
        if (!err.empty()) { set_error("ntrace_b200: " + err); cudaGetLastError(); return 1; }
        NT_CUDA(e);
    }
    g.wideValid = true;
    return 0;
}

// b200_auto: any-hit batches -> binary kernel (the reference's visiting order), closest-hit batches -> Wide4, except the camera rays the
// library generated itself (see Context::primLo)
inline int auto_kernel(const void* raysDev, int numRays, int needClosestHit)
{
    if (g.kernel != Kernel_Auto) return g.kernel;
    if (!needClosestHit) return Kernel_PersistentSpeculative;
    const char* p = (const char*)raysDev;
    if (p && g.primLo && p >= g.primLo && p + (size_t)numRays * 32 <= g.primHi) return Kernel_PersistentSpeculative;
    return Kernel_Wide4Persistent;
}

// every entry point but an overlapped deferred launch orders the main stream behind the launches still in flight on the two
// kernel streams (deferred mode 2), so that whatever it enqueues or waits for sees their results
int join_kernel_streams()
{
    if (!g.overlapPending) return 0;
    for (int k = 0; k < g.numKernelStreams; k++) NT_CUDA(cudaStreamWaitEvent(g.stream, g.kDone[k], 0));
    for (int i = 0; i < Context::kRing; i++) g.ring[i].live = false;
    g.overlapPending = false;
    return 0;
}

// forget the queued generators (their events stay valid): called by every entry point that synchronises the main stream
void clear_producers() { for (int i = 0; i < Context::kRing; i++) g.prod[i].live = false; }

struct Range { const void* p; size_t bytes; };
inline bool overlaps(const char* lo, const char* hi, const Range& r)
{
    return r.p && r.bytes && lo && (const char*)r.p < hi && lo < (const char*)r.p + r.bytes;
}
// order `stream` behind exactly those overlapped launches in flight that write what the caller reads or touch what it writes
int join_for(cudaStream_t stream, const Range* reads, int nr, const Range* writes, int nw)
{
    if (!g.overlapPending) return 0;
    for (int i = 0; i < Context::kRing; i++) {
        Context::InFlight& e = g.ring[i];
        if (!e.live) continue;
        bool hit = false;
        for (int k = 0; k < nw && !hit; k++) hit = overlaps(e.rLo, e.rHi, writes[k]) || overlaps(e.wLo, e.wHi, writes[k]);
        for (int k = 0; k < nr && !hit; k++) hit = overlaps(e.wLo, e.wHi, reads[k]);
        if (!hit) continue;
        NT_CUDA(cudaStreamWaitEvent(stream, e.ev, 0));
        if (stream == g.stream) e.live = false;       // everything queued on the main stream from here on is ordered behind it
    }
    return 0;
}
// may this call skip the library-wide join?  Only overlapped mode with nothing but device memory involved
bool fine_grained(std::initializer_list<const void*> ptrs)
{
    if (!(g.inited && g.deferred && g.overlap)) return false;
    for (const void* p : ptrs) if (p && !is_device_ptr(p)) return false;
    return true;
}

// after a synchronisation point: did any launch since the last check overflow a ray's traversal stack?
int check_trace_error()
{
    if (g.errHost && *g.errHost) {
        *g.errHost = 0;
        set_error("ntrace_b200: BVH too deep for the traversal stack (96 entries; the reference's STACK_SIZE is 64): results of the last launches are incomplete");
        return 1;
    }
    return 0;
}

int require_init(bool join = true)
{
    if (!g.inited) { set_error("ntrace_b200: nt_init() has not been called (no CUDA device selected; there is no CPU fallback)"); return 1; }
    // the caller may be another host thread, or the host framework may have switched devices since nt_init
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != g.device) NT_CUDA(cudaSetDevice(g.device));
    return join ? join_kernel_streams() : 0;
}

// index -> pixel table: 8x8 blocks in Morton order, Morton order inside a block, then the bottom
// stripe column by column and the right stripe row by row (reference behaviour: PixelTable.cpp:57-141).
void build_pixel_table(int w, int h, std::vector<int>& tab)
{
    tab.resize((size_t)w * h);
    size_t n = 0;
    const int bw = w & ~7, bh = h & ~7;
    int side = 1;
    while (side * 8 < bw || side * 8 < bh) side <<= 1;   // power-of-two block grid covering the bulk
    auto compact1by1 = [](unsigned v) {
        v &= 0x55555555u; v = (v | (v >> 1)) & 0x33333333u; v = (v | (v >> 2)) & 0x0f0f0f0fu;
        v = (v | (v >> 4)) & 0x00ff00ffu; v = (v | (v >> 8)) & 0x0000ffffu; return (int)v;
    };
    if (bw > 0 && bh > 0)
        for (unsigned code = 0; code < (unsigned)side * (unsigned)side; code++) {
            const int bx = compact1by1(code), by = compact1by1(code >> 1);
            if (bx * 8 >= bw || by * 8 >= bh) continue;
            for (unsigned in = 0; in < 64; in++) {
                const int ix = compact1by1(in), iy = compact1by1(in >> 1);
                tab[n++] = (by * 8 + iy) * w + (bx * 8 + ix);
            }
        }
    for (int x = 0; x < bw; x++) for (int y = bh; y < h; y++) tab[n++] = x + y * w;
    for (int y = 0; y < h; y++) for (int x = bw; x < w; x++) tab[n++] = x + y * w;
}

int ensure_pixel_table(int w, int h)
{
    if (g.ptW == w && g.ptH == h && g.pixelTable.p) return 0;
    std::vector<int> tab;
    build_pixel_table(w, h, tab);
    NT_CUDA(g.pixelTable.reserve(tab.size() * sizeof(int)));
    NT_CUDA(cudaMemcpyAsync(g.pixelTable.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, g.stream));
    NT_CUDA(cudaStreamSynchronize(g.stream));
    g.ptW = w; g.ptH = h;
    return 0;
}

} // namespace
} // namespace nt

using namespace nt;

extern "C" {

int nt_init(int device_ordinal)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g.inited && g.device == device_ordinal) return 0;
    if (g.inited) { set_error("ntrace_b200: already initialised on another device; call nt_shutdown() first"); return 1; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        set_error("ntrace_b200: no CUDA device available (this library has no CPU fallback)");
        return 1;
    }
    if (device_ordinal < 0 || device_ordinal >= count) { set_error("ntrace_b200: invalid device ordinal"); return 1; }
    NT_CUDA(cudaSetDevice(device_ordinal));
    cudaDeviceProp prop;
    NT_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
    if (prop.major < 10) {
        char buf[256];
        snprintf(buf, sizeof(buf), "ntrace_b200: device %d is sm_%d%d; this library is built for sm_100a only", device_ordinal, prop.major, prop.minor);
        set_error(buf);
        return 1;
    }
    g.device = device_ordinal;
    g.numSMs = prop.multiProcessorCount;
    NT_CUDA(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    NT_CUDA(cudaEventCreate(&g.evA));
    NT_CUDA(cudaEventCreate(&g.evB));
    for (int i = 0; i < 8; i++) NT_CUDA(cudaEventCreate(&g.userEv[i]));
    if (const char* e = getenv("NT_OVERLAP_STREAMS")) { const int v = atoi(e); if (v >= 2 && v <= Context::kMaxKernelStreams) g.numKernelStreams = v; }
    for (int k = 0; k < Context::kMaxKernelStreams; k++) {
        NT_CUDA(cudaStreamCreateWithFlags(&g.kStream[k], cudaStreamNonBlocking));
        NT_CUDA(cudaEventCreateWithFlags(&g.kDone[k], cudaEventDisableTiming));
    }
    NT_CUDA(cudaEventCreateWithFlags(&g.kFork, cudaEventDisableTiming));
    for (int i = 0; i < Context::kRing; i++) NT_CUDA(cudaEventCreateWithFlags(&g.ring[i].ev, cudaEventDisableTiming));
    for (int i = 0; i < Context::kRing; i++) NT_CUDA(cudaEventCreateWithFlags(&g.prod[i].ev, cudaEventDisableTiming));
    NT_CUDA(cudaStreamCreateWithFlags(&g.sIn, cudaStreamNonBlocking));
    NT_CUDA(cudaStreamCreateWithFlags(&g.sOut, cudaStreamNonBlocking));
    for (int i = 0; i < Context::kMaxChunks; i++) {
        NT_CUDA(cudaEventCreateWithFlags(&g.evIn[i], cudaEventDisableTiming));
        NT_CUDA(cudaEventCreate(&g.evKa[i]));
        NT_CUDA(cudaEventCreate(&g.evKb[i]));
    }
    for (int i = 0; i < Context::kAsyncSlots; i++) {
        NT_CUDA(cudaEventCreateWithFlags(&g.async[i].evIn, cudaEventDisableTiming));
        NT_CUDA(cudaEventCreate(&g.async[i].evKa));
        NT_CUDA(cudaEventCreate(&g.async[i].evKb));
        NT_CUDA(cudaEventCreateWithFlags(&g.async[i].evOut, cudaEventDisableTiming));
    }
    NT_CUDA(g.counters.reserve(256));
    NT_CUDA(cudaMemsetAsync(g.counters.p, 0, 256, g.stream));
    NT_CUDA(cudaHostAlloc((void**)&g.errHost, sizeof(int), cudaHostAllocMapped));
    *g.errHost = 0;
    NT_CUDA(cudaHostGetDevicePointer((void**)&g.errDev, g.errHost, 0));
    NT_CUDA(cudaStreamSynchronize(g.stream));
    g.launches = 0;
    g.inited = true;
    return 0;
}

void nt_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g.inited) return;
    cudaSetDevice(g.device);
    cudaStreamSynchronize(g.stream);
    cudaStreamSynchronize(g.sIn);
    cudaStreamSynchronize(g.sOut);
    for (int k = 0; k < Context::kMaxKernelStreams; k++) { cudaStreamSynchronize(g.kStream[k]); cudaStreamDestroy(g.kStream[k]); cudaEventDestroy(g.kDone[k]); }
    cudaEventDestroy(g.kFork);
    for (int i = 0; i < Context::kRing; i++) { cudaEventDestroy(g.ring[i].ev); cudaEventDestroy(g.prod[i].ev); }
    DevBuf* bufs[] = {&g.nodes, &g.woop, &g.triIndex, &g.sortedKeys, &g.sortedIdx, &g.stRays, &g.stResults, &g.stA, &g.stB,
                      &g.stC, &g.stD, &g.stE, &g.counters, &g.pixelTable, &g.sceneVerts, &g.sceneTris,
                      &g.srcNodes, &g.srcWoop, &g.srcIdx, &g.layoutScratch, &g.wideNodes, &g.wideScratch, &g.batchTables};
    for (DevBuf* b : bufs) b->release();
    release_build_scratch();
    release_sort_scratch();
    comm_destroy();
    reset_launch_caches();         // occupancy / carve-out decisions belong to the device they were taken on
    for (int i = 0; i < Context::kAsyncSlots; i++) {
        g.async[i].rays.release(); g.async[i].results.release();
        cudaEventDestroy(g.async[i].evIn); cudaEventDestroy(g.async[i].evKa); cudaEventDestroy(g.async[i].evKb); cudaEventDestroy(g.async[i].evOut);
    }
    cudaEventDestroy(g.evA); cudaEventDestroy(g.evB);
    if (g.errHost) cudaFreeHost(g.errHost);
    for (int i = 0; i < 8; i++) cudaEventDestroy(g.userEv[i]);
    for (int i = 0; i < Context::kMaxChunks; i++) { cudaEventDestroy(g.evIn[i]); cudaEventDestroy(g.evKa[i]); cudaEventDestroy(g.evKb[i]); }
    cudaStreamDestroy(g.sIn); cudaStreamDestroy(g.sOut);
    cudaStreamDestroy(g.stream);
    g = Context();
}

const char* nt_last_error(void) { return t_error.c_str(); }

int64_t nt_launch_count(void) { return g.launches; }

// ---- memory: what FW::Buffer needs from the CUDA runtime, so that a host above this ABI links no CUDA library itself
int nt_mem_alloc(size_t bytes, void** outDevicePtr)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (!outDevicePtr) { set_error("ntrace_b200: null output pointer"); return 1; }
    *outDevicePtr = nullptr;
    if (bytes == 0) return 0;
    NT_CUDA(cudaMalloc(outDevicePtr, bytes));
    return 0;
}
int nt_mem_free(void* devicePtr)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (devicePtr) { NT_CUDA(cudaStreamSynchronize(g.stream)); NT_CUDA(cudaFree(devicePtr)); }
    return 0;
}
int nt_mem_alloc_host(size_t bytes, void** outHostPtr)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (!outHostPtr) { set_error("ntrace_b200: null output pointer"); return 1; }
    *outHostPtr = nullptr;
    if (bytes == 0) return 0;
    NT_CUDA(cudaHostAlloc(outHostPtr, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    return 0;
}
int nt_mem_free_host(void* hostPtr)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (hostPtr) { NT_CUDA(cudaStreamSynchronize(g.stream)); NT_CUDA(cudaFreeHost(hostPtr)); }
    return 0;
}
int nt_memcpy(void* dst, const void* src, size_t bytes)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (bytes == 0) return 0;
    if (!dst || !src) { set_error("ntrace_b200: null pointer in nt_memcpy"); return 1; }
    NT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, g.stream));
    NT_CUDA(cudaStreamSynchronize(g.stream));
    return 0;
}
int nt_memset(void* devicePtr, int value, size_t bytes)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (bytes == 0) return 0;
    if (!devicePtr) { set_error("ntrace_b200: null pointer in nt_memset"); return 1; }
    NT_CUDA(cudaMemsetAsync(devicePtr, value, bytes, g.stream));
    NT_CUDA(cudaStreamSynchronize(g.stream));
    return 0;
}

int nt_event_record(int slot)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (slot < 0 || slot >= 8) { set_error("ntrace_b200: event slot out of range"); return 1; }
    NT_CUDA(cudaEventRecord(g.userEv[slot], g.stream));
    return 0;
}

int nt_event_elapsed(int slotA, int slotB, float* outSeconds)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (slotA < 0 || slotA >= 8 || slotB < 0 || slotB >= 8 || !outSeconds) { set_error("ntrace_b200: bad event arguments"); return 1; }
    NT_CUDA(cudaEventSynchronize(g.userEv[slotB]));
    float ms = 0.0f;
    NT_CUDA(cudaEventElapsedTime(&ms, g.userEv[slotA], g.userEv[slotB]));
    *outSeconds = ms * 1.0e-3f;
    return 0;
}

int nt_set_deferred(int mode)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (mode < 0 || mode > 2) { set_error("ntrace_b200: submission mode must be 0, 1 or 2"); return 1; }
    if (!mode && g.deferred) NT_CUDA(cudaStreamSynchronize(g.stream));         // (require_init joined the kernel streams)
    g.deferred = mode != 0;
    g.overlap = mode == 2;
    clear_producers();
    return 0;
}

int nt_synchronize(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    NT_CUDA(cudaStreamSynchronize(g.stream));
    clear_producers();
    return check_trace_error();
}

int nt_set_kernel(const char* name)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!name) { set_error("ntrace_b200: null kernel name"); return 1; }
    struct Entry { const char* name; int kernel; int layout; bool fast; };
    static const Entry table[] = {
        // IEEE arithmetic in the triangle test: bit-identical to the reference's CPU tracer (CudaBVH::trace / Intersect::RayTriangleWoop)
        {"b200_persistent_speculative_while_while", Kernel_PersistentSpeculative, Layout_Compact, false},
        {"b200_speculative_while_while", Kernel_PlainSpeculative, Layout_Compact, false},
        {"b200_persistent_speculative_while_while_compact2", Kernel_PersistentSpeculative, Layout_Compact2, false},
        // the arithmetic nvcc -use_fast_math gives the reference's GPU kernels (contracted FMAs, approximate 1/x): bit-identical to them
        {"b200_persistent_speculative_while_while_fastmath", Kernel_PersistentSpeculative, Layout_Compact, true},
        {"b200_persistent_speculative_while_while_compact2_fastmath", Kernel_PersistentSpeculative, Layout_Compact2, true},
        // the 4-wide quantised node array derived from the Compact / Compact2 BVH (nt_wide.cu); triangle test as above
        {"b200_wide4", Kernel_Wide4Persistent, Layout_Compact, false},
        {"b200_wide4_fastmath", Kernel_Wide4Persistent, Layout_Compact, true},
        {"b200_wide4_compact2", Kernel_Wide4Persistent, Layout_Compact2, false},
        // two rays per lane, phase-scheduled (nt_wide.cu "mr"): over the binary nodes (bit-identical to the one-ray kernels) and over Wide4
        {"b200_mr", Kernel_BinaryMr, Layout_Compact, false},
        {"b200_mr_fastmath", Kernel_BinaryMr, Layout_Compact, true},
        {"b200_mr_compact2", Kernel_BinaryMr, Layout_Compact2, false},
        // per batch the faster of the two on a B200: any-hit batches through the binary kernel (the reference's visiting order, so the reported
        // any-hit triangle is the reference's too), closest-hit batches through Wide4 (their result does not depend on the visiting order)
        {"b200_auto", Kernel_Auto, Layout_Compact, false},
        {"b200_auto_compact2", Kernel_Auto, Layout_Compact2, false},
        {"b200_sw", Kernel_BinarySw, Layout_Compact, false},
        {"b200_sw_fastmath", Kernel_BinarySw, Layout_Compact, true},
        {"b200_sw_compact2", Kernel_BinarySw, Layout_Compact2, false},
        {"b200_wide4_sw", Kernel_Wide4Sw, Layout_Compact, false},
        {"b200_wide4_sw_fastmath", Kernel_Wide4Sw, Layout_Compact, true},
        {"b200_wide4_sw_compact2", Kernel_Wide4Sw, Layout_Compact2, false},
        {"b200_wide4_mr", Kernel_Wide4Mr, Layout_Compact, false},
        {"b200_wide4_mr_fastmath", Kernel_Wide4Mr, Layout_Compact, true},
        {"b200_wide4_mr_compact2", Kernel_Wide4Mr, Layout_Compact2, false},
        // reference kernel file names (src/rt/kernels/*.cu) accepted as aliases with their layouts AND their arithmetic, so a config
        // that names one of them gets the results that kernel produces
        {"fermi_speculative_while_while", Kernel_PlainSpeculative, Layout_Compact, true},
        {"kepler_dynamic_fetch", Kernel_PersistentSpeculative, Layout_Compact2, true},
        // tesla_* ship with NODES/TRIANGLES_ARRAY_OF_STRUCTURES defined (tesla_persistent_while_while.cu:40-41): AOS_AOS.
        // The BVH is rewritten to the Compact form on the device (nt_layout.cu) and traversed by the one B200 kernel.
        {"tesla_persistent_while_while", Kernel_PersistentSpeculative, Layout_AOS_AOS, true},
        {"tesla_persistent_speculative_while_while", Kernel_PersistentSpeculative, Layout_AOS_AOS, true},
        {"tesla_persistent_packet", Kernel_PersistentSpeculative, Layout_AOS_AOS, true},
        // the three variants those files select by commenting the defines out (IEEE arithmetic, like every b200_* name)
        {"b200_persistent_speculative_while_while_aos_aos", Kernel_PersistentSpeculative, Layout_AOS_AOS, false},
        {"b200_persistent_speculative_while_while_aos_soa", Kernel_PersistentSpeculative, Layout_AOS_SOA, false},
        {"b200_persistent_speculative_while_while_soa_aos", Kernel_PersistentSpeculative, Layout_SOA_AOS, false},
        {"b200_persistent_speculative_while_while_soa_soa", Kernel_PersistentSpeculative, Layout_SOA_SOA, false},
    };
    for (const Entry& e : table)
        if (strcmp(e.name, name) == 0) {
            if (g.kernelLayout != e.layout) g.wideValid = false;
            g.kernel = e.kernel; g.kernelLayout = e.layout; g.fastMath = e.fast;
            return 0;
        }
    set_error(std::string("ntrace_b200: unknown kernel '") + name + "'");
    return 1;
}

int nt_desired_layout(void) { return g.kernelLayout; }

int nt_kernel_config(int32_t out4[4])
{
    KernelConfig c = trace_kernel_config(g.kernel, g.kernelLayout);
    out4[0] = c.bvhLayout; out4[1] = c.blockWidth; out4[2] = c.blockHeight; out4[3] = c.usePersistentThreads;
    return 0;
}

int nt_bvh_alloc(int layout, size_t nodeBytes, size_t woopBytes, size_t idxBytes)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (is_basic_layout(layout)) {
        if (nodeBytes < 64 || nodeBytes % 64 || woopBytes % 64 || idxBytes == 0 || idxBytes % 4 || woopBytes / 64 < idxBytes / 4) {
            set_error("ntrace_b200: inconsistent CudaBVH buffer sizes (AOS/SOA layouts: nodes and Woop triangles in 64 B records, one S32 index per triangle)");
            return 1;
        }
        g.haveBVH = false;             // a failed reserve must not leave the previous BVH's sizes over reallocated buffers
        g.generation++;
        NT_CUDA(g.srcNodes.reserve(nodeBytes));
        NT_CUDA(g.srcWoop.reserve(woopBytes));
        NT_CUDA(g.srcIdx.reserve(idxBytes));
        g.srcNodeBytes = nodeBytes; g.srcWoopBytes = woopBytes; g.srcIdxBytes = idxBytes;
        g.bvhLayout = layout;
        g.basic = true; g.converted = false; g.wideValid = false;
        g.haveBVH = true;
        g.builtTris = 0;
        return 0;
    }
    if (layout != Layout_Compact && layout != Layout_Compact2) { set_error("ntrace_b200: BVHLayout_CPU (no Woop data) cannot be traced on the device"); return 1; }
    if (nodeBytes < 64 || nodeBytes % 64 || woopBytes % 16 || idxBytes * 4 != woopBytes) {
        set_error("ntrace_b200: inconsistent CudaBVH buffer sizes (nodes multiple of 64 B, woop of 16 B, one index per woop float4)");
        return 1;
    }
    // child links are 32-bit and must stay below the EntrypointSentinel 0x76543210 that ends a traversal (a link at or above it
    // would make the kernel spin): byte offsets for Compact, offsets / 16 for Compact2; leaf links ~woopIndex need woopIndex < 2^31
    if (const char* why = link_range_error(layout, nodeBytes, woopBytes)) { set_error(why); return 1; }
    g.haveBVH = false;
    g.generation++;
    g.basic = false; g.converted = false; g.wideValid = false;
    NT_CUDA(g.nodes.reserve(nodeBytes));
    NT_CUDA(g.woop.reserve(woopBytes));
    NT_CUDA(g.triIndex.reserve(idxBytes));
    g.nodeBytes = nodeBytes; g.woopBytes = woopBytes; g.idxBytes = idxBytes;
    g.bvhLayout = layout;
    g.haveBVH = true;
    g.builtTris = 0;
    return 0;
}

int nt_bvh_upload(int layout, const void* nodes, size_t nodeBytes, const void* woop, size_t woopBytes,
                  const int32_t* triIndex, size_t idxBytes)
{
    if (!nodes || !woop || !triIndex) { set_error("ntrace_b200: null BVH buffer"); return 1; }
    if (nt_bvh_alloc(layout, nodeBytes, woopBytes, idxBytes)) return 1;
    std::lock_guard<std::mutex> lock(g_mutex);
    NT_CUDA(cudaMemcpyAsync(g.basic ? g.srcNodes.p : g.nodes.p, nodes, nodeBytes, cudaMemcpyDefault, g.stream));
    NT_CUDA(cudaMemcpyAsync(g.basic ? g.srcWoop.p : g.woop.p, woop, woopBytes, cudaMemcpyDefault, g.stream));
    NT_CUDA(cudaMemcpyAsync(g.basic ? g.srcIdx.p : g.triIndex.p, triIndex, idxBytes, cudaMemcpyDefault, g.stream));
    NT_CUDA(cudaStreamSynchronize(g.stream));
    if (g.basic && ensure_traversal_form()) { g.haveBVH = false; return 1; }     // validates the tree; a malformed upload is refused here
    return 0;
}

int nt_bvh_build(int builder, const float* vtxPos, int numVerts, const int32_t* triVtxIndex, int numTris,
                 const float bboxLo[3], const float bboxHi[3], int hlbvhBits, int leafSize, float epsilon,
                 float* outGpuSeconds)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (!vtxPos || !triVtxIndex || numVerts <= 0 || numTris <= 0) { set_error("ntrace_b200: empty scene"); return 1; }
    if (leafSize < 1) { set_error("ntrace_b200: leafSize must be >= 1"); return 1; }
    if (builder != NT_BUILDER_LBVH && builder != NT_BUILDER_HLBVH) { set_error("ntrace_b200: unknown builder"); return 1; }
    const void *dV, *dT;
    if (stage_in(vtxPos, (size_t)numVerts * 12, g.sceneVerts, &dV)) return 1;
    if (stage_in(triVtxIndex, (size_t)numTris * 12, g.sceneTris, &dT)) return 1;
    BuildParams p;
    p.builder = builder; p.hlbvhBits = hlbvhBits; p.leafSize = leafSize; p.epsilon = epsilon;
    p.collapse = g.collapseMode; p.collapseMaxLeaf = g.collapseMaxLeaf;
    p.layout = g.buildLayout;
    for (int i = 0; i < 3; i++) { p.lo[i] = bboxLo[i]; p.hi[i] = bboxHi[i]; }
    BuildOutput out;
    out.nodes = &g.nodes; out.woop = &g.woop; out.triIndex = &g.triIndex;
    out.sortedKeys = &g.sortedKeys; out.sortedIdx = &g.sortedIdx;
    out.nodeBytes = out.woopBytes = out.idxBytes = 0;
    g.haveBVH = false;
    g.generation++;
    g.basic = false; g.converted = false; g.wideValid = false;
    NT_CUDA(cudaEventRecord(g.evA, g.stream));
    int launches = 0;
    std::string err;
    cudaError_t e = build_bvh_device((const float*)dV, numVerts, (const int*)dT, numTris, p, out, g.stream, g.numSMs, &launches, &err, g.evB);
    g.launches += launches;
    if (e != cudaSuccess) {
        if (!err.empty()) { set_error("ntrace_b200: " + err); cudaGetLastError(); return 1; }
        NT_CUDA(e);
    }
    NT_CUDA(cudaEventSynchronize(g.evB));        // recorded by the builder behind its last kernel, before the size readback
    float ms = 0.0f;
    NT_CUDA(cudaEventElapsedTime(&ms, g.evA, g.evB));
    if (outGpuSeconds) *outGpuSeconds = ms * 1.0e-3f;
    g.nodeBytes = out.nodeBytes; g.woopBytes = out.woopBytes; g.idxBytes = out.idxBytes;
    g.bvhLayout = g.buildLayout;            // the reference's HLBVHBuilder emits Compact only (HLBVHBuilder.cpp:33); Compact2 is new
    g.haveBVH = true;
    g.builtTris = numTris;
    return 0;
}

int nt_bvh_set_collapse(int mode, int maxLeafSize)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (mode != 0 && mode != 1) { set_error("ntrace_b200: collapse mode must be 0 (reference leaf rule) or 1 (SAH)"); return 1; }
    if (maxLeafSize < 0) { set_error("ntrace_b200: negative maxLeafSize"); return 1; }
    g.collapseMode = mode; g.collapseMaxLeaf = maxLeafSize;
    return 0;
}

int nt_bvh_set_build_layout(int layout)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (layout != Layout_Compact && layout != Layout_Compact2) { set_error("ntrace_b200: the GPU builder emits BVHLayout_Compact or BVHLayout_Compact2"); return 1; }
    g.buildLayout = layout;
    return 0;
}

int nt_bvh_convert(int layout)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (!g.haveBVH) { set_error("CudaBVHTracer: No BVH!"); return 1; }
    if (layout != Layout_Compact && layout != Layout_Compact2) { set_error("ntrace_b200: the resident BVH can only be converted to BVHLayout_Compact / Compact2"); return 1; }
    g.wideValid = false;
    if (g.basic) {
        if (ensure_traversal_form()) return 1;                      // AOS/SOA -> Compact (nt_layout.cu)
        g.basic = false; g.converted = false;
        g.bvhLayout = Layout_Compact;
        g.generation++;
    }
    if (g.bvhLayout == layout) return 0;
    if (const char* why = link_range_error(layout, g.nodeBytes, g.woopBytes)) { set_error(why); return 1; }
    g.generation++;
    const bool toCompact2 = (layout == Layout_Compact2);
    NT_CUDA(rescale_compact_links(g.nodes.as<int4>(), g.nodeBytes / 64, toCompact2 ? 1 : 16, toCompact2 ? 16 : 1, g.stream));
    g.launches += 1;
    NT_CUDA(cudaStreamSynchronize(g.stream));
    g.bvhLayout = layout;
    return 0;
}

int nt_bvh_wide4_convert_host(int layout, const void* nodes, size_t nodeBytes, size_t woopBytes,
                              void* outWideNodes, size_t outCapacityBytes, size_t* outWideBytes, int* outMaxDepth)
{
    // pure host code: no device is touched, so this also runs where there is no GPU (it computes a data layout, it does not trace)
    if (!nodes || !outWideBytes) { set_error("ntrace_b200: null pointer in nt_bvh_wide4_convert_host"); return 1; }
    std::vector<uint32_t> w;
    std::string err;
    int depth = 0;
    if (convert_compact_to_wide4_host((const int32_t*)nodes, nodeBytes, layout, woopBytes / 16, w, &depth, &err)) { set_error("ntrace_b200: " + err); return 1; }
    *outWideBytes = w.size() * 4;
    if (outMaxDepth) *outMaxDepth = depth;
    if (outWideNodes) {
        if (outCapacityBytes < w.size() * 4) { set_error("ntrace_b200: output buffer too small for the Wide4 node array"); return 1; }
        memcpy(outWideNodes, w.data(), w.size() * 4);
    }
    return 0;
}

int nt_bvh_wide4_download(void* outWideNodes, size_t outCapacityBytes, size_t* outWideBytes, int* outMaxDepth)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (!outWideBytes) { set_error("ntrace_b200: null pointer in nt_bvh_wide4_download"); return 1; }
    if (!g.haveBVH) { set_error("CudaBVHTracer: No BVH!"); return 1; }
    if (join_kernel_streams()) return 1;
    if (g.basic && ensure_traversal_form()) return 1;
    if (!g.wideValid) {
        const int keep = g.kernel;
        g.kernel = Kernel_Wide4Persistent;
        const int rc = ensure_wide_form();
        g.kernel = keep;
        if (rc) return 1;
    }
    *outWideBytes = g.wideBytes;
    if (outMaxDepth) *outMaxDepth = g.wideDepth;
    if (outWideNodes) {
        if (outCapacityBytes < g.wideBytes) { set_error("ntrace_b200: output buffer too small for the Wide4 node array"); return 1; }
        NT_CUDA(cudaMemcpyAsync(outWideNodes, g.wideNodes.p, g.wideBytes, cudaMemcpyDefault, g.stream));
        NT_CUDA(cudaStreamSynchronize(g.stream));
    }
    return 0;
}

// ---- multi-GPU: communicator + BVH replication (nt_comm.cu) ---------------------------------------------------------------------
int nt_comm_unique_id(void* out128) { std::lock_guard<std::mutex> lock(g_mutex); return comm_unique_id(out128); }

int nt_comm_init(int numRanks, int rank, const void* uniqueId128)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;                     // the communicator lives on the device nt_init selected
    return comm_init(numRanks, rank, uniqueId128);
}

int nt_comm_destroy(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g.inited) { cudaSetDevice(g.device); cudaStreamSynchronize(g.stream); }
    return comm_destroy();
}

int nt_comm_allreduce(double* values, int count, int op)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    return comm_allreduce_f64(values, count, op, g.stream);
}

int nt_bvh_broadcast(int root, float* outSeconds)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (outSeconds) *outSeconds = 0.0f;
    if (require_init()) return 1;
    if (!comm_ready()) { set_error("ntrace_b200: no communicator (nt_comm_init)"); return 1; }
    if (root < 0 || root >= comm_size()) { set_error("ntrace_b200: broadcast root out of range"); return 1; }
    const bool isRoot = comm_rank() == root;
    if (isRoot && !g.haveBVH) { set_error("CudaBVHTracer: No BVH!"); return 1; }
    // header: layout + the three sizes ("+ a 16-byte header {layout, sizes}"; sizes are 64-bit here)
    long long meta[4] = {0, 0, 0, 0};
    if (isRoot) {
        meta[0] = g.bvhLayout;
        meta[1] = (long long)(g.basic ? g.srcNodeBytes : g.nodeBytes);
        meta[2] = (long long)(g.basic ? g.srcWoopBytes : g.woopBytes);
        meta[3] = (long long)(g.basic ? g.srcIdxBytes : g.idxBytes);
    }
    void* dMeta = g.counters.as<char>() + 128;
    NT_CUDA(cudaMemcpyAsync(dMeta, meta, sizeof(meta), cudaMemcpyHostToDevice, g.stream));
    if (comm_broadcast_bytes(dMeta, sizeof(meta), root, g.stream)) return 1;
    NT_CUDA(cudaMemcpyAsync(meta, dMeta, sizeof(meta), cudaMemcpyDeviceToHost, g.stream));
    NT_CUDA(cudaStreamSynchronize(g.stream));
    const int layout = (int)meta[0];
    const size_t nb = (size_t)meta[1], wb = (size_t)meta[2], ib = (size_t)meta[3];
    if (!isRoot) {
        // replica: make room exactly as nt_bvh_alloc does (the old BVH is gone from here on)
        g.haveBVH = false;
        g.generation++;
        g.wideValid = false; g.converted = false;
        g.basic = is_basic_layout(layout);
        if (!g.basic && layout != Layout_Compact && layout != Layout_Compact2) { set_error("ntrace_b200: broadcast header carries an unknown layout"); return 1; }
        NT_CUDA((g.basic ? g.srcNodes : g.nodes).reserve(nb));
        NT_CUDA((g.basic ? g.srcWoop : g.woop).reserve(wb));
        NT_CUDA((g.basic ? g.srcIdx : g.triIndex).reserve(ib));
        if (g.basic) { g.srcNodeBytes = nb; g.srcWoopBytes = wb; g.srcIdxBytes = ib; }
        else { g.nodeBytes = nb; g.woopBytes = wb; g.idxBytes = ib; }
        g.bvhLayout = layout;
        g.builtTris = 0;
    }
    NT_CUDA(cudaEventRecord(g.evA, g.stream));
    if (comm_broadcast_bytes((g.basic ? g.srcNodes : g.nodes).p, nb, root, g.stream)) return 1;
    if (comm_broadcast_bytes((g.basic ? g.srcWoop : g.woop).p, wb, root, g.stream)) return 1;
    if (comm_broadcast_bytes((g.basic ? g.srcIdx : g.triIndex).p, ib, root, g.stream)) return 1;
    NT_CUDA(cudaEventRecord(g.evB, g.stream));
    NT_CUDA(cudaStreamSynchronize(g.stream));
    float ms = 0.0f;
    NT_CUDA(cudaEventElapsedTime(&ms, g.evA, g.evB));
    if (outSeconds) *outSeconds = ms * 1.0e-3f;
    if (!isRoot) {
        g.haveBVH = true;
        if (g.basic && ensure_traversal_form()) { g.haveBVH = false; return 1; }
    }
    return 0;
}

int nt_hash_buffer(const void* ptr, size_t size, uint32_t* outHash)
{
    // FW::hashBuffer (src/framework/base/Hash.cpp:33-75; Bob Jenkins' 1996 mix, Hash.hpp:169-181): pure host code
    if (!outHash || (!ptr && size) || size > 0x7fffffffull) { set_error("ntrace_b200: bad arguments to nt_hash_buffer"); return 1; }
#define NT_JENKINS_MIX(a, b, c) \
    a -= b; a -= c; a ^= (c >> 13); b -= c; b -= a; b ^= (a << 8);  c -= a; c -= b; c ^= (b >> 13); \
    a -= b; a -= c; a ^= (c >> 12); b -= c; b -= a; b ^= (a << 16); c -= a; c -= b; c ^= (b >> 5);  \
    a -= b; a -= c; a ^= (c >> 3);  b -= c; b -= a; b ^= (a << 10); c -= a; c -= b; c ^= (b >> 15);
    const uint8_t* src = (const uint8_t*)ptr;
    uint32_t a = 0x9e3779b9u, b = 0x9e3779b9u, c = 0x9e3779b9u;
    size_t left = size;
    while (left >= 12) {
        a += (uint32_t)src[0] + ((uint32_t)src[1] << 8) + ((uint32_t)src[2] << 16) + ((uint32_t)src[3] << 24);
        b += (uint32_t)src[4] + ((uint32_t)src[5] << 8) + ((uint32_t)src[6] << 16) + ((uint32_t)src[7] << 24);
        c += (uint32_t)src[8] + ((uint32_t)src[9] << 8) + ((uint32_t)src[10] << 16) + ((uint32_t)src[11] << 24);
        NT_JENKINS_MIX(a, b, c);
        src += 12; left -= 12;
    }
    switch (left) {
    case 11: c += (uint32_t)src[10] << 16;   // fall through
    case 10: c += (uint32_t)src[9] << 8;     // fall through
    case 9:  c += (uint32_t)src[8];          // fall through
    case 8:  b += (uint32_t)src[7] << 24;    // fall through
    case 7:  b += (uint32_t)src[6] << 16;    // fall through
    case 6:  b += (uint32_t)src[5] << 8;     // fall through
    case 5:  b += (uint32_t)src[4];          // fall through
    case 4:  a += (uint32_t)src[3] << 24;    // fall through
    case 3:  a += (uint32_t)src[2] << 16;    // fall through
    case 2:  a += (uint32_t)src[1] << 8;     // fall through
    case 1:  a += (uint32_t)src[0];          // fall through
    default: break;
    }
    c += (uint32_t)left;
    NT_JENKINS_MIX(a, b, c);
#undef NT_JENKINS_MIX
    *outHash = c;
    return 0;
}

int nt_bvh_sah(double* outSah, int64_t* outNumInner, int64_t* outNumLeaves, int64_t* outNumTris)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (!outSah) { set_error("ntrace_b200: null output"); return 1; }
    if (!g.haveBVH) { set_error("CudaBVHTracer: No BVH!"); return 1; }
    if (ensure_traversal_form()) return 1;                 // AOS / SOA uploads are measured through their Compact form (same tree)
    struct { double terms, rootArea; unsigned long long leaves, tris; } h;
    void* d = g.counters.as<char>() + 192;
    NT_CUDA(cudaMemsetAsync(d, 0, 32, g.stream));
    NT_CUDA(launch_sah(g.nodes.as<float4>(), g.nodeBytes / 64, g.woop.as<float4>(), g.woopBytes / 16, d, g.stream));
    g.launches += 1;
    NT_CUDA(cudaMemcpyAsync(&h, d, 32, cudaMemcpyDeviceToHost, g.stream));
    NT_CUDA(cudaStreamSynchronize(g.stream));
    *outSah = 2.0 + (h.rootArea > 0.0 ? h.terms / h.rootArea : 0.0);       // the root's own 2 * Cn at P = 1
    if (outNumInner) *outNumInner = (int64_t)(g.nodeBytes / 64);
    if (outNumLeaves) *outNumLeaves = (int64_t)h.leaves;
    if (outNumTris) *outNumTris = (int64_t)h.tris;
    return 0;
}

int nt_bvh_generation(uint64_t* outGeneration)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init(false)) return 1;                 // touches no device memory: launches in flight are not joined
    if (!outGeneration) { set_error("ntrace_b200: null output"); return 1; }
    *outGeneration = g.haveBVH ? g.generation : 0;
    return 0;
}

int nt_bvh_sizes(size_t sizes[3], int* outLayout)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (!g.haveBVH) { set_error("CudaBVHTracer: No BVH!"); return 1; }
    if (g.basic) { sizes[0] = g.srcNodeBytes; sizes[1] = g.srcWoopBytes; sizes[2] = g.srcIdxBytes; }
    else { sizes[0] = g.nodeBytes; sizes[1] = g.woopBytes; sizes[2] = g.idxBytes; }
    if (outLayout) *outLayout = g.bvhLayout;
    return 0;
}

int nt_bvh_download(void* nodes, void* woop, int32_t* triIndex)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (!g.haveBVH) { set_error("CudaBVHTracer: No BVH!"); return 1; }
    if (g.basic) {
        if (nodes) NT_CUDA(cudaMemcpyAsync(nodes, g.srcNodes.p, g.srcNodeBytes, cudaMemcpyDefault, g.stream));
        if (woop) NT_CUDA(cudaMemcpyAsync(woop, g.srcWoop.p, g.srcWoopBytes, cudaMemcpyDefault, g.stream));
        if (triIndex) NT_CUDA(cudaMemcpyAsync(triIndex, g.srcIdx.p, g.srcIdxBytes, cudaMemcpyDefault, g.stream));
    } else {
        if (nodes) NT_CUDA(cudaMemcpyAsync(nodes, g.nodes.p, g.nodeBytes, cudaMemcpyDefault, g.stream));
        if (woop) NT_CUDA(cudaMemcpyAsync(woop, g.woop.p, g.woopBytes, cudaMemcpyDefault, g.stream));
        if (triIndex) NT_CUDA(cudaMemcpyAsync(triIndex, g.triIndex.p, g.idxBytes, cudaMemcpyDefault, g.stream));
    }
    NT_CUDA(cudaStreamSynchronize(g.stream));
    return 0;
}

int nt_bvh_device_ptrs(void* ptrs[3])
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (!g.haveBVH) { set_error("CudaBVHTracer: No BVH!"); return 1; }
    g.wideValid = false;                                                                                          // caller may write (broadcast)
    if (g.basic) { ptrs[0] = g.srcNodes.p; ptrs[1] = g.srcWoop.p; ptrs[2] = g.srcIdx.p; g.converted = false; }
    else { ptrs[0] = g.nodes.p; ptrs[1] = g.woop.p; ptrs[2] = g.triIndex.p; }
    return 0;
}

int nt_bvh_build_debug(uint32_t* sortedKeys, int32_t* sortedIdx, int numTris)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (g.builtTris <= 0 || numTris != g.builtTris) { set_error("ntrace_b200: no GPU build of that size is resident"); return 1; }
    if (sortedKeys) NT_CUDA(cudaMemcpyAsync(sortedKeys, g.sortedKeys.p, (size_t)numTris * 4, cudaMemcpyDefault, g.stream));
    if (sortedIdx) NT_CUDA(cudaMemcpyAsync(sortedIdx, g.sortedIdx.p, (size_t)numTris * 4, cudaMemcpyDefault, g.stream));
    NT_CUDA(cudaStreamSynchronize(g.stream));
    return 0;
}

int nt_trace_batch(const float* rays, int32_t* results, int numRays, int needClosestHit, float* outSeconds)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (outSeconds) *outSeconds = 0.0f;
    // overlapped deferred launches take device buffers and page-locked host buffers (traversed in place over PCIe)
    const bool overlapped = g.inited && g.deferred && g.overlap && rays && results && mapped_device_ptr(rays) && mapped_device_ptr(results);
    if (require_init(!overlapped)) return 1;
    if (numRays == 0) return 0;                                        // CudaBVHTracer.cpp:92-94
    if (numRays < 0 || !rays || !results) { set_error("ntrace_b200: invalid ray batch"); return 1; }
    if (numRays > 0x3fffffff) { set_error("ntrace_b200: a batch holds at most 2^30 - 1 rays (32-bit ray indexing in the kernels)"); return 1; }
    if (!g.haveBVH) { set_error("CudaBVHTracer: No BVH!"); return 1; }                          // :98-99
    if (g.bvhLayout != g.kernelLayout) { set_error("CudaBVHTracer: Incorrect BVH layout!"); return 1; }  // :100-101
    if (ensure_traversal_form() || ensure_wide_form()) return 1;

    // Pinned (page-locked, UVA-mapped) host buffers are traversed in place: the kernel reads rays and writes results
    // over PCIe (zero copy), which overlaps both transfers with the traversal inside ONE launch.  Pageable host memory
    // cannot be touched by the device and goes through the staged, chunked copy pipeline below.
    const void* raysDev = mapped_device_ptr(rays);
    void* resDev = mapped_device_ptr(results);
    const bool raysOnHost = (raysDev == nullptr), resOnHost = (resDev == nullptr);

    TraceLaunch a;
    a.kernel = auto_kernel(raysDev, numRays, needClosestHit);
    a.fast = g.fastMath ? 1 : 0; a.layout = g.basic ? (int)Layout_Compact : g.kernelLayout; a.anyHit = needClosestHit ? 0 : 1;
    a.nodes = g.nodes.as<float4>(); a.woop = g.woop.as<float4>(); a.triIndices = g.triIndex.as<int>();
    a.wideNodes = g.wideNodes.as<float4>();
    a.numSMs = g.numSMs; a.stream = g.stream; a.errorFlag = g.errDev;
    int launches = 0;

    if (raysOnHost || resOnHost) {
        // Host buffers (the reference's Buffer would migrate them): stage through device buffers in chunks so that the
        // H2D copy of chunk i+1, the kernel on chunk i and the D2H copy of chunk i-1 overlap on three streams.
        const float4* dRays = (const float4*)raysDev; int4* dRes = (int4*)resDev;
        if (raysOnHost) { NT_CUDA(g.stRays.reserve((size_t)numRays * 32)); dRays = g.stRays.as<float4>(); }
        if (resOnHost) { NT_CUDA(g.stResults.reserve((size_t)numRays * 16)); dRes = g.stResults.as<int4>(); }
        static const int wantChunks = [] { const char* e = getenv("NT_E2E_CHUNKS"); int v = e ? atoi(e) : 2; return v < 1 ? 1 : (v > Context::kMaxChunks ? Context::kMaxChunks : v); }();
        int chunk = (numRays + wantChunks - 1) / wantChunks;
        if (chunk < 32768) chunk = 32768;
        chunk = (chunk + 127) & ~127;
        const int numChunks = (numRays + chunk - 1) / chunk;          // <= wantChunks <= kMaxChunks
        for (int i = 0; i < numChunks; i++) {
            const int lo = i * chunk, cnt = (numRays - lo < chunk) ? numRays - lo : chunk;
            if (raysOnHost) {
                NT_CUDA(cudaMemcpyAsync((void*)(dRays + (size_t)lo * 2), rays + (size_t)lo * 8, (size_t)cnt * 32, cudaMemcpyHostToDevice, g.sIn));
                NT_CUDA(cudaEventRecord(g.evIn[i], g.sIn));
                NT_CUDA(cudaStreamWaitEvent(g.stream, g.evIn[i], 0));
            }
            a.numRays = cnt; a.rays = dRays + (size_t)lo * 2; a.results = dRes + lo;
            a.warpCounter = g.counters.as<int>() + 8 + i;
            NT_CUDA(cudaMemsetAsync(a.warpCounter, 0, sizeof(int), g.stream));
            NT_CUDA(cudaEventRecord(g.evKa[i], g.stream));
            int l = 0;
            NT_CUDA(launch_trace(a, &l));
            launches += l;
            NT_CUDA(cudaEventRecord(g.evKb[i], g.stream));
            if (resOnHost) {
                NT_CUDA(cudaStreamWaitEvent(g.sOut, g.evKb[i], 0));
                NT_CUDA(cudaMemcpyAsync(results + (size_t)lo * 4, dRes + lo, (size_t)cnt * 16, cudaMemcpyDeviceToHost, g.sOut));
            }
        }
        g.launches += launches;
        NT_CUDA(cudaStreamSynchronize(g.stream));
        NT_CUDA(cudaStreamSynchronize(g.sOut));
        float total = 0.0f;
        for (int i = 0; i < numChunks; i++) { float ms = 0.0f; NT_CUDA(cudaEventElapsedTime(&ms, g.evKa[i], g.evKb[i])); total += ms; }
        if (outSeconds) *outSeconds = total * 1.0e-3f;             // kernel time only, as the reference reports it
        return check_trace_error();
    }

    a.numRays = numRays; a.rays = (const float4*)raysDev; a.results = (int4*)resDev;
    if (overlapped) {
        const int k = g.overlapNext; g.overlapNext = (g.overlapNext + 1) % g.numKernelStreams;
        a.stream = g.kStream[k];
        // the rays' generator comes first.  When it was a queued nt_raygen_ao the launch waits for that kernel alone; otherwise for
        // everything the main stream has queued so far
        bool haveProducer = false;
        for (int i = 0; i < Context::kRing; i++) {
            Context::Producer& pr = g.prod[i];
            if (pr.live && (const char*)raysDev < pr.hi && pr.lo < (const char*)raysDev + (size_t)numRays * 32) {
                NT_CUDA(cudaStreamWaitEvent(a.stream, pr.ev, 0));
                haveProducer = true;
            }
        }
        if (!haveProducer) {
            NT_CUDA(cudaEventRecord(g.kFork, g.stream));
            NT_CUDA(cudaStreamWaitEvent(a.stream, g.kFork, 0));
        }
        // launches in flight on the OTHER kernel stream that share a buffer with this one (the same batch traced twice, say) come first too
        const Range rd{raysDev, (size_t)numRays * 32}, wr{resDev, (size_t)numRays * 16};
        if (join_for(a.stream, &rd, 1, &wr, 1)) return 1;
        a.warpCounter = g.counters.as<int>() + 4 + k;                  // one fetch counter per launch in flight
        NT_CUDA(cudaMemsetAsync(a.warpCounter, 0, sizeof(int), a.stream));
        NT_CUDA(launch_trace(a, &launches));
        NT_CUDA(cudaEventRecord(g.kDone[k], a.stream));
        Context::InFlight& slot = g.ring[g.ringNext];
        g.ringNext = (g.ringNext + 1) % Context::kRing;
        if (slot.live) NT_CUDA(cudaStreamWaitEvent(g.stream, slot.ev, 0));     // the ring remembers 8 launches: the one that falls out is joined
        slot.rLo = (const char*)raysDev; slot.rHi = slot.rLo + (size_t)numRays * 32;
        slot.wLo = (const char*)resDev; slot.wHi = slot.wLo + (size_t)numRays * 16;
        NT_CUDA(cudaEventRecord(slot.ev, a.stream));
        slot.live = true;
        g.overlapPending = true;
        g.launches += launches;
        return 0;
    }
    a.warpCounter = g.counters.as<int>();
    // the counter reset is issued before the first event so the timed interval is the kernel only
    NT_CUDA(cudaMemsetAsync(a.warpCounter, 0, sizeof(int), g.stream));
    if (g.deferred && is_device_ptr(rays) && is_device_ptr(results)) {
        NT_CUDA(launch_trace(a, &launches));
        g.launches += launches;
        return 0;
    }
    NT_CUDA(cudaEventRecord(g.evA, g.stream));
    NT_CUDA(launch_trace(a, &launches));
    NT_CUDA(cudaEventRecord(g.evB, g.stream));
    g.launches += launches;
    NT_CUDA(cudaStreamSynchronize(g.stream));
    float ms = 0.0f;
    NT_CUDA(cudaEventElapsedTime(&ms, g.evA, g.evB));
    if (outSeconds) *outSeconds = ms * 1.0e-3f;
    return check_trace_error();
}

// Several device-resident batches in ONE persistent launch.  A persistent launch on its own spends its last ~20 % waiting for the
// rays still in flight (profiles/r2_summary.md, section 12: 24 launches of 1 Mi diffuse rays 2 851 Mrays/s, the same rays in one launch
// 3 673); nt_set_deferred(2) hides most of that by overlapping consecutive launches on two streams, one launch over the frame's batches
// removes it.  The kernels are the same: ray index i of the launch is looked up in a small device table at ray fetch and at result store.
int nt_trace_batches(int numBatches, const float* const* rays, int32_t* const* results, const int32_t* numRays, int needClosestHit, float* outSeconds)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (outSeconds) *outSeconds = 0.0f;
    const bool overlapped = g.inited && g.deferred && g.overlap;
    if (require_init(!overlapped)) return 1;
    if (numBatches == 0) return 0;
    if (numBatches < 0 || !rays || !results || !numRays) { set_error("ntrace_b200: invalid batch list"); return 1; }
    if (!g.haveBVH) { set_error("CudaBVHTracer: No BVH!"); return 1; }
    if (g.bvhLayout != g.kernelLayout) { set_error("CudaBVHTracer: Incorrect BVH layout!"); return 1; }
    const int kernel = auto_kernel(nullptr, 0, needClosestHit);
    if (kernel != Kernel_PersistentSpeculative && kernel != Kernel_Wide4Persistent) {
        set_error("ntrace_b200: nt_trace_batches needs a persistent one-ray kernel (b200_persistent_speculative_while_while*, b200_wide4*, b200_auto*)");
        return 1;
    }
    if (ensure_traversal_form() || ensure_wide_form()) return 1;
    NT_CUDA(g.batchTables.reserve(sizeof(BatchTable) * Context::kBatchTableSlots * (Context::kMaxKernelStreams + 1)));
    float totalMs = 0.0f;
    int launches = 0;
    for (int first = 0; first < numBatches; first += kMaxBatches) {
        const int cnt = (numBatches - first < kMaxBatches) ? numBatches - first : kMaxBatches;
        BatchTable t;
        memset(&t, 0, sizeof(t));
        long long total = 0;
        int used = 0;
        Range rd[kMaxBatches], wr[kMaxBatches];
        const char *rLo = nullptr, *rHi = nullptr, *wLo = nullptr, *wHi = nullptr;
        for (int i = 0; i < cnt; i++) {
            const int n = numRays[first + i];
            if (n < 0) { set_error("ntrace_b200: negative batch size"); return 1; }
            if (n == 0) continue;
            const float* r = rays[first + i]; int32_t* o = results[first + i];
            if (!r || !o || !is_device_ptr(r) || !is_device_ptr(o)) { set_error("ntrace_b200: nt_trace_batches takes device buffers (use nt_trace_batch / nt_trace_batch_async for host memory)"); return 1; }
            if ((reinterpret_cast<size_t>(r) & 31) || (reinterpret_cast<size_t>(o) & 15)) { set_error("ntrace_b200: ray buffers must be 32-byte aligned, result buffers 16-byte aligned"); return 1; }
            t.start[used] = (int)total; t.rays[used] = (const float4*)r; t.results[used] = (int4*)o;
            rd[used] = Range{r, (size_t)n * 32}; wr[used] = Range{o, (size_t)n * 16};
            const char *a0 = (const char*)r, *a1 = a0 + (size_t)n * 32, *b0 = (const char*)o, *b1 = b0 + (size_t)n * 16;
            if (!rLo || a0 < rLo) rLo = a0; if (!rHi || a1 > rHi) rHi = a1;
            if (!wLo || b0 < wLo) wLo = b0; if (!wHi || b1 > wHi) wHi = b1;
            total += n; used++;
            if (total > 0x3fffffffLL) { set_error("ntrace_b200: a launch holds at most 2^30 - 1 rays (32-bit ray indexing in the kernels)"); return 1; }
        }
        if (used == 0) continue;
        t.count = used; t.start[used] = (int)total;
        TraceLaunch a;
        a.kernel = kernel;
        a.fast = g.fastMath ? 1 : 0; a.layout = g.basic ? (int)Layout_Compact : g.kernelLayout; a.anyHit = needClosestHit ? 0 : 1;
        a.nodes = g.nodes.as<float4>(); a.woop = g.woop.as<float4>(); a.triIndices = g.triIndex.as<int>();
        a.wideNodes = g.wideNodes.as<float4>();
        a.errorFlag = g.errDev;
        a.numSMs = g.numSMs;
        a.numRays = (int)total; a.rays = nullptr; a.results = nullptr;
        // overlapped mode: the launch goes to one of the kernel streams like a queued nt_trace_batch -- behind what the main stream has queued
        // and behind the launches in flight that touch its buffers -- so that its CTAs move in while the launch before it drains
        const int k = overlapped ? g.overlapNext : Context::kMaxKernelStreams;
        if (overlapped) {
            g.overlapNext = (g.overlapNext + 1) % g.numKernelStreams;
            a.stream = g.kStream[k];
            NT_CUDA(cudaEventRecord(g.kFork, g.stream));
            NT_CUDA(cudaStreamWaitEvent(a.stream, g.kFork, 0));
            a.warpCounter = g.counters.as<int>() + 4 + k;
        } else {
            a.stream = g.stream;
            a.warpCounter = g.counters.as<int>();
        }
        if (join_for(a.stream, rd, used, wr, used)) return 1;
        BatchTable* dT = g.batchTables.as<BatchTable>() + k * Context::kBatchTableSlots + g.batchTableNext[k];
        g.batchTableNext[k] = (g.batchTableNext[k] + 1) % Context::kBatchTableSlots;
        NT_CUDA(cudaMemcpyAsync(dT, &t, sizeof(BatchTable), cudaMemcpyHostToDevice, a.stream));     // pageable source: staged before the call returns
        a.batches = dT;
        NT_CUDA(cudaMemsetAsync(a.warpCounter, 0, sizeof(int), a.stream));
        int l = 0;
        if (overlapped) {
            NT_CUDA(launch_trace(a, &l));
            NT_CUDA(cudaEventRecord(g.kDone[k], a.stream));
            Context::InFlight& slot = g.ring[g.ringNext];
            g.ringNext = (g.ringNext + 1) % Context::kRing;
            if (slot.live) NT_CUDA(cudaStreamWaitEvent(g.stream, slot.ev, 0));
            slot.rLo = rLo; slot.rHi = rHi; slot.wLo = wLo; slot.wHi = wHi;        // the hull of the launch's buffers: conservative
            NT_CUDA(cudaEventRecord(slot.ev, a.stream));
            slot.live = true;
            g.overlapPending = true;
        } else if (g.deferred) {
            NT_CUDA(launch_trace(a, &l));
        } else {
            NT_CUDA(cudaEventRecord(g.evA, g.stream));
            NT_CUDA(launch_trace(a, &l));
            NT_CUDA(cudaEventRecord(g.evB, g.stream));
            NT_CUDA(cudaStreamSynchronize(g.stream));
            float ms = 0.0f;
            NT_CUDA(cudaEventElapsedTime(&ms, g.evA, g.evB));
            totalMs += ms;
        }
        launches += l;
    }
    g.launches += launches;
    if (g.deferred) return 0;
    if (outSeconds) *outSeconds = totalMs * 1.0e-3f;
    return check_trace_error();
}

// Asynchronous form of nt_trace_batch for a host loop that keeps several independent batches in flight (the batches of a
// frame are independent once the primary results exist).  Rays in pinned host memory are DMA'd into the slot's device
// staging on the copy-in stream, the kernel runs on the compute stream, results go back by DMA on the copy-out stream:
// with >= 2 slots the H2D copy of batch i+1, the traversal of batch i and the D2H copy of batch i-1 overlap, so the
// steady state is bound by the larger PCIe direction (32 B/ray in) instead of the sum of the three stages.
int nt_trace_batch_async(const float* rays, int32_t* results, int numRays, int needClosestHit, int slot)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (slot < 0 || slot >= Context::kAsyncSlots) { set_error("ntrace_b200: async slot out of range"); return 1; }
    Context::AsyncSlot& s = g.async[slot];
    if (s.busy) { set_error("ntrace_b200: async slot still in flight; call nt_trace_wait first"); return 1; }
    if (numRays <= 0 || numRays > 0x3fffffff || !rays || !results) { set_error("ntrace_b200: invalid ray batch"); return 1; }
    if (!g.haveBVH) { set_error("CudaBVHTracer: No BVH!"); return 1; }
    if (g.bvhLayout != g.kernelLayout) { set_error("CudaBVHTracer: Incorrect BVH layout!"); return 1; }
    if (ensure_traversal_form() || ensure_wide_form()) return 1;
    const bool raysDev = is_device_ptr(rays), resDev = is_device_ptr(results);
    if ((!raysDev && !mapped_device_ptr(rays)) || (!resDev && !mapped_device_ptr(results))) {
        set_error("ntrace_b200: asynchronous submission needs device or pinned (page-locked) host buffers");
        return 1;
    }
    static const bool zeroCopyResults = [] { const char* e = getenv("NT_ASYNC_RESULTS"); return e && strcmp(e, "zc") == 0; }();
    const float4* dRays = (const float4*)rays;
    if (!raysDev) {
        NT_CUDA(s.rays.reserve((size_t)numRays * 32));
        NT_CUDA(cudaMemcpyAsync(s.rays.p, rays, (size_t)numRays * 32, cudaMemcpyHostToDevice, g.sIn));
        NT_CUDA(cudaEventRecord(s.evIn, g.sIn));
        NT_CUDA(cudaStreamWaitEvent(g.stream, s.evIn, 0));
        dRays = s.rays.as<float4>();
    }
    int4* dRes = (int4*)results;
    s.copyOut = false;
    if (!resDev) {
        if (zeroCopyResults) dRes = (int4*)mapped_device_ptr(results);
        else { NT_CUDA(s.results.reserve((size_t)numRays * 16)); dRes = s.results.as<int4>(); s.copyOut = true; }
    }
    TraceLaunch a;
    a.kernel = auto_kernel(nullptr, numRays, needClosestHit);          // host-side batches carry no generator hint
    a.fast = g.fastMath ? 1 : 0; a.layout = g.basic ? (int)Layout_Compact : g.kernelLayout; a.anyHit = needClosestHit ? 0 : 1;
    a.nodes = g.nodes.as<float4>(); a.woop = g.woop.as<float4>(); a.triIndices = g.triIndex.as<int>();
    a.wideNodes = g.wideNodes.as<float4>();
    a.numSMs = g.numSMs; a.stream = g.stream; a.errorFlag = g.errDev;
    a.numRays = numRays; a.rays = dRays; a.results = dRes;
    a.warpCounter = g.counters.as<int>() + 32 + slot;
    NT_CUDA(cudaMemsetAsync(a.warpCounter, 0, sizeof(int), g.stream));
    NT_CUDA(cudaEventRecord(s.evKa, g.stream));
    int launches = 0;
    NT_CUDA(launch_trace(a, &launches));
    g.launches += launches;
    NT_CUDA(cudaEventRecord(s.evKb, g.stream));
    if (s.copyOut) {
        NT_CUDA(cudaStreamWaitEvent(g.sOut, s.evKb, 0));
        NT_CUDA(cudaMemcpyAsync(results, dRes, (size_t)numRays * 16, cudaMemcpyDeviceToHost, g.sOut));
        NT_CUDA(cudaEventRecord(s.evOut, g.sOut));
    }
    s.busy = true;
    return 0;
}

int nt_trace_wait(int slot, float* outSeconds)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (outSeconds) *outSeconds = 0.0f;
    if (require_init()) return 1;
    if (slot < 0 || slot >= Context::kAsyncSlots) { set_error("ntrace_b200: async slot out of range"); return 1; }
    Context::AsyncSlot& s = g.async[slot];
    if (!s.busy) return 0;
    s.busy = false;
    if (s.copyOut) {
        NT_CUDA(cudaEventSynchronize(s.evOut));
        NT_CUDA(cudaStreamWaitEvent(g.stream, s.evOut, 0));     // later work / events on the compute stream order after the copy-out
    } else NT_CUDA(cudaEventSynchronize(s.evKb));
    float ms = 0.0f;
    NT_CUDA(cudaEventElapsedTime(&ms, s.evKa, s.evKb));
    if (outSeconds) *outSeconds = ms * 1.0e-3f;                 // kernel time only, like nt_trace_batch
    return check_trace_error();
}

int nt_raygen_primary(float* rays, int32_t* idToSlot, int32_t* slotToID, const float origin[3],
                      const float nscreenToWorld[16], int w, int h, float maxDist, uint32_t randomSeed)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (w <= 0 || h <= 0 || !rays) { set_error("ntrace_b200: invalid primary ray request"); return 1; }
    if (ensure_pixel_table(w, h)) return 1;
    const size_t n = (size_t)w * h;
    void *dRays, *hRays, *dI2S, *hI2S, *dS2I, *hS2I;
    if (stage_out(rays, n * 32, g.stRays, &dRays, &hRays)) return 1;
    if (stage_out(idToSlot, n * 4, g.stA, &dI2S, &hI2S)) return 1;
    if (stage_out(slotToID, n * 4, g.stB, &dS2I, &hS2I)) return 1;
    PrimaryArgs a;
    a.rays = (float4*)dRays; a.idToSlot = (int*)dI2S; a.slotToID = (int*)dS2I;
    for (int i = 0; i < 3; i++) a.origin[i] = origin[i];
    for (int i = 0; i < 16; i++) a.n2w[i] = nscreenToWorld[i];
    a.w = w; a.h = h; a.maxDist = maxDist; a.seed = randomSeed;
    NT_CUDA(launch_raygen_primary(a, g.pixelTable.as<int>(), g.stream));
    g.launches += 1;
    if (!hRays) { g.primLo = (const char*)dRays; g.primHi = g.primLo + n * 32; } else { g.primLo = g.primHi = nullptr; }
    if (copy_back(hRays, dRays, n * 32) || copy_back(hI2S, dI2S, n * 4) || copy_back(hS2I, dS2I, n * 4)) return 1;
    NT_CUDA(cudaStreamSynchronize(g.stream));
    return 0;
}

int nt_raygen_ao(float* outRays, int32_t* outIDToSlot, int32_t* outSlotToID, const float* inRays,
                 const int32_t* inResults, const float* triNormals, int firstInputSlot, int numInputRays,
                 int numSamples, float maxDist, uint32_t randomSeed)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    // overlapped mode, device buffers only: wait for just the launches in flight that use these buffers, and do not block the host
    const bool fine = fine_grained({outRays, outIDToSlot, outSlotToID, inRays, inResults, triNormals});
    if (require_init(!fine)) return 1;
    if (numInputRays == 0) return 0;
    if (numInputRays < 0 || numSamples <= 0 || firstInputSlot < 0 || !outRays || !inRays || !inResults || !triNormals) {
        set_error("ntrace_b200: invalid AO ray request");
        return 1;
    }
    if (fine) {
        const size_t nO = (size_t)numInputRays * numSamples;
        const Range rd[2] = {{inRays + (size_t)firstInputSlot * 8, (size_t)numInputRays * 32}, {inResults + (size_t)firstInputSlot * 4, (size_t)numInputRays * 16}};
        const Range wr[3] = {{outRays, nO * 32}, {outIDToSlot, nO * 4}, {outSlotToID, nO * 4}};
        if (join_for(g.stream, rd, 2, wr, 3)) return 1;
    }
    if (is_device_ptr(inRays) != is_device_ptr(inResults)) { set_error("ntrace_b200: inRays and inResults must live on the same side"); return 1; }
    const size_t nOut = (size_t)numInputRays * numSamples;
    if (nOut > 0x3fffffffull) { set_error("ntrace_b200: AO batch too large"); return 1; }
    // host inputs: only the [first, first+count) window is staged, so the device slot base becomes 0
    const void *dInRays, *dInRes;
    int first = firstInputSlot;
    if (!is_device_ptr(inRays)) {
        if (stage_in(inRays + (size_t)firstInputSlot * 8, (size_t)numInputRays * 32, g.stC, &dInRays)) return 1;
        if (stage_in(inResults + (size_t)firstInputSlot * 4, (size_t)numInputRays * 16, g.stD, &dInRes)) return 1;
        first = 0;
    } else { dInRays = inRays; dInRes = inResults; }
    const void* dNormals = triNormals;
    if (!is_device_ptr(triNormals)) { set_error("ntrace_b200: triNormals must be a device pointer (use nt_tri_normals with a device output)"); return 1; }
    void *dOut, *hOut, *dA, *hA, *dB, *hB;
    if (stage_out(outRays, nOut * 32, g.stRays, &dOut, &hOut)) return 1;
    if (stage_out(outIDToSlot, nOut * 4, g.stA, &dA, &hA)) return 1;
    if (stage_out(outSlotToID, nOut * 4, g.stB, &dB, &hB)) return 1;
    AOArgs a;
    a.outRays = (float4*)dOut; a.outIDToSlot = (int*)dA; a.outSlotToID = (int*)dB;
    a.inRays = (const float4*)dInRays; a.inResults = (const int4*)dInRes; a.normals = (const float*)dNormals;
    a.firstInputSlot = first; a.numInputRays = numInputRays; a.numSamples = numSamples; a.maxDist = maxDist; a.seed = randomSeed;
    a.order = g.raygenOrder;
    if (g.primLo && (const char*)dOut < g.primHi && g.primLo < (const char*)dOut + nOut * 32) g.primLo = g.primHi = nullptr;
    NT_CUDA(launch_raygen_ao(a, g.stream));
    g.launches += 1;
    if (fine) {                               // queued on the main stream; the trace of these rays waits for this kernel (see nt_trace_batch)
        Context::Producer& pr = g.prod[g.prodNext];
        g.prodNext = (g.prodNext + 1) % Context::kRing;
        pr.lo = (const char*)dOut; pr.hi = pr.lo + nOut * 32;
        NT_CUDA(cudaEventRecord(pr.ev, g.stream));
        pr.live = true;
        return 0;
    }
    if (copy_back(hOut, dOut, nOut * 32) || copy_back(hA, dA, nOut * 4) || copy_back(hB, dB, nOut * 4)) return 1;
    NT_CUDA(cudaStreamSynchronize(g.stream));
    return 0;
}

int nt_raygen_set_order(int mode)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (mode != 0 && mode != 1) { set_error("ntrace_b200: ray generation order must be 0 (reference slot order) or 1 (direction-coherent tiles)"); return 1; }
    g.raygenOrder = mode;
    return 0;
}

int nt_raygen_shadow(float* outRays, int32_t* outIDToSlot, int32_t* outSlotToID, const float* inRays, const int32_t* inResults,
                     int firstInputSlot, int numInputRays, int numSamples, const float lightPos[3], float lightRadius, uint32_t randomSeed)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (numInputRays == 0) return 0;
    if (numInputRays < 0 || numSamples <= 0 || firstInputSlot < 0 || !outRays || !inRays || !inResults || !lightPos) {
        set_error("ntrace_b200: invalid shadow ray request");
        return 1;
    }
    if (is_device_ptr(inRays) != is_device_ptr(inResults)) { set_error("ntrace_b200: inRays and inResults must live on the same side"); return 1; }
    const size_t nOut = (size_t)numInputRays * numSamples;
    if (nOut > 0x3fffffffull) { set_error("ntrace_b200: shadow batch too large"); return 1; }
    const void *dInRays, *dInRes;
    int first = firstInputSlot;
    if (!is_device_ptr(inRays)) {
        if (stage_in(inRays + (size_t)firstInputSlot * 8, (size_t)numInputRays * 32, g.stC, &dInRays)) return 1;
        if (stage_in(inResults + (size_t)firstInputSlot * 4, (size_t)numInputRays * 16, g.stD, &dInRes)) return 1;
        first = 0;
    } else { dInRays = inRays; dInRes = inResults; }
    void *dOut, *hOut, *dA, *hA, *dB, *hB;
    if (stage_out(outRays, nOut * 32, g.stRays, &dOut, &hOut)) return 1;
    if (stage_out(outIDToSlot, nOut * 4, g.stA, &dA, &hA)) return 1;
    if (stage_out(outSlotToID, nOut * 4, g.stB, &dB, &hB)) return 1;
    ShadowArgs a;
    a.outRays = (float4*)dOut; a.outIDToSlot = (int*)dA; a.outSlotToID = (int*)dB;
    a.inRays = (const float4*)dInRays; a.inResults = (const int4*)dInRes;
    a.firstInputSlot = first; a.numInputRays = numInputRays; a.numSamples = numSamples;
    for (int i = 0; i < 3; i++) a.lightPos[i] = lightPos[i];
    a.lightRadius = lightRadius; a.seed = randomSeed;
    NT_CUDA(launch_raygen_shadow(a, g.stream));
    g.launches += 1;
    if (copy_back(hOut, dOut, nOut * 32) || copy_back(hA, dA, nOut * 4) || copy_back(hB, dB, nOut * 4)) return 1;
    NT_CUDA(cudaStreamSynchronize(g.stream));
    return 0;
}

int nt_ray_sort(float* rays, int32_t* idToSlot, int32_t* slotToID, int numRays)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (numRays <= 1) return 0;
    if (!rays || !idToSlot || !slotToID) { set_error("ntrace_b200: null ray buffer"); return 1; }
    const bool dev = is_device_ptr(rays);
    if (dev != is_device_ptr(idToSlot) || dev != is_device_ptr(slotToID)) { set_error("ntrace_b200: ray buffer and id maps must live on the same side"); return 1; }
    float4* dRays = (float4*)rays; int* dI2S = idToSlot; int* dS2I = slotToID;
    if (!dev) {
        NT_CUDA(g.stRays.reserve((size_t)numRays * 32)); NT_CUDA(g.stA.reserve((size_t)numRays * 4)); NT_CUDA(g.stB.reserve((size_t)numRays * 4));
        dRays = g.stRays.as<float4>(); dI2S = g.stA.as<int>(); dS2I = g.stB.as<int>();
        NT_CUDA(cudaMemcpyAsync(dRays, rays, (size_t)numRays * 32, cudaMemcpyHostToDevice, g.stream));
        NT_CUDA(cudaMemcpyAsync(dS2I, slotToID, (size_t)numRays * 4, cudaMemcpyHostToDevice, g.stream));
    }
    int launches = 0;
    cudaError_t e = ray_sort_device(dRays, dI2S, dS2I, numRays, g.stream, g.numSMs, &launches);
    g.launches += launches;
    NT_CUDA(e);
    if (!dev) {
        NT_CUDA(cudaMemcpyAsync(rays, dRays, (size_t)numRays * 32, cudaMemcpyDeviceToHost, g.stream));
        NT_CUDA(cudaMemcpyAsync(idToSlot, dI2S, (size_t)numRays * 4, cudaMemcpyDeviceToHost, g.stream));
        NT_CUDA(cudaMemcpyAsync(slotToID, dS2I, (size_t)numRays * 4, cudaMemcpyDeviceToHost, g.stream));
    }
    NT_CUDA(cudaStreamSynchronize(g.stream));
    return 0;
}

int nt_count_hits(const int32_t* results, int numRays, int* outHits)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (!outHits) { set_error("ntrace_b200: null output"); return 1; }
    *outHits = 0;
    if (numRays == 0) return 0;
    if (numRays < 0 || !results) { set_error("ntrace_b200: invalid result buffer"); return 1; }
    const void* dRes;
    if (stage_in(results, (size_t)numRays * 16, g.stResults, &dRes)) return 1;
    int* counter = g.counters.as<int>() + 1;
    NT_CUDA(launch_count_hits((const int4*)dRes, numRays, counter, g.stream));
    g.launches += 1;
    NT_CUDA(cudaMemcpyAsync(outHits, counter, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    NT_CUDA(cudaStreamSynchronize(g.stream));
    return 0;
}

int nt_tri_normals(const float* vtxPos, int numVerts, const int32_t* triVtxIndex, int numTris, float* outNormals)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (require_init()) return 1;
    if (!vtxPos || !triVtxIndex || !outNormals || numVerts <= 0 || numTris <= 0) { set_error("ntrace_b200: empty scene"); return 1; }
    const void *dV, *dT; void *dN, *hN;
    if (stage_in(vtxPos, (size_t)numVerts * 12, g.sceneVerts, &dV)) return 1;
    if (stage_in(triVtxIndex, (size_t)numTris * 12, g.sceneTris, &dT)) return 1;
    if (stage_out(outNormals, (size_t)numTris * 12, g.stE, &dN, &hN)) return 1;
    NT_CUDA(launch_tri_normals((const float*)dV, (const int*)dT, numTris, (float*)dN, g.stream));
    g.launches += 1;
    if (copy_back(hN, dN, (size_t)numTris * 12)) return 1;
    NT_CUDA(cudaStreamSynchronize(g.stream));
    return 0;
}

} // extern "C"
