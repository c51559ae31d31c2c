// FW::Renderer — the frame / batch loop of the reference renderer, restricted to the tracing path (no display, no shading).
// Reference: src/rt/cuda/Renderer.hpp:44-190, Renderer.cpp:134-160 (setParams), :168-232 (getCudaBVH incl. the bvhcache file),
// :405-579 (beginFrame / nextBatch / traceBatch), :676-710 (getTotalNumRays).
#pragma once
#include <vector>
#include <memory>
#include "ntrace/CameraControls.hpp"
#include "ntrace/CudaBVHTracer.hpp"
#include "ntrace/Environment.hpp"
#include "ntrace/RayGen.hpp"
#include <fstream>
#include <memory>

namespace FW
{
class Renderer
{
public:
    enum RayType { RayType_Primary = 0, RayType_AO, RayType_Diffuse, RayType_Max };

    struct Params
    {
        String kernelName; RayType rayType; F32 aoRadius; S32 numSamples; bool sortSecondary;
        Params() : kernelName("b200_persistent_speculative_while_while"), rayType(RayType_Primary), aoRadius(5.0f), numSamples(32), sortSecondary(false) {}
    };

    // the seed used when Raygen.random is false: the reference hashes Random(0).getU32() (a RANROT value); every parity
    // check of this repo uses this constant on both sides instead
    enum { FixedSecondarySeed = 0x9E3779B9 };

    Renderer() : m_raygen(1 << 20), m_scene(NULL), m_useCachePath(false), m_rank(0), m_numGpus(1), m_batchCounter(0), m_genIdx(0), m_queueHead(0), m_queueLen(0), m_genDone(false), m_pipelined(false), m_cameraFar(0.0f), m_newBatch(true), m_batchRays(NULL), m_batchStart(0)   // Renderer.cpp:45
    {
        m_cudaTracer.reset(new CudaBVHTracer());
        m_builder = "HLBVH";
        Environment::GetSingleton()->GetStringValue("Renderer.builder", m_builder);
        m_cachePath = "bvhcache";
    }

    void setScene(Scene* scene) { m_scene = scene; m_accelStruct.reset(); m_cudaTracer->setScene(scene); }
    Scene* getScene() const { return m_scene; }
    void setParams(const Params& p) { m_params = p; m_cudaTracer->setKernel(p.kernelName); }
    const Params& getParams() const { return m_params; }
    void setHLBVHParams(const HLBVHParams& p) { m_hlbvh = p; m_accelStruct.reset(); }
    void setCudaBVH(CudaBVH* bvh) { m_accelStruct.reset(bvh); }           // takes ownership: a prebuilt / deserialised BVH

    // NEW (no reference counterpart).  false: the reference's loop, every traceBatch() returns its kernel seconds.  true: nextBatch()
    // cycles three secondary RayBuffers (the reference has one, Renderer.hpp m_secondaryRays, and traces batch i before it generates
    // batch i+1) and generates up to Prefetch batches ahead of the one it hands out, traceBatch() only queues (returns 0) and
    // beginTiming() / endTiming() bracket the batch loop: later batches are generated while batch i is traced and consecutive launches
    // overlap at their tails (nt_set_deferred(2); the library orders every call behind the launches / the generator that use ITS
    // buffers only).  Results are bit-identical to the synchronous loop.
    enum { NumSecondary = 3, Prefetch = 2 };
    void setPipelined(bool on) { m_pipelined = on; }

    // NEW: Renderer.numGpus (the reference is single-GPU).  One process per GPU, all running the same loop after nt_comm_init:
    // rank 0 builds (or imports) the BVH and nt_bvh_broadcast replicates it; the secondary batches of a frame are dealt round-robin —
    // batch b is generated and traced by rank b mod numGpus only, the others just advance the batching — and the one primary batch of a
    // primary-ray frame goes to rank 0.  getTotalNumRays() stays the FRAME's ray count on every rank.
    void setMultiGpu(int rank, int numGpus)
    {
        if (numGpus < 1 || rank < 0 || rank >= numGpus) fail("Renderer: bad rank / numGpus");
        m_rank = rank; m_numGpus = numGpus;
        m_accelStruct.reset();
    }
    int getRank() const { return m_rank; }
    int getNumGpus() const { return m_numGpus; }
    F32 getBroadcastTime() const { return m_broadcastTime; }
    void beginTiming()
    {
        if (m_pipelined) {
            for (int i = 0; i < NumSecondary; i++) m_secondary[i].reserve(m_raygen.getMaxBatchSize());      // no allocation inside the loop
            ntCheck(nt_set_deferred(2));
        }
        ntCheck(nt_event_record(6));
    }
    F32 endTiming()                                                         // device seconds since beginTiming(); waits for everything queued
    {
        float sec = 0.0f;
        ntCheck(nt_event_record(7));
        ntCheck(nt_event_elapsed(6, 7, &sec));
        if (m_pipelined) { ntCheck(nt_synchronize()); ntCheck(nt_set_deferred(0)); }
        return sec;
    }
    CudaBVHTracer& getCudaTracer() { return *m_cudaTracer; }
    RayBuffer& getPrimaryRays() { return m_primaryRays; }

    // Renderer.cpp:168-232.  GPU builders only (the CPU SAH / Split builders are the oracle's business); a BVH whose cache
    // file exists is loaded from it, a freshly built one is written when Renderer.cacheDataStructure is set.
    CudaAS* getCudaBVH()
    {
        BVHLayout layout = m_cudaTracer->getDesiredBVHLayout();
        if (m_accelStruct && m_accelStruct->getLayout() == layout) return m_accelStruct.get();
        if (m_numGpus > 1) {
            if (m_rank == 0) buildOrImport(layout);
            float sec = 0.0f;
            ntCheck(nt_bvh_broadcast(0, &sec));                      // collective: every rank is here
            m_broadcastTime = sec;
            if (m_rank != 0) m_accelStruct.reset(CudaBVH::adoptResident());
            else if (!m_accelStruct->isResident()) m_cudaTracer->setBVH(m_accelStruct.get());   // (an imported cache file was uploaded before the broadcast)
            return m_accelStruct.get();
        }
        return buildOrImport(layout);
    }

private:
    CudaAS* buildOrImport(BVHLayout layout)
    {
        if (!m_scene) fail("Renderer: no scene");
        if (m_builder != "HLBVH" && m_builder != "LBVH") fail("Unsupported BVH builder %s (this host builds HLBVH | LBVH on the GPU)", m_builder.c_str());
        if (layout != BVHLayout_Compact) fail("HLBVHBuilder output is BVHLayout_Compact only (HLBVHBuilder.cpp:33)");
        bool cache = false;
        Environment::GetSingleton()->GetBoolValue("Renderer.cacheDataStructure", cache);
        String cacheFile = m_cacheFileOverride;
        if (cache && cacheFile.empty() && m_useCachePath) cacheFile = getCacheFileName(layout);
        if (cache && !cacheFile.empty()) {
            std::ifstream in(cacheFile.c_str(), std::ios::binary);
            if (in) {
                m_accelStruct.reset(new CudaBVH(in));
                if (m_accelStruct->getLayout() == layout) {
                    if (m_numGpus > 1) m_cudaTracer->setBVH(m_accelStruct.get());      // resident before the broadcast
                    return m_accelStruct.get();
                }
            }
        }
        HLBVHParams p = m_hlbvh;
        if (m_builder == "LBVH") { p.hlbvh = false; p.hlbvhBits = 10; }
        m_accelStruct.reset(new HLBVHBuilder(m_scene, p));             // Renderer.cpp:201-209
        if (cache && !cacheFile.empty()) { std::ofstream out(cacheFile.c_str(), std::ios::binary); if (out) m_accelStruct->serialize(out); }
        return m_accelStruct.get();
    }

public:
    // Cache file: setCacheFile() names it outright; setCachePath() turns on the reference's naming "<cachePath>/<hash>_<builder>.dat"
    // (Renderer.cpp:173-178; the reference's path is "bvhcache").
    void setCacheFile(const String& path) { m_cacheFileOverride = path; }
    void setCachePath(const String& dir) { m_cachePath = dir; m_useCachePath = true; }

    static U32 jenkins3(U32 a, U32 b, U32 c, int which)                     // FW_JENKINS_MIX (Hash.hpp:172-181)
    {
        a -= b; a -= c; a ^= (c >> 13); b -= c; b -= a; b ^= (a << 8);  c -= a; c -= b; c ^= (b >> 13);
        a -= b; a -= c; a ^= (c >> 12); b -= c; b -= a; b ^= (a << 16); c -= a; c -= b; c ^= (b >> 5);
        a -= b; a -= c; a ^= (c >> 3);  b -= c; b -= a; b ^= (a << 10); c -= a; c -= b; c ^= (b >> 15);
        return which == 0 ? a : which == 1 ? b : c;
    }
    static U32 hashBits(U32 a, U32 b = 0x9e3779b9u, U32 c = 0) { c += 0x9e3779b9u; return jenkins3(a, b, c, 2); }        // Hash.hpp:183
    static U32 hashBits(U32 a, U32 b, U32 c, U32 d, U32 e = 0, U32 f = 0)                                                   // Hash.hpp:184
    {
        c += 0x9e3779b9u;
        const U32 a1 = jenkins3(a, b, c, 0), b1 = jenkins3(a, b, c, 1), c1 = jenkins3(a, b, c, 2);
        return jenkins3(a1 + d, b1 + e, c1 + f, 2);
    }
    static U32 hashBuffer(const void* p, size_t n) { uint32_t h = 0; ntCheck(nt_hash_buffer(p, n, &h)); return h; }
    static U32 floatBits(F32 f) { U32 u; memcpy(&u, &f, 4); return u; }

    // "<cachePath>/<hash>_<builder>.dat" as Renderer::getCudaBVH forms it: hashBits(Scene::hash (Scene.cpp:171-179: this path carries no
    // material colours, those two buffers enter as empty), Platform("GPU") with leaf preferences (1, 1) (Platform.hpp:162,
    // Renderer.cpp:88-89), BuildParams::computeHash (splitAlpha 1e-5, BVH.hpp:141-144), layout, hash of Renderer.dataStructure)
    String getCacheFileName(BVHLayout layout)
    {
        if (!m_scene) fail("Renderer: no scene");
        Buffer& ti = m_scene->getTriVtxIndexBuffer(); Buffer& tn = m_scene->getTriNormalBuffer(); Buffer& vp = m_scene->getVtxPosBuffer();
        const U32 empty = hashBuffer(NULL, 0);
        const U32 sceneHash = hashBits(hashBuffer(ti.getPtr(), (size_t)ti.getSize()), hashBuffer(tn.getPtr(), (size_t)tn.getSize()), empty, empty,
                                       hashBuffer(vp.getPtr(), (size_t)vp.getSize()));
        const U32 platform = hashBits(hashBuffer("GPU", 3), floatBits(1.0f), floatBits(1.0f), hashBits(1, 1, 1, 1));
        const U32 params = hashBits(floatBits(1.0e-5f));
        String ds = "BVH";
        Environment::GetSingleton()->GetStringValue("Renderer.dataStructure", ds);
        char buf[32];
        snprintf(buf, sizeof(buf), "%08x", hashBits(sceneHash, platform, params, (U32)layout, hashBuffer(ds.c_str(), ds.size())));
        return m_cachePath + "/" + buf + "_" + m_builder + ".dat";
    }

    void beginFrame(const CameraControls& camera, int w, int h)           // Renderer.cpp:405-500 without GL
    {
        m_cudaTracer->setBVH(getCudaBVH());
        m_raygen.primary(m_primaryRays, camera.getPosition(), camera.getNScreenToWorld(w, h), w, h, camera.getFar(), 0);
        if (m_params.rayType != RayType_Primary) m_cudaTracer->traceBatch(m_primaryRays);     // :481-484
        m_cameraFar = camera.getFar();
        m_newBatch = true;
        m_batchRays = NULL;
        m_batchStart = 0;
        m_queueHead = m_queueLen = 0;
        m_genDone = false;
        m_batchCounter = 0;
    }

    bool nextBatch()                                                        // Renderer.cpp:504-566
    {
        if (m_batchRays) m_batchStart += m_batchRays->getSize();
        m_batchRays = NULL;
        if (m_pipelined && m_params.rayType != RayType_Primary) {
            const bool closest = (m_params.rayType == RayType_Diffuse);
            while (m_queueLen < Prefetch && !m_genDone) {
                if (m_numGpus > 1 && (m_batchCounter++ % m_numGpus) != m_rank) {       // another rank's batch
                    if (!m_raygen.skipAo(m_primaryRays, m_params.numSamples, m_newBatch)) m_genDone = true;
                    continue;
                }
                RayBuffer& sec = m_secondary[m_genIdx % NumSecondary];
                if (!m_raygen.ao(sec, m_primaryRays, *m_scene, m_params.numSamples, closest ? m_cameraFar : m_params.aoRadius, m_newBatch, FixedSecondarySeed)) { m_genDone = true; break; }
                sec.setNeedClosestHit(closest);
                if (m_params.sortSecondary) sec.mortonSort();
                m_queue[(m_queueHead + m_queueLen++) % NumSecondary] = &sec;
                m_genIdx++;
            }
            if (!m_queueLen) return false;
            m_batchRays = m_queue[m_queueHead];
            m_queueHead = (m_queueHead + 1) % NumSecondary;
            m_queueLen--;
            return true;
        }
        if (m_numGpus > 1 && m_params.rayType != RayType_Primary)                        // skip the batches dealt to other ranks
            while ((m_batchCounter++ % m_numGpus) != m_rank)
                if (!m_raygen.skipAo(m_primaryRays, m_params.numSamples, m_newBatch)) return false;
        switch (m_params.rayType) {
        case RayType_Primary:
            if (!m_newBatch) return false;
            m_newBatch = false;
            if (m_rank != 0) return false;                                              // the frame's one primary batch is rank 0's
            m_batchRays = &m_primaryRays;
            break;
        case RayType_AO: {
            RayBuffer& sec = m_secondary[0];
            if (!m_raygen.ao(sec, m_primaryRays, *m_scene, m_params.numSamples, m_params.aoRadius, m_newBatch, FixedSecondarySeed)) return false;
            m_batchRays = &sec;
            break;
        }
        case RayType_Diffuse: {
            RayBuffer& sec = m_secondary[0];
            if (!m_raygen.ao(sec, m_primaryRays, *m_scene, m_params.numSamples, m_cameraFar, m_newBatch, FixedSecondarySeed)) return false;
            sec.setNeedClosestHit(true);
            m_batchRays = &sec;
            break;
        }
        default:
            fail("Renderer: unsupported ray type");
        }
        if (m_params.sortSecondary) m_batchRays->mortonSort();             // :561-562 (the reference's condition is a tautology)
        return true;
    }

    F32 traceBatch()                                                        // Renderer.cpp:568-579
    {
        if (!m_batchRays) fail("Renderer: no batch");
        return m_cudaTracer->traceBatch(*m_batchRays);
    }
    RayBuffer* getBatchRays() { return m_batchRays; }

    // NEW (no reference counterpart; the reference's Renderer holds ONE secondary RayBuffer).  After beginFrame(): generate EVERY batch of
    // the frame for the current ray type (this rank's share with numGpus > 1), each into its own RayBuffer of a grow-only pool; returns the
    // number of batches.  Same batches, same rays as the nextBatch() loop produces one after the other.
    int prepareFrame()
    {
        m_frameBatches.clear();
        if (m_params.rayType == RayType_Primary) { if (m_rank == 0) m_frameBatches.push_back(&m_primaryRays); return (int)m_frameBatches.size(); }
        if (m_params.rayType != RayType_AO && m_params.rayType != RayType_Diffuse) fail("Renderer: unsupported ray type");
        const bool closest = (m_params.rayType == RayType_Diffuse);
        bool newBatch = true;
        int counter = 0;
        for (;;) {
            if (m_numGpus > 1 && (counter++ % m_numGpus) != m_rank) {                 // another rank's batch
                if (!m_raygen.skipAo(m_primaryRays, m_params.numSamples, newBatch)) break;
                continue;
            }
            if (m_frameBatches.size() == m_framePool.size()) m_framePool.push_back(std::unique_ptr<RayBuffer>(new RayBuffer()));
            RayBuffer& rb = *m_framePool[m_frameBatches.size()];
            if (!m_raygen.ao(rb, m_primaryRays, *m_scene, m_params.numSamples, closest ? m_cameraFar : m_params.aoRadius, newBatch, FixedSecondarySeed)) break;
            rb.setNeedClosestHit(closest);
            if (m_params.sortSecondary) rb.mortonSort();
            m_frameBatches.push_back(&rb);
        }
        return (int)m_frameBatches.size();
    }
    // NEW: the batches prepareFrame() generated, traced by ONE persistent launch (CudaBVHTracer::traceBatches = nt_trace_batches): the same
    // results per ray as traceBatch() on each, without the ramp-up and drain of every launch but one.  Returns the kernel seconds.
    F32 traceFrame()
    {
        if (m_frameBatches.empty()) return 0.0f;
        if (m_frameBatches.size() == 1) return m_cudaTracer->traceBatch(*m_frameBatches[0]);
        return m_cudaTracer->traceBatches(m_frameBatches);
    }
    const std::vector<RayBuffer*>& getFrameBatches() const { return m_frameBatches; }

    S32 getTotalNumRays()                                                   // Renderer.cpp:676-710
    {
        if (m_params.rayType == RayType_Primary) return m_primaryRays.getSize();
        int hits = 0;
        ntCheck(nt_count_hits((const int32_t*)m_primaryRays.getResultBuffer().getCudaPtr(), m_primaryRays.getSize(), &hits));
        return hits * m_params.numSamples;
    }

private:
    RayGen m_raygen;
    std::unique_ptr<CudaBVHTracer> m_cudaTracer;
    std::unique_ptr<CudaBVH> m_accelStruct;
    Scene* m_scene;
    Params m_params;
    HLBVHParams m_hlbvh;
    String m_builder, m_cachePath, m_cacheFileOverride;
    RayBuffer m_primaryRays, m_secondary[NumSecondary];
    std::vector<std::unique_ptr<RayBuffer> > m_framePool;
    std::vector<RayBuffer*> m_frameBatches;
    bool m_useCachePath;
    int m_rank, m_numGpus, m_batchCounter;
    F32 m_broadcastTime = 0.0f;
    RayBuffer* m_queue[NumSecondary];
    int m_genIdx, m_queueHead, m_queueLen;
    bool m_genDone;
    bool m_pipelined;
    F32 m_cameraFar;
    bool m_newBatch;
    RayBuffer* m_batchRays;
    S32 m_batchStart;
};
}
