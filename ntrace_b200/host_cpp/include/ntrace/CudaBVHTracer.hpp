// FW::CudaVirtualTracer / FW::CudaBVHTracer — the tracer object Renderer drives.
// Reference: src/rt/cuda/CudaVirtualTracer.hpp:11-26 (interface), src/rt/cuda/CudaBVHTracer.cpp:52-84 (setKernel +
// queryConfig), :88-168 (traceBatch: empty batch -> 0, "No BVH!", "Incorrect BVH layout!", GPU seconds around the kernel).
#pragma once
#include <vector>
#include "ntrace/CudaBVH.hpp"
#include "ntrace/RayBuffer.hpp"

namespace FW
{
struct KernelConfig { int bvhLayout, blockWidth, blockHeight, usePersistentThreads; };    // CudaTracerKernels.hpp:69-75

class CudaVirtualTracer
{
public:
    virtual ~CudaVirtualTracer() {}
    virtual void setKernel(const String& name) = 0;
    virtual BVHLayout getDesiredBVHLayout() const = 0;
    virtual void setBVH(CudaAS* bvh) = 0;
    virtual void setScene(Scene* scene) = 0;
    virtual F32 traceBatch(RayBuffer& rays) = 0;                     // GPU seconds
};

class CudaBVHTracer : public CudaVirtualTracer
{
public:
    CudaBVHTracer() : m_bvh(NULL), m_scene(NULL) { setKernel("b200_persistent_speculative_while_while"); }

    virtual void setKernel(const String& name)
    {
        if (name == m_kernelName) return;
        ntCheck(nt_set_kernel(name.c_str()));
        m_kernelName = name;
        int32_t c[4];
        ntCheck(nt_kernel_config(c));
        m_kernelConfig.bvhLayout = c[0]; m_kernelConfig.blockWidth = c[1]; m_kernelConfig.blockHeight = c[2]; m_kernelConfig.usePersistentThreads = c[3];
    }
    virtual BVHLayout getDesiredBVHLayout() const { return (BVHLayout)m_kernelConfig.bvhLayout; }
    const KernelConfig& getKernelConfig() const { return m_kernelConfig; }

    virtual void setBVH(CudaAS* as)
    {
        m_bvh = as;
        if (!as) return;
        CudaBVH* bvh = dynamic_cast<CudaBVH*>(as);
        if (bvh && bvh->getGeneration() && bvh->getGeneration() == CudaBVH::currentGeneration()) return;   // already the resident BVH
        if (bvh && bvh->isResident() && !bvh->hasHostCopy())
            fail("CudaBVHTracer: this BVH was built on the device and has been replaced by a later build / upload; rebuild it or keep a host copy");
        Buffer& n = as->getNodeBuffer(); Buffer& w = as->getTriWoopBuffer(); Buffer& i = as->getTriIndexBuffer();
        ntCheck(nt_bvh_upload((int)as->getLayout(), n.getPtr(), (size_t)n.getSize(), w.getPtr(), (size_t)w.getSize(),
                              (const int32_t*)i.getPtr(), (size_t)i.getSize()));
        if (bvh) bvh->setGeneration(CudaBVH::currentGeneration());
    }
    virtual void setScene(Scene* scene) { m_scene = scene; }

    virtual F32 traceBatch(RayBuffer& rays)
    {
        if (!rays.getSize()) return 0.0f;                          // CudaBVHTracer.cpp:92-94
        if (!m_bvh) fail("CudaBVHTracer: No BVH!");
        if (m_bvh->getLayout() != getDesiredBVHLayout()) fail("CudaBVHTracer: Incorrect BVH layout!");
        if (CudaBVH* bvh = dynamic_cast<CudaBVH*>(m_bvh))
            if (bvh->getGeneration() && bvh->getGeneration() != CudaBVH::currentGeneration()) setBVH(m_bvh);   // another handle became resident in between
        float sec = 0.0f;
        ntCheck(nt_trace_batch((const float*)rays.getRayBuffer().getCudaPtr(), (int32_t*)rays.getResultBuffer().getMutableCudaPtrDiscard(),
                               rays.getSize(), rays.getNeedClosestHit() ? 1 : 0, &sec));
        return sec;
    }

    // NEW (nt_trace_batches): several RayBuffers with the same closest / any-hit flag in ONE persistent launch; same results per ray as
    // traceBatch on each, without the ramp-up and drain of every launch but one.  Device-resident buffers.
    F32 traceBatches(const std::vector<RayBuffer*>& batches)
    {
        std::vector<const float*> rays; std::vector<int32_t*> results; std::vector<int32_t> counts;
        int closest = -1;
        for (size_t i = 0; i < batches.size(); i++) {
            RayBuffer& b = *batches[i];
            if (!b.getSize()) continue;
            if (closest >= 0 && closest != (b.getNeedClosestHit() ? 1 : 0)) fail("CudaBVHTracer: the batches of one launch share the closest / any-hit flag");
            closest = b.getNeedClosestHit() ? 1 : 0;
            rays.push_back((const float*)b.getRayBuffer().getCudaPtr());
            results.push_back((int32_t*)b.getResultBuffer().getMutableCudaPtrDiscard());
            counts.push_back(b.getSize());
        }
        if (rays.empty()) return 0.0f;
        if (!m_bvh) fail("CudaBVHTracer: No BVH!");
        if (m_bvh->getLayout() != getDesiredBVHLayout()) fail("CudaBVHTracer: Incorrect BVH layout!");
        if (CudaBVH* bvh = dynamic_cast<CudaBVH*>(m_bvh))
            if (bvh->getGeneration() && bvh->getGeneration() != CudaBVH::currentGeneration()) setBVH(m_bvh);
        float sec = 0.0f;
        ntCheck(nt_trace_batches((int)rays.size(), rays.data(), results.data(), counts.data(), closest, &sec));
        return sec;
    }

private:
    CudaAS* m_bvh;
    Scene* m_scene;
    String m_kernelName;
    KernelConfig m_kernelConfig;
};
}
