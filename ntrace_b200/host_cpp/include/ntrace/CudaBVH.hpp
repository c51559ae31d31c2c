// FW::CudaAS / FW::CudaBVH / FW::HLBVHBuilder — the acceleration structure objects of the tracing path.
// Reference: src/rt/cuda/CudaAS.hpp:20-64 (interface), src/rt/cuda/CudaBVH.hpp:137-152 + CudaBVH.cpp:105-125 (the three
// buffers, layout tag, bvhcache stream format "S32 layout; {S64 size; bytes} x 3", Buffer.cpp:349-381),
// src/rt/bvh/HLBVH/HLBVHBuilder.hpp:25-41 (HLBVHParams, GPU build straight into the Compact layout).
#pragma once
#include "ntrace/Scene.hpp"
#include <istream>
#include <ostream>

namespace FW
{
class CudaAS
{
public:
    virtual ~CudaAS() {}
    virtual BVHLayout getLayout() const = 0;
    virtual Buffer& getNodeBuffer() = 0;
    virtual Buffer& getTriWoopBuffer() = 0;
    virtual Buffer& getTriIndexBuffer() = 0;
    virtual void serialize(std::ostream& out) = 0;
};

class CudaBVH : public CudaAS
{
public:
    explicit CudaBVH(BVHLayout layout = BVHLayout_Compact) : m_layout(layout), m_resident(false), m_generation(0) {}
    explicit CudaBVH(std::istream& in) : m_layout(BVHLayout_Max), m_resident(false), m_generation(0)              // CudaBVH.cpp:105-112
    {
        S32 layout = 0;
        in.read((char*)&layout, 4);
        if (!in || layout < 0 || layout >= BVHLayout_Max) fail("Corrupt CudaBVH stream!");
        m_layout = (BVHLayout)layout;
        Buffer* b[3] = {&m_nodes, &m_triWoop, &m_triIndex};
        for (int i = 0; i < 3; i++) {
            S64 size = 0;
            in.read((char*)&size, 8);
            if (!in || size < 0) fail("Corrupt CudaBVH stream!");
            b[i]->resizeDiscard(size);
            if (size) in.read((char*)b[i]->getMutablePtr(), size);
            if (!in) fail("Corrupt CudaBVH stream!");
        }
    }

    virtual BVHLayout getLayout() const { return m_layout; }
    virtual Buffer& getNodeBuffer() { materialise(); return m_nodes; }
    virtual Buffer& getTriWoopBuffer() { materialise(); return m_triWoop; }
    virtual Buffer& getTriIndexBuffer() { materialise(); return m_triIndex; }

    virtual void serialize(std::ostream& out)                                                         // CudaBVH.cpp:116-125
    {
        materialise();
        S32 layout = (S32)m_layout;
        out.write((const char*)&layout, 4);
        Buffer* b[3] = {&m_nodes, &m_triWoop, &m_triIndex};
        for (int i = 0; i < 3; i++) {
            S64 size = b[i]->getSize();
            out.write((const char*)&size, 8);
            if (size) out.write((const char*)b[i]->getPtr(), size);
        }
    }

    // true while the buffers live only inside the library (after a GPU build)
    bool isResident() const { return m_resident; }
    // The library holds ONE resident BVH; nt_bvh_generation() changes whenever it is replaced.  A handle remembers the generation
    // that made ITS buffers the resident ones: setBVH / traceBatch compare it and upload again (or refuse) instead of silently
    // tracing whatever a later build left in the library.
    uint64_t getGeneration() const { return m_generation; }
    void setGeneration(uint64_t g) { m_generation = g; }
    bool hasHostCopy() { return m_nodes.getSize() > 0; }
    static uint64_t currentGeneration() { uint64_t g = 0; ntCheck(nt_bvh_generation(&g)); return g; }
    // a handle for the BVH that is resident in the library right now (a replica after nt_bvh_broadcast)
    static CudaBVH* adoptResident()
    {
        size_t sz[3]; int layout = 0;
        ntCheck(nt_bvh_sizes(sz, &layout));
        CudaBVH* b = new CudaBVH((BVHLayout)layout);
        b->m_resident = true;
        b->m_generation = currentGeneration();
        return b;
    }

protected:
    void materialise()
    {
        if (!m_resident || m_nodes.getSize()) return;
        if (m_generation && currentGeneration() != m_generation)
            fail("CudaBVH: this handle's BVH is no longer the resident one (a later build / upload replaced it before it was downloaded)");
        size_t sz[3]; int layout = 0;
        ntCheck(nt_bvh_sizes(sz, &layout));
        m_nodes.resizeDiscard((S64)sz[0]); m_triWoop.resizeDiscard((S64)sz[1]); m_triIndex.resizeDiscard((S64)sz[2]);
        ntCheck(nt_bvh_download(m_nodes.getMutableCudaPtrDiscard(), m_triWoop.getMutableCudaPtrDiscard(), (int32_t*)m_triIndex.getMutableCudaPtrDiscard()));
    }

    BVHLayout m_layout;
    bool m_resident;
    uint64_t m_generation;
    Buffer m_nodes, m_triWoop, m_triIndex;
};

struct HLBVHParams
{
    bool hlbvh; S32 hlbvhBits; S32 leafSize; F32 epsilon;
    HLBVHParams(bool h = true, S32 bits = 4, S32 leaf = 8, F32 eps = 0.001f) : hlbvh(h), hlbvhBits(bits), leafSize(leaf), epsilon(eps) {}
};

class HLBVHBuilder : public CudaBVH
{
public:
    HLBVHBuilder(Scene* scene, const HLBVHParams& params = HLBVHParams()) : CudaBVH(BVHLayout_Compact), m_gpuTime(0.0f)
    {
        if (!scene) fail("HLBVHBuilder: no scene");
        const bool lbvh = !params.hlbvh || params.hlbvhBits == 10;                                    // HLBVHBuilder.cpp:44-47
        Vec3f lo, hi;
        scene->getBBox(lo, hi);
        ntCheck(nt_bvh_build(lbvh ? NT_BUILDER_LBVH : NT_BUILDER_HLBVH, (const float*)scene->getVtxPosBuffer().getCudaPtr(), scene->getNumVertices(),
                             (const int32_t*)scene->getTriVtxIndexBuffer().getCudaPtr(), scene->getNumTriangles(), lo.getPtr(), hi.getPtr(),
                             params.hlbvhBits, params.leafSize, params.epsilon, &m_gpuTime));
        m_resident = true;
        m_generation = currentGeneration();
        size_t sz[3]; int layout = 0;
        ntCheck(nt_bvh_sizes(sz, &layout));                       // Compact, or Compact2 after nt_bvh_set_build_layout(5)
        m_layout = (BVHLayout)layout;
    }
    F32 getGPUTime() const { return m_gpuTime; }

private:
    F32 m_gpuTime;
};
}
