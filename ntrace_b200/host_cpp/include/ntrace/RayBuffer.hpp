// FW::RayBuffer — rays, results and the id <-> slot maps.  Reference: src/rt/ray/RayBuffer.hpp:38-195, RayBuffer.cpp:38-163.
#pragma once
#include "ntrace/Buffer.hpp"

namespace FW
{
class RayBuffer
{
public:
    explicit RayBuffer(S32 n = 0, bool closestHit = true) : m_size(0), m_needClosestHit(closestHit) { resize(n); }

    S32 getSize() const { return m_size; }
    void resize(S32 n)                                       // storage only grows (RayBuffer.cpp:38-51)
    {
        if (n < 0) fail("RayBuffer: negative size");
        if (n < m_size) { m_size = n; return; }
        m_size = n;
        if (m_rays.getSize() < (S64)n * (S64)sizeof(Ray)) {
            m_rays.resize((S64)n * sizeof(Ray));
            m_results.resize((S64)n * sizeof(RayResult));
            m_IDToSlot.resize((S64)n * sizeof(S32));
            m_slotToID.resize((S64)n * sizeof(S32));
        }
    }
    void reserve(S32 n) { const S32 s = m_size; if (n > s) { resize(n); m_size = s; } }       // grow the storage, keep the size
    void setRay(S32 slot, const Ray& ray) { setRay(slot, ray, slot); }
    void setRay(S32 slot, const Ray& ray, S32 id)            // RayBuffer.cpp:55-62
    {
        if (slot < 0 || slot >= m_size || id < 0 || id >= m_size) fail("RayBuffer: slot out of range");
        ((Ray*)m_rays.getMutablePtr())[slot] = ray;
        ((S32*)m_IDToSlot.getMutablePtr())[id] = slot;
        ((S32*)m_slotToID.getMutablePtr())[slot] = id;
    }
    void setResult(S32 slot, const RayResult& r) { getMutableResultForSlot(slot) = r; }

    const Ray& getRayForSlot(S32 slot) const { return ((const Ray*)m_rays.getPtr())[slot]; }
    const Ray& getRayForID(S32 id) const { return getRayForSlot(getSlotForID(id)); }
    const RayResult& getResultForSlot(S32 slot) const { return ((const RayResult*)m_results.getPtr())[slot]; }
    RayResult& getMutableResultForSlot(S32 slot) { return ((RayResult*)m_results.getMutablePtr())[slot]; }
    const RayResult& getResultForID(S32 id) const { return getResultForSlot(getSlotForID(id)); }
    S32 getSlotForID(S32 id) const { return ((const S32*)m_IDToSlot.getPtr())[id]; }
    S32 getIDForSlot(S32 slot) const { return ((const S32*)m_slotToID.getPtr())[slot]; }

    void setNeedClosestHit(bool c) { m_needClosestHit = c; }
    bool getNeedClosestHit() const { return m_needClosestHit; }

    // RayBuffer::mortonSort (RayBuffer.cpp:103-163): on the GPU here (nt_ray_sort)
    void mortonSort()
    {
        if (m_size > 1)
            ntCheck(nt_ray_sort((float*)m_rays.getMutableCudaPtr(), (int32_t*)m_IDToSlot.getMutableCudaPtr(),
                                (int32_t*)m_slotToID.getMutableCudaPtr(), m_size));
    }

    Buffer& getRayBuffer() { return m_rays; }
    Buffer& getResultBuffer() { return m_results; }
    Buffer& getIDToSlotBuffer() { return m_IDToSlot; }
    Buffer& getSlotToIDBuffer() { return m_slotToID; }

private:
    S32 m_size;
    bool m_needClosestHit;
    Buffer m_rays, m_results, m_IDToSlot, m_slotToID;
};
}
