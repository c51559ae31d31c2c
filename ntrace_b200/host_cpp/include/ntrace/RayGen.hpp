// FW::RayGen — primary / AO (diffuse) / shadow ray batches.  Reference: src/rt/ray/RayGen.hpp:47-170, RayGen.cpp:45-74
// (primary), :114-147 (shadow), :198-232 (ao), :582-600 (batching).
#pragma once
#include "ntrace/RayBuffer.hpp"
#include "ntrace/Scene.hpp"

namespace FW
{
class RayGen
{
public:
    explicit RayGen(S32 maxBatchSize = 8 * 1024 * 1024) : m_maxBatchSize(maxBatchSize), m_shadowStartIdx(0), m_aoStartIdx(0) {}

    void primary(RayBuffer& orays, const Vec3f& origin, const Mat4f& nscreenToWorld, S32 w, S32 h, float maxDist, U32 randomSeed = 0)
    {
        orays.resize(w * h);
        orays.setNeedClosestHit(true);
        ntCheck(nt_raygen_primary((float*)orays.getRayBuffer().getMutableCudaPtrDiscard(), (int32_t*)orays.getIDToSlotBuffer().getMutableCudaPtrDiscard(),
                                  (int32_t*)orays.getSlotToIDBuffer().getMutableCudaPtrDiscard(), origin.getPtr(), nscreenToWorld.getPtr(), w, h, maxDist, randomSeed));
    }

    // advance the batching of ao() by one batch without generating it (multi-GPU: the batch belongs to another rank); false at the end
    bool skipAo(RayBuffer& irays, int numSamples, bool& newBatch)
    {
        S32 lo, hi;
        return batching(irays.getSize(), numSamples, m_aoStartIdx, newBatch, lo, hi);
    }

    // true while the batch continues (reference signature: newBatch is in/out)
    bool ao(RayBuffer& orays, RayBuffer& irays, Scene& scene, int numSamples, float maxDist, bool& newBatch, U32 randomSeed = 0)
    {
        S32 lo, hi;
        if (!batching(irays.getSize(), numSamples, m_aoStartIdx, newBatch, lo, hi)) return false;
        orays.resize((hi - lo) * numSamples);
        orays.setNeedClosestHit(false);
        ntCheck(nt_raygen_ao((float*)orays.getRayBuffer().getMutableCudaPtrDiscard(), (int32_t*)orays.getIDToSlotBuffer().getMutableCudaPtrDiscard(),
                             (int32_t*)orays.getSlotToIDBuffer().getMutableCudaPtrDiscard(), (const float*)irays.getRayBuffer().getCudaPtr(),
                             (const int32_t*)irays.getResultBuffer().getCudaPtr(), (const float*)scene.getTriNormalBuffer().getCudaPtr(),
                             lo, hi - lo, numSamples, maxDist, randomSeed));
        return true;
    }

    bool shadow(RayBuffer& orays, RayBuffer& irays, int numSamples, const Vec3f& lightPos, float lightRadius, bool& newBatch, U32 randomSeed = 0)
    {
        S32 lo, hi;
        if (!batching(irays.getSize(), numSamples, m_shadowStartIdx, newBatch, lo, hi)) return false;
        orays.resize((hi - lo) * numSamples);
        orays.setNeedClosestHit(false);
        ntCheck(nt_raygen_shadow((float*)orays.getRayBuffer().getMutableCudaPtrDiscard(), (int32_t*)orays.getIDToSlotBuffer().getMutableCudaPtrDiscard(),
                                 (int32_t*)orays.getSlotToIDBuffer().getMutableCudaPtrDiscard(), (const float*)irays.getRayBuffer().getCudaPtr(),
                                 (const int32_t*)irays.getResultBuffer().getCudaPtr(), lo, hi - lo, numSamples, lightPos.getPtr(), lightRadius, randomSeed));
        return true;
    }

    // RayGen.cpp:582-600
    bool batching(S32 numInputRays, S32 numSamples, S32& startIdx, bool& newBatch, S32& lo, S32& hi) const
    {
        if (newBatch) { newBatch = false; startIdx = 0; }
        if (startIdx == numInputRays) return false;
        lo = startIdx;
        hi = lo + m_maxBatchSize / numSamples;
        if (hi > numInputRays) hi = numInputRays;
        startIdx = hi;
        return true;
    }

private:
public:
    S32 getMaxBatchSize() const { return m_maxBatchSize; }
private:
    S32 m_maxBatchSize;
    S32 m_shadowStartIdx, m_aoStartIdx;
};
}
