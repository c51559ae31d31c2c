// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  pixel_table PINNED against the compiled reference (oracle/_ref, CPU); ray generators PINNED on the GPU box: bit-identical to the IEEE build of the reference's RayGenKernels.cu run on the B200 (tests/test_gpu_reference_kernels.py).
#pragma once
#include "orc_bvh.hpp"

namespace orc {

void pixel_table(int w, int h, int32_t* indexToPixel, int32_t* pixelToIndex);                 // PixelTable.cpp:57-141
void raygen_primary(Ray* rays, int32_t* idToSlot, int32_t* slotToID, V3 origin, const M4& nscreenToWorld,
                    int w, int h, float maxDist, uint32_t randomSeed);                          // RayGen.cpp:76-112
void raygen_ao(Ray* outRays, int32_t* outIDToSlot, int32_t* outSlotToID,
               const Ray* inRays, const RayResult* inResults, const V3* normals,
               int firstInputSlot, int numInputRays, int numSamples, float maxDist, uint32_t randomSeed);  // RayGenKernels.cu:129-236
void raygen_shadow(Ray* outRays, int32_t* outIDToSlot, int32_t* outSlotToID, const Ray* inRays, const RayResult* inResults,
                   int firstInputSlot, int numInputRays, int numSamples, V3 lightPos, float lightRadius, uint32_t randomSeed);  // RayGenKernels.cu:240-302
int  count_hits(const RayResult* results, int n);                                              // RendererKernels.cu:174-224
void tri_normals(const Scene& sc, V3* out);                                                    // Scene.cpp:112

} // namespace orc

namespace orc {
void ray_morton_keys(const Ray* rays, int n, uint32_t* keys6, float aabb[6]);                     // RayBufferKernels.cu:62-163
void ray_morton_order(const Ray* rays, int n, int truncated, int32_t* order, uint64_t* key64Out);  // RayBuffer.cpp:88-149
}
