// ORACLE — TEST INFRASTRUCTURE ONLY.  Runs the reference's OWN ray generation kernels (src/rt/ray/RayGenKernels.cu:
// rayGenPrimaryKernel, rayGenAOKernel, rayGenShadowKernel), compiled unmodified from /root/reference for sm_100a, with the
// host side of RayGen::primary / ao / shadow (RayGen.cpp:45-74, 114-147, 198-232): fill the __constant__ input struct, launch
// one thread per input ray.  Built twice by `make -C oracle ref_gpu`: with the reference's -use_fast_math
// (libref_raygen_fast.so) and with IEEE flags (-fmad=false, libref_raygen_ieee.so: the same source without the approximate
// division / sine, which is the arithmetic the B200 kernels and the restated oracle use).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdarg>
#include <ctime>
#include <new>
#include <string>
#include <fstream>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <stdarg.h>
#include <time.h>
#include <vector_types.h>
#include <vector_functions.h>
// RayGenKernels.cu says `using namespace FW;` at file scope; FW:: declares its own sqrt/sin/exp overloads, which makes the
// host-side math stubs nvcc appends to every translation unit ambiguous.  Confining the file (and the framework headers it
// pulls in, all of whose system includes are already satisfied above) to a namespace keeps that directive local.
namespace refk {
#include "ray/RayGenKernels.cu"
}
using refk::FW::RayGenPrimaryInput; using refk::FW::RayGenAOInput; using refk::FW::RayGenShadowInput;
using refk::c_RayGenPrimaryInput; using refk::c_RayGenAOInput; using refk::c_RayGenShadowInput;
using refk::rayGenPrimaryKernel; using refk::rayGenAOKernel; using refk::rayGenShadowKernel;

static char s_err[512] = "";
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(s_err, sizeof(s_err), "%s: %s", #call, cudaGetErrorString(e_)); return 1; } } while (0)
extern "C" const char* ref_gpu_error(void) { return s_err; }

static dim3 gridFor(long long threads, dim3 block) { return dim3((unsigned)((threads + block.x * block.y - 1) / (block.x * block.y))); }

extern "C" int ref_raygen_primary(void* dRays, void* dIdToSlot, void* dSlotToID, const void* dIndexToPixel, const float* origin, const float* n2wRowMajor,
                                  int w, int h, float maxDist, unsigned randomSeed)
{
    // the FW:: vector types have __device__-only members under nvcc: fill the struct through raw storage on the host
    alignas(16) static unsigned char raw[sizeof(RayGenPrimaryInput)];
    RayGenPrimaryInput& in = *reinterpret_cast<RayGenPrimaryInput*>(raw);
    float* o = reinterpret_cast<float*>(&in.origin);
    float* m = reinterpret_cast<float*>(&in.nscreenToWorld);            // Mat4f stores columns: m00, m10, m20, m30, m01, ... (Math.hpp:700-715)
    for (int i = 0; i < 3; i++) o[i] = origin[i];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) m[c * 4 + r] = n2wRowMajor[r * 4 + c];
    in.w = w; in.h = h; in.maxDist = maxDist;
    in.rays = (CUdeviceptr)dRays; in.idToSlot = (CUdeviceptr)dIdToSlot; in.slotToID = (CUdeviceptr)dSlotToID; in.indexToPixel = (CUdeviceptr)dIndexToPixel;
    in.randomSeed = randomSeed;
    CK(cudaMemcpyToSymbol(c_RayGenPrimaryInput, &in, sizeof(in)));
    dim3 block(32, 4);
    rayGenPrimaryKernel<<<gridFor((long long)w * h, block), block>>>();
    CK(cudaDeviceSynchronize());
    return 0;
}

extern "C" int ref_raygen_ao(void* dOutRays, void* dOutIDToSlot, void* dOutSlotToID, const void* dInRays, const void* dInResults, const void* dNormals,
                             int firstInputSlot, int numInputRays, int numSamples, float maxDist, unsigned randomSeed)
{
    RayGenAOInput in;
    in.firstInputSlot = firstInputSlot; in.numInputRays = numInputRays; in.numSamples = numSamples; in.maxDist = maxDist; in.randomSeed = randomSeed;
    in.inRays = (CUdeviceptr)dInRays; in.inResults = (CUdeviceptr)dInResults; in.outRays = (CUdeviceptr)dOutRays;
    in.outIDToSlot = (CUdeviceptr)dOutIDToSlot; in.outSlotToID = (CUdeviceptr)dOutSlotToID; in.normals = (CUdeviceptr)dNormals;
    CK(cudaMemcpyToSymbol(c_RayGenAOInput, &in, sizeof(in)));
    dim3 block(32, 4);
    rayGenAOKernel<<<gridFor(numInputRays, block), block>>>();
    CK(cudaDeviceSynchronize());
    return 0;
}

extern "C" int ref_raygen_shadow(void* dOutRays, void* dOutIDToSlot, void* dOutSlotToID, const void* dInRays, const void* dInResults,
                                 int firstInputSlot, int numInputRays, int numSamples, const float* lightPos, float lightRadius, unsigned randomSeed)
{
    RayGenShadowInput in;
    in.firstInputSlot = firstInputSlot; in.numInputRays = numInputRays; in.numSamples = numSamples;
    in.lightPositionX = lightPos[0]; in.lightPositionY = lightPos[1]; in.lightPositionZ = lightPos[2]; in.lightRadius = lightRadius; in.randomSeed = randomSeed;
    in.inRays = (CUdeviceptr)dInRays; in.inResults = (CUdeviceptr)dInResults; in.outRays = (CUdeviceptr)dOutRays;
    in.outIDToSlot = (CUdeviceptr)dOutIDToSlot; in.outSlotToID = (CUdeviceptr)dOutSlotToID;
    CK(cudaMemcpyToSymbol(c_RayGenShadowInput, &in, sizeof(in)));
    dim3 block(32, 4);
    rayGenShadowKernel<<<gridFor(numInputRays, block), block>>>();
    CK(cudaDeviceSynchronize());
    return 0;
}
