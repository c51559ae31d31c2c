// ORACLE — TEST INFRASTRUCTURE ONLY.  Runs the reference's OWN ray-sort kernels (src/rt/ray/RayBufferKernels.cu:
// findAABBKernel, genMortonKeysKernel), compiled unmodified from /root/reference for sm_100a, with the host side of
// RayBuffer::mortonSort (RayBuffer.cpp:103-163): constants filled, the launches it makes.  The 192-bit keys they produce
// are what the reference then sorts on the CPU; the B200 ray sort radix-sorts bits [83,147) of the same key.
// Built with the reference's -use_fast_math (libref_raybuf_fast.so) and with IEEE flags (libref_raybuf_ieee.so).
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
namespace refk {                       // the file says `using namespace FW;` at file scope, see ref_gpu_raygen.cu
#include "ray/RayBufferKernels.cu"
}
using refk::FW::FindAABBInput; using refk::FW::FindAABBOutput; using refk::FW::GenMortonKeysInput; using refk::FW::MortonKey;
using refk::c_FindAABBInput; using refk::c_FindAABBOutput; using refk::c_GenMortonKeysInput;
using refk::c_ReorderRaysInput;
using refk::findAABBKernel; using refk::genMortonKeysKernel; using refk::reorderRaysKernel;

static char s_err[512] = "";
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(s_err, sizeof(s_err), "%s: %s", #call, cudaGetErrorString(e_)); return 1; } } while (0)
extern "C" const char* ref_gpu_error(void) { return s_err; }

// findAABBKernel: box over ray origins and end points o + d * tmax -> outLoHi[6]
extern "C" int ref_ray_aabb(const void* dRays, int numRays, float* outLoHi)
{
    FindAABBInput in;
    in.numRays = numRays; in.raysPerThread = 32; in.inRays = (CUdeviceptr)dRays;
    CK(cudaMemcpyToSymbol(c_FindAABBInput, &in, sizeof(in)));
    const float big = 3.402823466e+38f;
    float init[8] = {big, big, big, -big, -big, -big, 0.0f, 0.0f};                 // FindAABBOutput{aabbLo, aabbHi}, padded to int4s
    CK(cudaMemcpyToSymbol(c_FindAABBOutput, init, sizeof(c_FindAABBOutput)));
    const int threads = (numRays - 1) / in.raysPerThread + 1;
    dim3 block(refk::FW::FindAABB_BlockWidth, refk::FW::FindAABB_BlockHeight);
    const int blocks = (threads + block.x * block.y - 1) / (block.x * block.y);
    findAABBKernel<<<blocks, block>>>();
    CK(cudaDeviceSynchronize());
    float out[8];
    CK(cudaMemcpyFromSymbol(out, c_FindAABBOutput, sizeof(c_FindAABBOutput)));
    for (int i = 0; i < 6; i++) outLoHi[i] = out[i];
    return 0;
}

// genMortonKeysKernel: hostKeys = numRays x {S32 oldSlot, U32 hash[6]}
extern "C" int ref_ray_keys(const void* dRays, int numRays, const float* lo, const float* hi, void* hostKeys)
{
    alignas(16) static unsigned char raw[sizeof(c_GenMortonKeysInput)];
    memset(raw, 0, sizeof(raw));
    GenMortonKeysInput& in = *reinterpret_cast<GenMortonKeysInput*>(raw);
    in.numRays = numRays;
    float* plo = reinterpret_cast<float*>(&in.aabbLo); float* phi = reinterpret_cast<float*>(&in.aabbHi);
    for (int i = 0; i < 3; i++) { plo[i] = lo[i]; phi[i] = hi[i]; }
    void* dKeys = nullptr;
    CK(cudaMalloc(&dKeys, (size_t)numRays * sizeof(MortonKey)));
    in.inRays = (CUdeviceptr)dRays; in.outKeys = (CUdeviceptr)dKeys;
    CK(cudaMemcpyToSymbol(c_GenMortonKeysInput, raw, sizeof(raw)));
    dim3 block(32, 4);
    genMortonKeysKernel<<<(numRays + 127) / 128, block>>>();
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hostKeys, dKeys, (size_t)numRays * sizeof(MortonKey), cudaMemcpyDeviceToHost));
    cudaFree(dKeys);
    return 0;
}
