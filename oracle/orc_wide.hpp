// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU emulation of the product's Wide4 traversal (ntrace_b200/csrc/nt_wide.cu).  The Wide4 node array is NOT part of the reference
// (NTrace has binary trees only); it is a derived form of the reference's Compact CudaBVH.  This file therefore restates the
// PRODUCT's node decode and child ordering operation for operation (so the CUDA kernel can be required to be bit-identical to it),
// while the triangle test, the leaf format and the result convention are the reference's (src/rt/cuda/CudaBVH.cpp:1083-1126,
// 1183-1225 via orc::ray_triangle_woop).  Parity of the Wide4 path against the reference is then established by comparing this
// emulation with orc::trace_compact on the same Compact BVH (tests/test_wide4.py).
#pragma once
#include "orc_bvh.hpp"

namespace orc {

// results: (id, t) as trace_compact; counters (optional): per ray [wide nodes visited, triangles tested, leaves entered]
void trace_wide4(const uint32_t* wnodes, const int32_t* woop, const int32_t* triIndex,
                 const Ray* rays, RayResult* results, int n, bool closest, uint32_t* counters, int nthreads);

// structural check of a Wide4 array against the Compact tree it was derived from: every leaf link of the binary tree appears exactly
// once, every decoded child box contains the binary tree's box of the same subtree.  Returns 0 when consistent, else an error code;
// out[0] = wide nodes, out[1] = leaf links, out[2] = max depth, out[3] = worst overhang of a decoded box in quantisation steps
int check_wide4(const uint32_t* wnodes, size_t numWide, const int32_t* nodes, size_t nodeBytes, int layout, double out[4]);

} // namespace orc
