// ORACLE BUILD SHIM — shadows bvh/OcclusionBVHBuilder.hpp (needs a rendered visibility pass; out of scope).  BVH::BVH only
// names the type in a branch the oracle never takes.
#pragma once
#include "bvh/BVH.hpp"
namespace FW
{
class OcclusionBVHBuilder
{
public:
    OcclusionBVHBuilder(BVH&, const BVH::BuildParams&) {}
    BVHNode* run(void) { fail("OcclusionBVH is not available in the oracle build"); return NULL; }
};
}
