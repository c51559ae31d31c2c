// ORACLE BUILD SHIM — shadows src/rt/Environment.h with the one query BVH::BVH makes ("Renderer.builder").
#pragma once
#include <string>
using std::string;   // MSVC headers leak std names into the reference's translation units (BVH.cpp:38 writes `string`)
class Environment
{
public:
    static Environment* GetSingleton(void) { static Environment e; return &e; }
    bool GetStringValue(const char* name, std::string& value) const { if (std::string(name) == "Renderer.builder") { value = builder; return true; } return false; }
    std::string builder = "SplitBVH";
};
