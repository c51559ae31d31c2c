// ORACLE BUILD SHIM — shadows base/MulticoreLauncher.hpp (Win32 threads).  base/Sort.cpp only needs the Task type and
// push(); the shim runs pushed tasks inline (the builders call sort() with multicore == false anyway).
#pragma once
#include "base/Defs.hpp"
namespace FW
{
class MulticoreLauncher
{
public:
    struct Task;
    typedef void (*TaskFunc)(Task& task);
    struct Task { MulticoreLauncher* launcher; TaskFunc func; void* data; int idx; void* result; };
    void push(TaskFunc func, void* data, int firstIdx = 0, int numTasks = 1)
    {
        for (int i = 0; i < numTasks; i++) { Task t; t.launcher = this; t.func = func; t.data = data; t.idx = firstIdx + i; t.result = NULL; func(t); }
    }
};
}
