// ORACLE BUILD SHIM — test infrastructure only.  Shadows src/framework/base/DLLImports.hpp of the reference (which pulls in
// <windows.h>, <mmsystem.h>, <shlwapi.h>) so that the reference's *algorithm* sources compile unmodified with g++ on Linux.
// No algorithm lives here: only the platform layer the reference takes from CUDA / Win32 / MSVC headers:
//   * CUDA vector types and CUdeviceptr come from the toolkit headers;
//   * <string.h>/<stdlib.h> are visible transitively under MSVC;
//   * MSVC's one-phase template lookup lets Hash.hpp call hash<T>() before declaring it: forward-declared here.
#pragma once
#include "base/Defs.hpp"
#include <string.h>
#include <stdlib.h>
#include <stdarg.h>
#include <stdio.h>
#include <time.h>
#include <cuda.h>
#include <vector_types.h>
#include <vector_functions.h>
// MSVC CRT spellings used by base/String.cpp
static inline int _vscprintf(const char* fmt, va_list ap) { va_list c; va_copy(c, ap); int n = vsnprintf(NULL, 0, fmt, c); va_end(c); return n; }
static inline int vsprintf_s(char* buf, size_t n, const char* fmt, va_list ap) { return vsnprintf(buf, n, fmt, ap); }
static inline int ctime_s(char* buf, size_t n, const time_t* t) { char tmp[32]; ctime_r(t, tmp); strncpy(buf, tmp, n); if (n) buf[n - 1] = 0; return 0; }
static inline int vsnprintf_s(char* buf, size_t n, size_t, const char* fmt, va_list ap) { return vsnprintf(buf, n, fmt, ap); }
// Win32 high-resolution counter, used by base/Random.cpp only to pick a seed when none is given
typedef union { long long QuadPart; struct { unsigned LowPart; int HighPart; }; } LARGE_INTEGER;
static inline int QueryPerformanceCounter(LARGE_INTEGER* t) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); t->QuadPart = (long long)ts.tv_sec * 1000000000LL + ts.tv_nsec; return 1; }
namespace FW
{
class String;
template <class T> inline U32 hash(const T& value);
}
