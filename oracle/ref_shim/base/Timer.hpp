// ORACLE BUILD SHIM — shadows base/Timer.hpp (QueryPerformanceCounter) with a std::chrono clock; same interface.
#pragma once
#include "base/DLLImports.hpp"
#include <chrono>
namespace FW
{
class Timer
{
public:
    explicit Timer(bool started = false) : m_start(started ? now() : -1.0), m_total(0.0) {}
    void start(void) { m_start = now(); }
    void unstart(void) { m_start = -1.0; }
    F32 getElapsed(void) { if (m_start < 0.0) m_start = now(); return (F32)(now() - m_start); }
    F32 end(void) { double t = now(); F32 e = (m_start < 0.0) ? 0.0f : (F32)(t - m_start); m_total += e; m_start = t; return e; }
    F32 getTotal(void) const { return (F32)m_total; }
    void clearTotal(void) { m_total = 0.0; }
    static void staticInit(void) {}
private:
    static double now(void) { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    double m_start, m_total;
};
}
