// FW::Buffer — byte buffer mirrored between the CPU and the GPU with lazy migration.
// Reference: src/framework/gpu/Buffer.hpp:58-190, Buffer.cpp:383-520 (getPtr / getMutablePtr / getCudaPtr /
// getMutableCudaPtr: the side asked for is brought up to date, the "Mutable" forms mark the other side stale).
// The GL module of the reference buffer is not on the tracing path and is not mirrored.
#pragma once
#include "ntrace/Base.hpp"

namespace FW
{
class Buffer
{
public:
    enum Module { CPU = 1 << 0, Cuda = 1 << 2 };

    Buffer() : m_size(0), m_cpu(nullptr), m_cpuCap(0), m_cuda(nullptr), m_cudaCap(0), m_valid(CPU) {}
    Buffer(const void* ptr, S64 size) : Buffer() { set(ptr, size); }
    Buffer(const Buffer& o) : Buffer() { *this = o; }
    ~Buffer() { release(); }
    Buffer& operator=(const Buffer& o)
    {
        if (&o != this) { resizeDiscard(o.m_size); if (m_size) memcpy(getMutablePtr(), o.getPtr(), (size_t)m_size); }
        return *this;
    }

    S64 getSize() const { return m_size; }
    void reset() { release(); }
    void resizeDiscard(S64 size) { m_size = size; m_valid = CPU; ensureCpu(size, false); }
    void resize(S64 size)
    {
        if (size == m_size) return;
        if (size > m_size) {                                 // keep contents on whichever side holds them (Buffer.cpp:205-240)
            if (m_valid & CPU) ensureCpu(size, true);
            if (m_valid & Cuda) ensureCuda(size, true);
        }
        m_size = size;
    }
    void set(const void* ptr, S64 size) { resizeDiscard(size); if (ptr && size) memcpy(m_cpu, ptr, (size_t)size); }
    void clear(int value = 0) { ensureCpu(m_size, false); if (m_size) memset(m_cpu, value, (size_t)m_size); m_valid = CPU; }

    const U8* getPtr(S64 ofs = 0) const { const_cast<Buffer*>(this)->validateCpu(); return m_cpu + ofs; }
    U8* getMutablePtr(S64 ofs = 0) { validateCpu(); m_valid = CPU; return m_cpu + ofs; }
    const void* getCudaPtr(S64 ofs = 0) const { const_cast<Buffer*>(this)->validateCuda(); return (const U8*)m_cuda + ofs; }
    void* getMutableCudaPtr(S64 ofs = 0) { validateCuda(); m_valid = Cuda; return (U8*)m_cuda + ofs; }
    // like getMutableCudaPtr() when the old contents are about to be overwritten completely: no upload
    void* getMutableCudaPtrDiscard() { ensureCuda(m_size, false); m_valid = Cuda; return m_cuda; }

private:
    void release()
    {
        if (m_cpu) { free(m_cpu); m_cpu = nullptr; m_cpuCap = 0; }
        if (m_cuda) { nt_mem_free(m_cuda); m_cuda = nullptr; m_cudaCap = 0; }
        m_size = 0; m_valid = CPU;
    }
    void ensureCpu(S64 size, bool keep)
    {
        if (size <= m_cpuCap) return;
        U8* p = (U8*)malloc((size_t)size);
        if (!p) fail("Out of memory!");
        if (keep && m_cpu && m_size) memcpy(p, m_cpu, (size_t)m_size);
        free(m_cpu);
        m_cpu = p; m_cpuCap = size;
    }
    void ensureCuda(S64 size, bool keep)
    {
        if (size <= m_cudaCap) return;
        void* p = nullptr;
        ntCheck(nt_mem_alloc((size_t)size, &p));
        if (keep && m_cuda && m_size) ntCheck(nt_memcpy(p, m_cuda, (size_t)m_size));
        if (m_cuda) ntCheck(nt_mem_free(m_cuda));
        m_cuda = p; m_cudaCap = size;
    }
    void validateCpu()
    {
        ensureCpu(m_size, (m_valid & CPU) != 0);
        if (!(m_valid & CPU)) { if (m_size) ntCheck(nt_memcpy(m_cpu, m_cuda, (size_t)m_size)); m_valid |= CPU; }
    }
    void validateCuda()
    {
        ensureCuda(m_size, (m_valid & Cuda) != 0);
        if (!(m_valid & Cuda)) { if (m_size) ntCheck(nt_memcpy(m_cuda, m_cpu, (size_t)m_size)); m_valid |= Cuda; }
    }

    S64 m_size;
    U8* m_cpu; S64 m_cpuCap;
    void* m_cuda; S64 m_cudaCap;
    int m_valid;              // which side(s) hold the current contents
};
}
