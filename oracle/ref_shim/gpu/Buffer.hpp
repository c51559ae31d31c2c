// ORACLE BUILD SHIM — shadows gpu/Buffer.hpp (CPU/GL/CUDA mirrored buffer) with a host-only byte buffer exposing the
// members the CPU path touches.
#pragma once
#include "base/Math.hpp"
#include "io/Stream.hpp"
#include <vector>
#include <cstring>
namespace FW
{
class Buffer
{
public:
    enum Module { CPU = 1, GL = 2, Cuda = 4 };
    Buffer(void) {}
    Buffer(const void* ptr, S64 size) { set(ptr, size); }
    S64 getSize(void) const { return (S64)m_data.size(); }
    void resize(S64 size) { m_data.resize((size_t)size); }
    void resizeDiscard(S64 size) { m_data.assign((size_t)size, 0); }
    void reset(void) { m_data.clear(); }
    void set(const void* ptr, S64 size) { m_data.resize((size_t)size); if (ptr && size) memcpy(m_data.data(), ptr, (size_t)size); }
    void clear(int value = 0) { if (!m_data.empty()) memset(m_data.data(), value, m_data.size()); }
    const U8* getPtr(S64 ofs = 0) const { return m_data.data() + ofs; }
    U8* getMutablePtr(S64 ofs = 0) { return m_data.data() + ofs; }
    void setOwner(Module, bool) {}
private:
    std::vector<U8> m_data;
};
// Buffer::readFromStream / writeToStream (gpu/Buffer.cpp:349-381): S64 size, then the bytes
inline InputStream& operator>>(InputStream& s, Buffer& b) { S64 n = 0; s >> n; b.resizeDiscard(n); if (n) s.readFully(b.getMutablePtr(), (int)n); return s; }
inline OutputStream& operator<<(OutputStream& s, const Buffer& b) { s << b.getSize(); if (b.getSize()) s.write(b.getPtr(), (int)b.getSize()); return s; }
}
