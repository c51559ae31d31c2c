// ORACLE BUILD SHIM — shadows src/rt/Scene.hpp (mesh import, materials, texture atlas) with the flat buffers the CPU
// builders and tracers read: triVtxIndex (Vec3i per triangle) and vtxPos (Vec3f per vertex).
#pragma once
#include "base/Math.hpp"
#include "gpu/Buffer.hpp"
namespace FW
{
class Scene
{
public:
    Scene(const Vec3f* verts, int numVerts, const Vec3i* tris, int numTris)
    : m_numTriangles(numTris), m_numVertices(numVerts), m_triVtxIndex(tris, (S64)numTris * sizeof(Vec3i)), m_vtxPos(verts, (S64)numVerts * sizeof(Vec3f)) {}
    int getNumTriangles(void) const { return m_numTriangles; }
    int getNumVertices(void) const { return m_numVertices; }
    Buffer& getTriVtxIndexBuffer(void) { return m_triVtxIndex; }
    Buffer& getVtxPosBuffer(void) { return m_vtxPos; }
private:
    int m_numTriangles, m_numVertices;
    Buffer m_triVtxIndex, m_vtxPos;
};
}
