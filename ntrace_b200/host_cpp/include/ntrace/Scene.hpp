// FW::Scene — the flat scene buffers the tracing path reads, plus mesh ingestion with the reference's ordering rules.
// Reference: src/rt/Scene.hpp:110-190 / Scene.cpp:101-136 (triVtxIndex, vtxPos, triNormal; submeshes concatenated in
// order), src/framework/io/MeshWavefrontIO.cpp:412-485 (vertex dedup by (pos,tex,normal) index triple in first-use order,
// fan triangulation, faces flushed per usemtl, default submesh for faces without a known material).
#pragma once
#include "ntrace/Buffer.hpp"
#include <fstream>
#include <map>
#include <sstream>
#include <tuple>

namespace FW
{
class Scene
{
public:
    Scene(const std::vector<Vec3f>& verts, const std::vector<Vec3i>& tris)
    : m_numTriangles((S32)tris.size()), m_numVertices((S32)verts.size())
    {
        if (verts.empty() || tris.empty()) fail("Scene: empty mesh");
        m_vtxPos.set(verts.data(), (S64)verts.size() * sizeof(Vec3f));
        m_triVtxIndex.set(tris.data(), (S64)tris.size() * sizeof(Vec3i));
        m_triNormal.resizeDiscard((S64)tris.size() * sizeof(Vec3f));
        // Scene.cpp:112: normalize(cross(v1 - v0, v2 - v0)); computed on the device
        ntCheck(nt_tri_normals((const float*)m_vtxPos.getCudaPtr(), m_numVertices, (const int32_t*)m_triVtxIndex.getCudaPtr(),
                               m_numTriangles, (float*)m_triNormal.getMutableCudaPtrDiscard()));
        for (size_t i = 0; i < verts.size(); i++) m_bbox.grow(verts[i]);
    }

    int getNumTriangles() const { return m_numTriangles; }
    int getNumVertices() const { return m_numVertices; }
    Buffer& getTriVtxIndexBuffer() { return m_triVtxIndex; }
    Buffer& getTriNormalBuffer() { return m_triNormal; }
    Buffer& getVtxPosBuffer() { return m_vtxPos; }
    void getBBox(Vec3f& lo, Vec3f& hi) const { lo = m_bbox.min(); hi = m_bbox.max(); }

    // ---- ingestion ----------------------------------------------------------------------------------------------
    static Scene* importMesh(const std::string& path)
    {
        std::vector<Vec3f> v; std::vector<Vec3i> t;
        if (path.size() > 7 && path.compare(path.size() - 7, 7, ".ntmesh") == 0) loadNtMesh(path, v, t);
        else loadWavefront(path, v, t);
        return new Scene(v, t);
    }

    // binary container written by ntrace_b200/mesh_io.py: "NTMESH1\0", int64 numVerts, int64 numTris, verts f32x3, tris i32x3
    static void loadNtMesh(const std::string& path, std::vector<Vec3f>& verts, std::vector<Vec3i>& tris)
    {
        std::ifstream f(path.c_str(), std::ios::binary);
        if (!f) fail("Cannot open file '%s'!", path.c_str());
        char magic[8]; int64_t nv = 0, nt = 0;
        f.read(magic, 8); f.read((char*)&nv, 8); f.read((char*)&nt, 8);
        if (!f || memcmp(magic, "NTMESH1\0", 8) != 0 || nv <= 0 || nt <= 0) fail("'%s' is not an .ntmesh file", path.c_str());
        verts.resize((size_t)nv); tris.resize((size_t)nt);
        f.read((char*)verts.data(), nv * 12); f.read((char*)tris.data(), nt * 12);
        if (!f) fail("'%s' is truncated", path.c_str());
    }

    static void loadWavefront(const std::string& path, std::vector<Vec3f>& verts, std::vector<Vec3i>& tris)
    {
        std::ifstream f(path.c_str());
        if (!f) fail("Cannot open file '%s'!", path.c_str());
        std::string dir = path.substr(0, path.find_last_of("/\\") == std::string::npos ? 0 : path.find_last_of("/\\") + 1);
        std::vector<Vec3f> positions;
        int numTex = 0, numNrm = 0;
        std::map<std::tuple<int, int, int>, int> vertHash;
        std::vector<std::vector<Vec3i> > submeshes;
        std::map<std::string, int> materials;                  // name -> submesh index, -1 = known but unused so far
        int submesh = -1, defaultSubmesh = -1;
        std::vector<Vec3i> indexTmp;
        auto flush = [&]() { if (submesh != -1) submeshes[submesh].insert(submeshes[submesh].end(), indexTmp.begin(), indexTmp.end()); indexTmp.clear(); };
        auto trim = [](std::string s) { size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n"); return a == std::string::npos ? std::string() : s.substr(a, b - a + 1); };
        std::string raw;
        while (std::getline(f, raw)) {
            std::string line = trim(raw);
            if (line.empty() || line[0] == '#') continue;
            if (line.compare(0, 2, "v ") == 0) {
                std::istringstream ss(line.substr(2));
                double x = 0.0, y = 0.0, z = 0.0; ss >> x >> y >> z;      // double, then one rounding to fp32
                positions.push_back(Vec3f((F32)x, (F32)y, (F32)z));
            } else if (line.compare(0, 3, "vt ") == 0) numTex++;
            else if (line.compare(0, 3, "vn ") == 0) numNrm++;
            else if (line.compare(0, 2, "f ") == 0) {
                std::istringstream ss(line.substr(2));
                std::string tok;
                std::vector<int> tmp;
                const int sizes[3] = {(int)positions.size(), numTex, numNrm};
                while (ss >> tok) {
                    int ptn[3] = {0, 0, 0};
                    size_t start = 0;
                    for (int i = 0; i < 3; i++) {
                        int v = 0;
                        if (start <= tok.size()) {
                            size_t end = tok.find('/', start);
                            std::string part = tok.substr(start, end == std::string::npos ? std::string::npos : end - start);
                            if (!part.empty()) v = atoi(part.c_str());
                            start = (end == std::string::npos) ? tok.size() + 1 : end + 1;
                        }
                        v = (v < 0) ? v + sizes[i] : v - 1;
                        if (v < 0 || v >= sizes[i]) v = -1;
                        ptn[i] = v;
                    }
                    auto key = std::make_tuple(ptn[0], ptn[1], ptn[2]);
                    auto it = vertHash.find(key);
                    int idx;
                    if (it == vertHash.end()) {
                        idx = (int)verts.size();
                        vertHash[key] = idx;
                        verts.push_back(ptn[0] != -1 ? positions[ptn[0]] : Vec3f(0.0f));
                    } else idx = it->second;
                    tmp.push_back(idx);
                }
                if (submesh == -1) {
                    if (defaultSubmesh == -1) { defaultSubmesh = (int)submeshes.size(); submeshes.push_back(std::vector<Vec3i>()); }
                    submesh = defaultSubmesh;
                }
                for (size_t i = 2; i < tmp.size(); i++) indexTmp.push_back(Vec3i(tmp[0], tmp[i - 1], tmp[i]));
            } else if (line.compare(0, 7, "usemtl ") == 0) {
                std::string name = trim(line.substr(6));
                if (submesh != -1) { flush(); submesh = -1; }
                auto it = materials.find(name);
                if (it != materials.end()) {
                    if (it->second == -1) { it->second = (int)submeshes.size(); submeshes.push_back(std::vector<Vec3i>()); }
                    submesh = it->second;
                    indexTmp.clear();
                }
            } else if (line.compare(0, 7, "mtllib ") == 0) {
                std::ifstream m((dir + trim(line.substr(6))).c_str());
                std::string ml;
                while (std::getline(m, ml)) {
                    ml = trim(ml);
                    if (ml.compare(0, 7, "newmtl ") == 0 || ml == "newmtl") { std::string n = trim(ml.substr(6)); if (!materials.count(n)) materials[n] = -1; }
                }
            }
        }
        flush();
        for (size_t s = 0; s < submeshes.size(); s++) tris.insert(tris.end(), submeshes[s].begin(), submeshes[s].end());
        if (verts.empty() || tris.empty()) fail("'%s' holds no triangles", path.c_str());
    }

private:
    S32 m_numTriangles, m_numVertices;
    Buffer m_triVtxIndex, m_triNormal, m_vtxPos;
    AABB m_bbox;
};
}
