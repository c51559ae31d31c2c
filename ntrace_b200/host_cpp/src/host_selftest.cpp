// Host-only checks of the C++ mirror (no GPU, no library calls): prints one JSON object that tests/test_host_cpp.py
// compares with the Python mirror and with the values SURVEY.md App. D lists.
//   host_selftest <config.conf> <mesh.obj> [-D...]
#include "ntrace/Base.hpp"
#include "ntrace/CameraControls.hpp"
#include "ntrace/Environment.hpp"
#include <fstream>
#include <map>
#include <sstream>
#include <tuple>

// Scene.hpp constructs device buffers; the loaders are static and GPU-free, so only they are pulled in here
#include "ntrace/Scene.hpp"
#include "ntrace/Renderer.hpp"
#include <iterator>

using namespace FW;

static void printMat(const char* name, const Mat4f& m)
{
    printf("\"%s\": [", name);
    for (int i = 0; i < 16; i++) printf("%s%.9g", i ? ", " : "", m.getPtr()[i]);
    printf("]");
}

int main(int argc, char** argv)
{
    try {
        // host_selftest --cache <in.dat> <out.dat>: the bvhcache stream through CudaBVH(std::istream&) and serialize() (no GPU needed)
        if (argc == 4 && std::string(argv[1]) == "--cache") {
            std::ifstream in(argv[2], std::ios::binary);
            if (!in) fail("cannot open %s", argv[2]);
            CudaBVH bvh(in);
            std::ofstream out(argv[3], std::ios::binary);
            bvh.serialize(out);
            printf("{\"layout\": %d, \"nodeBytes\": %lld, \"woopBytes\": %lld, \"idxBytes\": %lld}\n", (int)bvh.getLayout(),
                   (long long)bvh.getNodeBuffer().getSize(), (long long)bvh.getTriWoopBuffer().getSize(), (long long)bvh.getTriIndexBuffer().getSize());
            return 0;
        }
        // host_selftest --hash <file>: FW::hashBuffer of a file and the hashBits overloads on fixed arguments (cache file naming)
        if (argc == 3 && std::string(argv[1]) == "--hash") {
            std::ifstream in(argv[2], std::ios::binary);
            std::vector<char> d((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
            printf("{\"hashBuffer\": %u, \"hashBits1\": %u, \"hashBits3\": %u, \"hashBits4\": %u, \"hashBits6\": %u}\n",
                   Renderer::hashBuffer(d.data(), d.size()), Renderer::hashBits(12345u), Renderer::hashBits(1u, 2u, 3u),
                   Renderer::hashBits(1u, 2u, 3u, 4u), Renderer::hashBits(0xdeadbeefu, 7u, 0x80000000u, 4u, 5u, 6u));
            return 0;
        }
        if (argc < 3) fail("usage: host_selftest <config.conf> <mesh.obj> [-D...]");
        Environment env;
        std::vector<char*> args;
        args.push_back(argv[0]); args.push_back(argv[1]);
        for (int i = 3; i < argc; i++) args.push_back(argv[i]);
        env.Parse((int)args.size(), args.data());
        printf("{");
        const char* keys[] = {"App.frameWidth", "App.frameHeight", "Benchmark.scene", "Benchmark.camera", "Benchmark.kernel", "Benchmark.warmupRepeats",
                              "Benchmark.measureRepeats", "Renderer.rayType", "Renderer.samples", "Renderer.sortRays", "Renderer.builder", "Raygen.aoRadius",
                              "SubdivisionRayCaster.numPrimitives", "Renderer.dataStructure"};
        printf("\"env\": {");
        bool first = true;
        for (size_t i = 0; i < sizeof(keys) / sizeof(keys[0]); i++) {
            std::string v;
            if (!env.GetStringValue(keys[i], v)) continue;
            std::string esc;
            for (size_t k = 0; k < v.size(); k++) { if (v[k] == '"' || v[k] == '\\') esc += '\\'; esc += v[k]; }
            printf("%s\"%s\": \"%s\"", first ? "" : ", ", keys[i], esc.c_str());
            first = false;
        }
        printf("}, \"cameras\": {");
        const char* sigs[][2] = {{"conference", "6omr/04j3200bR6Z/0/3ZEAz/x4smy19///c/05frY109Qx7w////m100"},
                                 {"fairyforest", "cIxMx/sK/Ty/EFu3z/5m9mWx/YPA5z/8///m007toC10AnAHx///Uy200"},
                                 {"sibenik", "ytIa02G35kz1i:ZZ/0//iSay/5W6Ex19///c/05frY109Qx7w////m100"},
                                 {"sanmiguel", "Yciwz1oRQmz/Xvsm005CwjHx/b70nx18tVI7005frY108Y/:x/v3/z100"}};
        for (int i = 0; i < 4; i++) {
            CameraControls c;
            c.decodeSignature(sigs[i][1]);
            printf("%s\"%s\": {\"position\": [%.9g, %.9g, %.9g], \"forward\": [%.9g, %.9g, %.9g], \"up\": [%.9g, %.9g, %.9g], \"fov\": %.9g, \"near\": %.9g, \"far\": %.9g, ",
                   i ? ", " : "", sigs[i][0], c.getPosition().x, c.getPosition().y, c.getPosition().z, c.getForward().x, c.getForward().y, c.getForward().z,
                   c.getUp().x, c.getUp().y, c.getUp().z, c.getFOV(), c.getNear(), c.getFar());
            printMat("nscreenToWorld", c.getNScreenToWorld(1024, 768));
            printf(", ");
            printMat("worldToClip", c.getWorldToClip());
            printf("}");
        }
        printf("}, ");
        std::vector<Vec3f> verts; std::vector<Vec3i> tris;
        Scene::loadWavefront(argv[2], verts, tris);
        printf("\"verts\": [");
        for (size_t i = 0; i < verts.size(); i++) printf("%s[%.9g, %.9g, %.9g]", i ? ", " : "", verts[i].x, verts[i].y, verts[i].z);
        printf("], \"tris\": [");
        for (size_t i = 0; i < tris.size(); i++) printf("%s[%d, %d, %d]", i ? ", " : "", tris[i].x, tris[i].y, tris[i].z);
        printf("]}\n");
    } catch (const std::exception& e) {
        fprintf(stderr, "host_selftest: %s\n", e.what());
        return 1;
    }
    return 0;
}
