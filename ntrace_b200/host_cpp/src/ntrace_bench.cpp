// ntrace_bench — the reference's benchmark mode on the B200 tracing path, in C++ over the C ABI.
//
// Mirror of FW::runBenchmark (src/rt/App.cpp:842-1007) with the option plumbing of FW::init (App.cpp:1011-1083): for every
// kernel x ray type x camera: setParams -> beginFrame -> getTotalNumRays -> while nextBatch(): traceBatch once, then
// warmupRepeats untimed and measureRepeats timed repeats; per (kernel, ray type) the #SUM_RENDER_TIME / #SUM_RENDER_KRAYS
// records that the reference's tests/table-tests.sh greps are appended to App.stats (the reference's "Mrays" variable holds
// Krays/s, App.cpp:969-975: the record keeps that unit, the printed table shows true Mrays/s).
//
//   ntrace_bench config.conf -DBenchmark.scene=scene.obj -DBenchmark.camera=<signature>[;<signature>...]
//                -DRenderer.rayType=primary;AO;diffuse -DRenderer.builder=HLBVH -DBenchmark.kernel=<name>[;<name>...]
//
// Keys beyond the reference's: Benchmark.dumpPrefix=<path> writes, for the last camera of every (kernel, ray type), the rays
// and results of the LAST batch as raw arrays (<path>.<kernel>.<type>.rays / .results) so a test can check them;
// Benchmark.device=<ordinal>.
#include "ntrace/NTrace.hpp"
#include <algorithm>
#include <cstdlib>
#include <unistd.h>

using namespace FW;

static std::vector<std::string> splitList(const std::string& s, const char* seps)
{
    std::vector<std::string> out;
    std::string cur;
    for (size_t i = 0; i <= s.size(); i++) {
        bool sep = (i == s.size()) || strchr(seps, s[i]) != NULL;
        if (sep) { if (!cur.empty()) out.push_back(cur); cur.clear(); }
        else cur += s[i];
    }
    return out;
}

static const char* namedSignature(const std::string& name)
{
    // src/rt/App.cpp:51-58, config.conf:11
    static const char* table[][2] = {
        {"conference", "6omr/04j3200bR6Z/0/3ZEAz/x4smy19///c/05frY109Qx7w////m100"},
        {"fairyforest", "cIxMx/sK/Ty/EFu3z/5m9mWx/YPA5z/8///m007toC10AnAHx///Uy200"},
        {"sibenik", "ytIa02G35kz1i:ZZ/0//iSay/5W6Ex19///c/05frY109Qx7w////m100"},
        {"sanmiguel", "Yciwz1oRQmz/Xvsm005CwjHx/b70nx18tVI7005frY108Y/:x/v3/z100"}};
    for (size_t i = 0; i < sizeof(table) / sizeof(table[0]); i++)
        if (name == table[i][0]) return table[i][1];
    return NULL;
}

static void dumpBuffer(const std::string& path, Buffer& b, S64 bytes)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) fail("Cannot write '%s'", path.c_str());
    fwrite(b.getPtr(), 1, (size_t)bytes, f);
    fclose(f);
}

static void runBenchmark(Environment& env)
{
    int device = 0, w = 1024, h = 768, warmupRepeats = 1, measureRepeats = 5, samples = 8;
    float aoRadius = 5.0f;
    bool sortRays = true;
    std::string sceneFile, cameraSpec, kernelSpec = "b200_persistent_speculative_while_while", rayTypeSpec = "primary", ds, statsFile = "stats.log", dumpPrefix;
    env.GetIntValue("Benchmark.device", device);
    env.GetIntValue("App.frameWidth", w); env.GetIntValue("App.frameHeight", h);
    env.GetIntValue("Benchmark.warmupRepeats", warmupRepeats); env.GetIntValue("Benchmark.measureRepeats", measureRepeats);
    env.GetIntValue("Renderer.samples", samples); env.GetFloatValue("Raygen.aoRadius", aoRadius); env.GetBoolValue("Renderer.sortRays", sortRays);
    env.GetStringValue("Benchmark.kernel", kernelSpec); env.GetStringValue("Renderer.rayType", rayTypeSpec);
    env.GetStringValue("App.stats", statsFile); env.GetStringValue("Benchmark.dumpPrefix", dumpPrefix);
    bool pipelined = false;
    env.GetBoolValue("Benchmark.pipelined", pipelined);
    bool frameLaunch = false;                                                // NEW knob: all batches of a frame in one persistent launch
    env.GetBoolValue("Benchmark.frameLaunch", frameLaunch);
    if (env.GetStringValue("Renderer.dataStructure", ds) && ds != "BVH") fail("Incorrect data structure type!  (only Renderer.dataStructure=BVH is on this path)");
    if (!env.GetStringValue("Benchmark.scene", sceneFile) || sceneFile.empty()) fail("Benchmark.scene is not set");
    if (!env.GetStringValue("Benchmark.camera", cameraSpec) || cameraSpec.empty()) fail("Benchmark.camera is empty");

    // Renderer.numGpus (new): one process per GPU.  Rank and world size come from the launcher's environment (RANK / WORLD_SIZE /
    // LOCAL_RANK, as torchrun and most MPI wrappers export them) or from -DBenchmark.rank / -DRenderer.numGpus; the NCCL unique id
    // travels through the file Benchmark.commFile (rank 0 writes it, the others wait for it)
    int numGpus = 1, rank = 0;
    env.GetIntValue("Renderer.numGpus", numGpus);
    if (getenv("WORLD_SIZE")) numGpus = atoi(getenv("WORLD_SIZE"));
    if (getenv("RANK")) rank = atoi(getenv("RANK"));
    env.GetIntValue("Benchmark.rank", rank);
    if (numGpus > 1) { device = getenv("LOCAL_RANK") ? atoi(getenv("LOCAL_RANK")) : rank; }
    ntCheck(nt_init(device));
    if (numGpus > 1) {
        std::string commFile;
        if (!env.GetStringValue("Benchmark.commFile", commFile) || commFile.empty()) fail("Renderer.numGpus > 1 needs Benchmark.commFile (a path every rank can reach)");
        char id[128];
        if (rank == 0) {
            ntCheck(nt_comm_unique_id(id));
            const std::string tmp = commFile + ".tmp";
            FILE* f = fopen(tmp.c_str(), "wb");
            if (!f || fwrite(id, 1, 128, f) != 128) fail("Cannot write '%s'", tmp.c_str());
            fclose(f);
            if (rename(tmp.c_str(), commFile.c_str()) != 0) fail("Cannot publish '%s'", commFile.c_str());
        } else {
            bool got = false;
            for (int i = 0; i < 1200 && !got; i++) {
                FILE* f = fopen(commFile.c_str(), "rb");
                if (f) { got = fread(id, 1, 128, f) == 128; fclose(f); }
                if (!got) usleep(100000);
            }
            if (!got) fail("rank %d: no NCCL unique id at '%s'", rank, commFile.c_str());
        }
        ntCheck(nt_comm_init(numGpus, rank, id));
    }
    std::vector<CameraControls> cameras;
    std::vector<std::string> camTokens = splitList(cameraSpec, ";");
    for (size_t i = 0; i < camTokens.size(); i++) {
        std::string c = camTokens[i];
        while (!c.empty() && isspace((unsigned char)c[0])) c.erase(0, 1);
        while (!c.empty() && isspace((unsigned char)c[c.size() - 1])) c.erase(c.size() - 1);
        if (c.empty()) continue;
        CameraControls cam;
        cam.decodeSignature(namedSignature(c) ? namedSignature(c) : c);
        cameras.push_back(cam);
    }
    if (cameras.empty()) fail("Benchmark.camera is empty");
    std::vector<std::string> kernels = splitList(kernelSpec, ";");
    std::vector<std::string> rayTypes = splitList(rayTypeSpec, "; ");
    std::vector<Renderer::RayType> rayTypeIds;
    for (size_t i = 0; i < rayTypes.size(); i++) {
        std::string r = rayTypes[i];
        std::transform(r.begin(), r.end(), r.begin(), ::tolower);
        if (r == "primary") rayTypeIds.push_back(Renderer::RayType_Primary);
        else if (r == "ao") rayTypeIds.push_back(Renderer::RayType_AO);
        else if (r == "diffuse") rayTypeIds.push_back(Renderer::RayType_Diffuse);
        else fail("Unsupported ray type %s", rayTypes[i].c_str());
    }

    printf("Running benchmark for \"%s\".\n\n", sceneFile.c_str());
    std::unique_ptr<Scene> scene(Scene::importMesh(sceneFile));
    Renderer renderer;
    renderer.setScene(scene.get());
    if (numGpus > 1) renderer.setMultiGpu(rank, numGpus);
    std::string cacheFile;
    if (env.GetStringValue("Benchmark.cacheFile", cacheFile)) renderer.setCacheFile(cacheFile);
    std::string cachePath;
    if (env.GetStringValue("Benchmark.cachePath", cachePath) && !cachePath.empty()) renderer.setCachePath(cachePath);     // files named <hash>_<builder>.dat
    int hlbvhBits = 4, leafSize = 8;
    env.GetIntValue("HLBVH.bits", hlbvhBits); env.GetIntValue("HLBVH.leafSize", leafSize);
    renderer.setHLBVHParams(HLBVHParams(true, hlbvhBits, leafSize, 0.001f));
    bool collapse = false;
    env.GetBoolValue("HLBVH.collapse", collapse);
    ntCheck(nt_bvh_set_collapse(collapse ? 1 : 0, leafSize));
    bool coherent = false;                                                    // NEW knob: slot order of the secondary rays (nt_raygen_set_order)
    env.GetBoolValue("Raygen.coherentOrder", coherent);
    ntCheck(nt_raygen_set_order(coherent ? 1 : 0));

    FILE* stats = fopen(statsFile.c_str(), "a");
    if (!stats) fail("Cannot open stats file '%s'", statsFile.c_str());
    std::vector<double> results;
    for (size_t k = 0; k < kernels.size(); k++)
        for (size_t r = 0; r < rayTypeIds.size(); r++) {
            long long totalRays = 0;
            double totalTime = 0.0;
            for (size_t c = 0; c < cameras.size(); c++) {
                printf("%s, %s, camera %d...\n", kernels[k].c_str(), rayTypes[r].c_str(), (int)c);
                Renderer::Params params;
                params.kernelName = kernels[k]; params.rayType = rayTypeIds[r]; params.numSamples = samples;
                params.aoRadius = aoRadius; params.sortSecondary = sortRays;
                renderer.setParams(params);
                renderer.setPipelined(pipelined);
                RayBuffer* last = NULL;
                if (frameLaunch) {
                    // NEW knob Benchmark.frameLaunch: every batch of the frame is generated into its own buffer (Renderer::prepareFrame), then
                    // ONE persistent launch traces them (traceFrame = nt_trace_batches); the kernel seconds of that launch are what is summed
                    renderer.setPipelined(false);
                    renderer.beginFrame(cameras[c], w, h);
                    totalRays += (long long)renderer.getTotalNumRays() * measureRepeats;
                    renderer.prepareFrame();
                    for (int i = 0; i < 1 + warmupRepeats; i++) renderer.traceFrame();
                    for (int i = 0; i < measureRepeats; i++) totalTime += renderer.traceFrame();
                    if (!renderer.getFrameBatches().empty()) last = renderer.getFrameBatches().back();
                } else if (pipelined) {
                    // NEW knob Benchmark.pipelined: whole frames are repeated instead of single batches; the batches of a frame are queued
                    // back to back on alternating buffers and the device time of the whole batch loop is what is summed
                    for (int rep = 0; rep < warmupRepeats + measureRepeats; rep++) {
                        renderer.beginFrame(cameras[c], w, h);
                        if (rep == 0) totalRays += (long long)renderer.getTotalNumRays() * measureRepeats;
                        renderer.beginTiming();
                        while (renderer.nextBatch()) { renderer.traceBatch(); last = renderer.getBatchRays(); }
                        const double sec = renderer.endTiming();
                        if (rep >= warmupRepeats) totalTime += sec;
                    }
                } else {
                    renderer.beginFrame(cameras[c], w, h);
                    totalRays += (long long)renderer.getTotalNumRays() * measureRepeats;
                    while (renderer.nextBatch()) {
                        renderer.traceBatch();
                        for (int i = 0; i < warmupRepeats; i++) renderer.traceBatch();
                        for (int i = 0; i < measureRepeats; i++) totalTime += renderer.traceBatch();
                        last = renderer.getBatchRays();
                    }
                }
                if (!dumpPrefix.empty() && c + 1 == cameras.size() && last) {
                    std::string base = dumpPrefix + "." + kernels[k] + "." + rayTypes[r];
                    dumpBuffer(base + ".rays", last->getRayBuffer(), (S64)last->getSize() * 32);
                    dumpBuffer(base + ".results", last->getResultBuffer(), (S64)last->getSize() * 16);
                }
            }
            if (numGpus > 1) ntCheck(nt_comm_allreduce(&totalTime, 1, 1));                           // the slowest rank's time; the rays are the frame's
            double krays = totalTime > 0.0 ? (double)totalRays / totalTime * 1.0e-3 : 0.0;
            results.push_back(krays);
            if (rank == 0) fprintf(stats, "#SUM_RENDER_TIME\n%g\n#SUM_RENDER_KRAYS\n%g\n", totalTime, krays);      // pushStat (Defs.hpp:166-172)
        }
    fclose(stats);
    if (numGpus > 1) {
        if (rank == 0) printf("%d GPUs: BVH broadcast %.3f ms\n", numGpus, renderer.getBroadcastTime() * 1e3);
        ntCheck(nt_comm_destroy());
        if (rank != 0) return;
    }

    printf("Done.\n\n%-42s", "Kernel");
    for (size_t r = 0; r < rayTypes.size(); r++) printf("%-14s", rayTypes[r].c_str());
    printf("  [Mrays/s]\n%-42s", "---");
    for (size_t r = 0; r < rayTypes.size(); r++) printf("%-14s", "---");
    printf("\n");
    for (size_t k = 0; k < kernels.size(); k++) {
        printf("%-42s", kernels[k].c_str());
        for (size_t r = 0; r < rayTypes.size(); r++) printf("%-14.2f", results[k * rayTypes.size() + r] * 1.0e-3);
        printf("\n");
    }
    printf("%-42s", "---");
    for (size_t r = 0; r < rayTypes.size(); r++) printf("%-14s", "---");
    printf("\n\n");
    nt_shutdown();
}

int main(int argc, char** argv)
{
    try {
        Environment env;
        env.Parse(argc, argv);
        Environment::SetSingleton(&env);
        bool benchmark = true;
        env.GetBoolValue("App.benchmark", benchmark);
        if (!benchmark) fail("only App.benchmark=true is supported (the interactive GUI is out of scope)");
        runBenchmark(env);
    } catch (const std::exception& e) {
        fprintf(stderr, "ntrace_bench: %s\n", e.what());
        return 1;
    }
    return 0;
}
