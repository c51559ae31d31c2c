/* Plain-C consumer of the drop-in boundary: compiles include/ntrace_b200.h as C (no C++/torch types in the ABI), links
 * libntrace_b200.so, builds a BVH for a small quad grid on the GPU and traces host-resident rays through it.
 * Exit code 0 = all checks passed; prints the reason otherwise.  Built and run by tests/test_cabi.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ntrace_b200.h"

#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "FAIL %s: %s\n", #call, nt_last_error()); return 2; } } while (0)

int main(void)
{
    enum { G = 16, NV = (G + 1) * (G + 1), NT = 2 * G * G, NR = 64 };
    static float verts[NV * 3];
    static int32_t tris[NT * 3];
    static float rays[NR * 8];
    static int32_t results[NR * 4];
    int i, j, t = 0, hits = 0, errors = 0;
    float lo[3] = {0.0f, 0.0f, 0.0f}, hi[3] = {1.0f, 1.0f, 0.0f}, sec = 0.0f;
    size_t sizes[3];
    int layout = -1;
    int32_t cfg[4];

    /* error path first: tracing before init must fail loudly, not fall back */
    if (nt_trace_batch(rays, results, NR, 1, &sec) == 0) { fprintf(stderr, "FAIL: trace before nt_init succeeded\n"); return 2; }
    if (strstr(nt_last_error(), "nt_init") == NULL) { fprintf(stderr, "FAIL: unexpected error text: %s\n", nt_last_error()); return 2; }

    CHECK(nt_init(0));
    for (j = 0; j <= G; j++)
        for (i = 0; i <= G; i++) {
            float* v = verts + 3 * (j * (G + 1) + i);
            v[0] = (float)i / G; v[1] = (float)j / G; v[2] = 0.0f;
        }
    for (j = 0; j < G; j++)
        for (i = 0; i < G; i++) {
            int a = j * (G + 1) + i, b = a + 1, c = a + G + 1, d = c + 1;
            tris[t++] = a; tris[t++] = b; tris[t++] = d;
            tris[t++] = a; tris[t++] = d; tris[t++] = c;
        }
    CHECK(nt_set_kernel("b200_persistent_speculative_while_while"));
    CHECK(nt_kernel_config(cfg));
    if (cfg[0] != NT_LAYOUT_COMPACT || nt_desired_layout() != NT_LAYOUT_COMPACT || cfg[3] != 1) { fprintf(stderr, "FAIL: kernel config\n"); return 2; }
    /* no BVH yet: the reference's error string */
    if (nt_trace_batch(rays, results, NR, 1, &sec) == 0 || strstr(nt_last_error(), "No BVH") == NULL) { fprintf(stderr, "FAIL: missing-BVH error\n"); return 2; }
    CHECK(nt_bvh_build(NT_BUILDER_LBVH, verts, NV, tris, NT, lo, hi, 10, 4, 0.001f, &sec));
    CHECK(nt_bvh_sizes(sizes, &layout));
    if (layout != NT_LAYOUT_COMPACT || sizes[0] % 64 != 0 || sizes[1] != 4 * sizes[2]) { fprintf(stderr, "FAIL: buffer sizes\n"); return 2; }

    /* rays straight down at cell centres (hit) and outside the grid (miss) */
    for (i = 0; i < NR; i++) {
        float* r = rays + 8 * i;
        float x = (i % 8 + 0.3f) / 8.0f + ((i >= 56) ? 5.0f : 0.0f), y = (i / 8 + 0.6f) / 8.0f;
        r[0] = x; r[1] = y; r[2] = 2.0f; r[3] = 0.0f;
        r[4] = 0.0f; r[5] = 0.0f; r[6] = -1.0f; r[7] = 10.0f;
    }
    CHECK(nt_trace_batch(rays, results, NR, 1, &sec));
    for (i = 0; i < NR; i++) {
        int32_t id = results[4 * i];
        float tt; memcpy(&tt, &results[4 * i + 1], 4);
        if (i < 56) {
            float x = rays[8 * i], y = rays[8 * i + 1];
            int cx = (int)(x * G), cy = (int)(y * G);
            int expect0 = 2 * (cy * G + cx), ok = (id == expect0 || id == expect0 + 1);
            hits++;
            if (!ok || fabsf(tt - 2.0f) > 1e-5f) { errors++; fprintf(stderr, "ray %d: id %d (expected %d/%d) t %f\n", i, id, expect0, expect0 + 1, tt); }
        } else if (id != -1 || tt != 10.0f) { errors++; fprintf(stderr, "ray %d should miss: id %d t %f\n", i, id, tt); }
    }
    CHECK(nt_count_hits(results, NR, &j));
    if (j != hits) { fprintf(stderr, "FAIL: count_hits %d != %d\n", j, hits); return 2; }
    if (nt_launch_count() <= 0) { fprintf(stderr, "FAIL: no kernel launches counted\n"); return 2; }
    nt_shutdown();
    if (errors) return 1;
    printf("cabi_smoke OK: %d rays, %d hits, kernel %.3f ms\n", NR, hits, sec * 1e3f);
    return 0;
}
