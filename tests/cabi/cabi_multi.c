/* Plain-C, one-process-per-GPU consumer of the multi-GPU part of the boundary (no Python, no torch, no MPI in the process):
 *     cabi_multi <rank> <numRanks> <idFile>
 * Rank 0 asks the library for an NCCL unique id and publishes it through <idFile> (written under a temporary name, then renamed);
 * the other ranks poll for the file.  Every rank selects GPU <rank>, joins the communicator, rank 0 builds a BVH on its GPU,
 * nt_bvh_broadcast replicates it, every rank traces ITS contiguous slice of one ray batch and checks it against the
 * analytic answer; hit counts and the slowest rank's kernel time are combined with nt_comm_allreduce.
 * Exit code 0 = all checks passed.  Built and run by tests/test_cabi.py (numRanks 1 on a single GPU, 2 when two are visible). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "ntrace_b200.h"

#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "rank %d FAIL %s: %s\n", rank, #call, nt_last_error()); return 2; } } while (0)

int main(int argc, char** argv)
{
    enum { G = 64, NV = (G + 1) * (G + 1), NT = 2 * G * G, NR = 4096 };
    static float verts[NV * 3];
    static int32_t tris[NT * 3];
    static float rays[NR * 8];
    static int32_t results[NR * 4];
    int rank, nranks, i, j, t = 0, lo_slot, hi_slot, per, layout = -1, errors = 0;
    char id[128], tmpname[1024];
    float lo[3] = {0.0f, 0.0f, 0.0f}, hi[3] = {1.0f, 1.0f, 0.0f}, sec = 0.0f, bsec = 0.0f;
    size_t sizes[3] = {0, 0, 0}, sizes0[3];
    double red[2];
    FILE* f;

    if (argc != 4) { fprintf(stderr, "usage: cabi_multi <rank> <numRanks> <idFile>\n"); return 2; }
    rank = atoi(argv[1]); nranks = atoi(argv[2]);
    CHECK(nt_init(rank));
    /* a broadcast without a communicator must fail loudly */
    if (nt_bvh_broadcast(0, &bsec) == 0 || strstr(nt_last_error(), "communicator") == NULL) { fprintf(stderr, "FAIL: broadcast without communicator\n"); return 2; }
    if (rank == 0) {
        CHECK(nt_comm_unique_id(id));
        snprintf(tmpname, sizeof(tmpname), "%s.tmp", argv[3]);
        f = fopen(tmpname, "wb");
        if (!f || fwrite(id, 1, 128, f) != 128) { fprintf(stderr, "FAIL: cannot write %s\n", tmpname); return 2; }
        fclose(f);
        if (rename(tmpname, argv[3]) != 0) { fprintf(stderr, "FAIL: rename\n"); return 2; }
    } else {
        for (i = 0; i < 600; i++) {                      /* up to 60 s */
            f = fopen(argv[3], "rb");
            if (f) { size_t n = fread(id, 1, 128, f); fclose(f); if (n == 128) break; }
            usleep(100000);
        }
        if (i == 600) { fprintf(stderr, "rank %d FAIL: no unique id file\n", rank); return 2; }
    }
    CHECK(nt_comm_init(nranks, rank, id));

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
    if (rank == 0) {
        CHECK(nt_bvh_build(NT_BUILDER_HLBVH, verts, NV, tris, NT, lo, hi, 4, 4, 0.001f, &sec));
        CHECK(nt_bvh_sizes(sizes0, &layout));
    }
    CHECK(nt_bvh_broadcast(0, &bsec));
    CHECK(nt_bvh_sizes(sizes, &layout));
    if (layout != NT_LAYOUT_COMPACT || sizes[0] == 0 || sizes[0] % 64 != 0 || sizes[1] != 4 * sizes[2]) { fprintf(stderr, "rank %d FAIL: replica sizes\n", rank); return 2; }
    /* every rank must hold the same number of bytes as the builder */
    red[0] = (double)sizes[0] + (double)sizes[1] + (double)sizes[2]; red[1] = -red[0];
    CHECK(nt_comm_allreduce(red, 2, 1));
    if (red[0] != -red[1]) { fprintf(stderr, "rank %d FAIL: replicas differ in size\n", rank); return 2; }

    /* one batch; rank g traces the contiguous slot range [g * ceil(N / G), min(N, (g + 1) * ceil(N / G))) */
    for (i = 0; i < NR; i++) {
        float* r = rays + 8 * i;
        float x = ((i % 64) + 0.37f) / 64.0f + ((i >= NR - 256) ? 3.0f : 0.0f), y = ((i / 64) + 0.61f) / 64.0f;
        r[0] = x; r[1] = y; r[2] = 2.0f; r[3] = 0.0f;
        r[4] = 0.0f; r[5] = 0.0f; r[6] = -1.0f; r[7] = 10.0f;
    }
    per = (NR + nranks - 1) / nranks;
    lo_slot = rank * per; hi_slot = lo_slot + per; if (hi_slot > NR) hi_slot = NR; if (lo_slot > NR) lo_slot = NR;
    CHECK(nt_set_kernel("b200_persistent_speculative_while_while"));
    CHECK(nt_trace_batch(rays + 8 * lo_slot, results + 4 * lo_slot, hi_slot - lo_slot, 1, &sec));
    red[0] = 0.0;
    for (i = lo_slot; i < hi_slot; i++) {
        int32_t idr = results[4 * i];
        float tt; memcpy(&tt, &results[4 * i + 1], 4);
        if (i >= NR - 256) { if (idr != -1) errors++; }
        else { if (idr < 0 || idr >= NT || fabsf(tt - 2.0f) > 1e-5f) errors++; else red[0] += 1.0; }
    }
    if (errors) { fprintf(stderr, "rank %d FAIL: %d wrong results in its slice\n", rank, errors); return 2; }
    red[1] = 0.0;
    CHECK(nt_comm_allreduce(red, 1, 0));                  /* hits over all ranks */
    if (red[0] != (double)(NR - 256)) { fprintf(stderr, "rank %d FAIL: %g hits over all ranks, expected %d\n", rank, red[0], NR - 256); return 2; }
    red[0] = sec;
    CHECK(nt_comm_allreduce(red, 1, 1));                  /* the slowest rank's kernel time */
    if (!(red[0] >= sec && red[0] > 0.0)) { fprintf(stderr, "rank %d FAIL: max reduce\n", rank); return 2; }
    CHECK(nt_comm_destroy());
    nt_shutdown();
    if (rank == 0) printf("cabi_multi OK: %d ranks, BVH %zu bytes broadcast in %.3f ms, %d rays sharded\n", nranks, sizes[0] + sizes[1] + sizes[2], bsec * 1e3, NR);
    return 0;
}
