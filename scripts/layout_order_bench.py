"""Experiment: does the ORDER of the 64-byte nodes in memory matter to the traversal kernel?  Builds the bench's BVH on the GPU,
renumbers its inner nodes on the host in several orders (links rewritten, tree unchanged), uploads each and times the trace.
Usage: python scripts/layout_order_bench.py [scene=conference] -> gpurun_out/layout_order_<scene>.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import camera, capi, host, scenes  # noqa: E402


def reorder(nodes, order_kind, rng):
    n = nodes.reshape(-1, 16)
    N = len(n)
    links = n[:, 12:14]
    child = np.where(links >= 0, links // 64, -1)
    new_of_old = np.full(N, -1, np.int64)
    if order_kind == "original":
        return nodes.copy()
    if order_kind == "random":
        perm = np.concatenate([[0], 1 + rng.permutation(N - 1)])
        new_of_old[perm] = np.arange(N)
    elif order_kind == "preorder":
        k, stack = 0, [0]
        while stack:
            x = stack.pop()
            new_of_old[x] = k; k += 1
            c0, c1 = child[x]
            if c1 >= 0: stack.append(c1)
            if c0 >= 0: stack.append(c0)
    elif order_kind == "siblings":                      # children of a node adjacent (what createCompact does), depth-first
        k, stack = 1, [0]
        new_of_old[0] = 0
        while stack:
            x = stack.pop()
            c0, c1 = child[x]
            if c0 >= 0: new_of_old[c0] = k; k += 1
            if c1 >= 0: new_of_old[c1] = k; k += 1
            if c1 >= 0: stack.append(c1)
            if c0 >= 0: stack.append(c0)
    elif order_kind == "bfs":
        k, q = 0, [0]
        while q:
            nq = []
            for x in q:
                new_of_old[x] = k; k += 1
                for c in child[x]:
                    if c >= 0: nq.append(c)
            q = nq
    assert (new_of_old >= 0).all()
    out = np.zeros_like(n)
    out[new_of_old] = n
    l = out[:, 12:14]
    out[:, 12:14] = np.where(l >= 0, new_of_old[np.maximum(l // 64, 0)] * 64, l)
    return out.reshape(-1)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "conference"
    host.init(0)
    verts, tris, cam_name = scenes.config_scene(name)
    cam = camera.named_camera(cam_name)
    scene = host.Scene(verts, tris)
    lo, hi = scene.getBBox()
    capi.bvh_set_collapse(1, 8)
    capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, lo, hi, 2, 8, 0.001)
    capi.bvh_set_collapse(0, 0)
    nodes, woop, idx, _ = capi.bvh_download()
    tracer = host.CudaBVHTracer()
    bvh = host.CudaBVH(layout=4); bvh.resident = True
    tracer.setBVH(bvh)
    prim = host.RayBuffer()
    host.RayGen().primary(prim, cam.position, camera.nscreen_to_world(cam, 1024, 768), 1024, 768, cam.far)
    tracer.traceBatch(prim)
    ao, diff = host.RayBuffer(), host.RayBuffer()
    g1, g2 = host.RayGen(1 << 20), host.RayGen(1 << 20)
    new = True
    for _ in range(10):
        ok, new = g1.ao(ao, prim, scene, 32, 5.0, new, host.FIXED_AO_SEED)
    new = True
    for _ in range(10):
        ok, new = g2.ao(diff, prim, scene, 32, cam.far, new, host.FIXED_AO_SEED)
    diff.setNeedClosestHit(True)
    rng = np.random.default_rng(0)
    out = {"scene": name, "rows": []}
    base = None
    for kind in ("original", "preorder", "siblings", "bfs", "random"):
        capi.bvh_upload(4, reorder(nodes, kind, rng), woop, idx)
        tracer.setBVH(host.CudaBVH(layout=4).__class__(layout=4)) if False else None
        b = host.CudaBVH(layout=4); b.resident = True
        tracer.setBVH(b)
        row = {"order": kind}
        for rt, rb in (("primary", prim), ("AO", ao), ("diffuse", diff)):
            tracer.traceBatch(rb)
            sec = min(tracer.traceBatch(rb) for _ in range(7))
            row[rt] = rb.getSize() / sec * 1e-6
        if kind == "original":
            base = prim.results_host()[:, :2].copy()
        else:
            tracer.traceBatch(prim)
            assert np.array_equal(prim.results_host()[:, :2], base)
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/layout_order_{name}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
