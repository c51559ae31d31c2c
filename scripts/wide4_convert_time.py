import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from ntrace_b200 import capi, host, scenes
host.init(0)
for name, gen in [("room283k", lambda: scenes.room(283_000, 2)), ("room10.5M", lambda: scenes.room(10_500_000, 4, wall_frac=0.2))]:
    v, t = gen(); lo, hi = scenes.bbox(v)
    dv = torch.from_numpy(v).cuda(); dt = torch.from_numpy(t).cuda()
    capi.bvh_set_collapse(1, 8)
    capi.bvh_build(1, dv, dt, lo, hi, 2, 8, 0.001)
    capi.set_kernel("b200_wide4")
    for rep in range(3):
        capi.bvh_build(1, dv, dt, lo, hi, 2, 8, 0.001)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        size = capi.C.c_size_t(0); depth = capi.C.c_int(0)
        capi._check(capi.lib().nt_bvh_wide4_download(None, capi.C.c_size_t(0), capi.C.byref(size), capi.C.byref(depth)))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(name, "device wide4 conversion (host wall incl. sync) ms", (t1 - t0) * 1e3, "wide nodes", size.value // 64, "depth", depth.value, flush=True)
    if "283k" in name:
        nodes, woop, idx, lay = capi.bvh_download()
        t0 = time.perf_counter(); capi.bvh_wide4_convert_host(lay, nodes, woop.nbytes); t1 = time.perf_counter()
        print(name, "host conversion ms", (t1 - t0) * 1e3 / 2)
