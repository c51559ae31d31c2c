"""dev tool: build time of the bench tree (HLBVH(2) + collapse) and HLBVH(4) at 283 K / 10.5 M for the library named by NTRACE_B200_LIB"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import capi, host, scenes
import torch
host.init(0)
out = []
for name, gen in [("room283k", lambda: scenes.room(283_000, 2)), ("room10.5M", lambda: scenes.room(10_500_000, 4, wall_frac=0.2))]:
    v, t = gen(); lo, hi = scenes.bbox(v)
    dv = torch.from_numpy(v).cuda(); dt = torch.from_numpy(t).cuda()
    for label, bits, col in (("H2c", 2, 1), ("H4", 4, 0)):
        capi.bvh_set_collapse(col, 8)
        ts = [capi.bvh_build(1, dv, dt, lo, hi, bits, 8, 0.001) for _ in range(8)]
        out.append(f"{name}/{label} {np.min(ts[1:]) * 1e3:.3f}")
print(os.environ.get("NTRACE_B200_LIB", "default"), os.environ.get("NT_TOP_CTAS", "-"), " ".join(out))
