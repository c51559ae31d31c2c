"""dev tool: build time of the two 10 M scenes and the 283 K room for the library named by NTRACE_B200_LIB."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import capi, host, scenes  # noqa: E402

import torch
host.init(0)
out = []
for name, gen in [("room283k", lambda: scenes.room(283_000, 2)), ("soup10M", lambda: scenes.soup_uniform(10_000_000, 5)),
                  ("room10.5M", lambda: scenes.room(10_500_000, 4, wall_frac=0.2))]:
    v, t = gen()
    lo, hi = scenes.bbox(v)
    dv = torch.from_numpy(v).cuda(); dt = torch.from_numpy(t).cuda()
    torch.cuda.synchronize()
    for label, builder, bits in (("L", 0, 10), ("H", 1, 4)):
        ts = [capi.bvh_build(builder, dv, dt, lo, hi, bits, 8, 0.001) for _ in range(7)]
        out.append(f"{name}/{label} {np.min(ts[1:]) * 1e3:.3f}")
print(os.environ.get("NTRACE_B200_LIB", "default"), " ".join(out))
