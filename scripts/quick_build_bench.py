"""Early measurement helper: GPU LBVH / HLBVH build time and Mtris/s for a few sizes. Usage: python scripts/quick_build_bench.py [--big]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import capi, host, scenes  # noqa: E402


def main():
    import torch
    host.init(0)
    cases = [("room283k", lambda: scenes.room(283_000, 2)), ("soup1M", lambda: scenes.soup_uniform(1_000_000, 5)),
             ("soup10M", lambda: scenes.soup_uniform(10_000_000, 5)), ("room10.5M", lambda: scenes.room(10_500_000, 4, wall_frac=0.2))]
    if "--big" in sys.argv:
        cases.append(("soup50M", lambda: scenes.soup_uniform(50_000_000, 5)))
    for name, gen in cases:
        v, t = gen()
        lo, hi = scenes.bbox(v)
        dv = torch.from_numpy(v).cuda(); dt = torch.from_numpy(t).cuda()
        torch.cuda.synchronize()
        for label, builder, bits, leaf in (("LBVH leaf8", 0, 10, 8), ("LBVH leaf1", 0, 10, 1), ("HLBVH bits4 leaf8", 1, 4, 8)):
            if leaf == 1 and len(t) > 30_000_000:
                continue                      # > 31 M inner nodes: needs the Compact2 build layout (scripts/rebuild_sweep.py covers it)
            ts = [capi.bvh_build(builder, dv, dt, lo, hi, bits, leaf, 0.001) for _ in range(5)]
            (nb, wb, ib), _ = capi.bvh_sizes()
            print(f"{name} {label}: {np.min(ts[1:]) * 1e3:.3f} ms best ({len(t) / np.min(ts[1:]) * 1e-6:.1f} Mtris/s), first {ts[0] * 1e3:.2f} ms, nodes {nb // 64}", flush=True)


if __name__ == "__main__":
    main()
