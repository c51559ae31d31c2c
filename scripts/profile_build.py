"""Two GPU builds of a triangle soup for ncu (10 M by default: builder kernels in their HBM-bound regime):
python scripts/profile_build.py <lbvh|hlbvh> [numTris | room:<numTris>]
  -- capture the second build (`ncu --set full -k regex:'morton|radix|topology|finalize|emit|hlbvh|cluster' ...`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import capi, host, scenes  # noqa: E402

import torch
host.init(0)
what = sys.argv[2] if len(sys.argv) > 2 else "10000000"
if what.startswith("room:"):
    v, t = scenes.room(int(what[5:]), 2)
else:
    v, t = scenes.soup_uniform(int(what), 5)
lo, hi = scenes.bbox(v)
dv = torch.from_numpy(v).cuda(); dt = torch.from_numpy(t).cuda()
torch.cuda.synchronize()
hl = sys.argv[1] == "hlbvh"
for _ in range(2):
    s = capi.bvh_build(1 if hl else 0, dv, dt, lo, hi, 4 if hl else 10, 8, 0.001)
print(f"{sys.argv[1]} {what}: {s * 1e3:.3f} ms")
