"""Two GPU builds of a 10 M-triangle soup for ncu (builder kernels in their HBM-bound regime):
python scripts/profile_build.py <lbvh|hlbvh>   -- capture the second build (`ncu --set full -k regex:'morton|radix|topology|finalize|emit|hlbvh|cluster' ...`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import capi, host, scenes  # noqa: E402

import torch
host.init(0)
v, t = scenes.soup_uniform(10_000_000, 5)
lo, hi = scenes.bbox(v)
dv = torch.from_numpy(v).cuda(); dt = torch.from_numpy(t).cuda()
torch.cuda.synchronize()
hl = sys.argv[1] == "hlbvh"
for _ in range(2):
    s = capi.bvh_build(1 if hl else 0, dv, dt, lo, hi, 4 if hl else 10, 8, 0.001)
print(f"{sys.argv[1]} 10M: {s * 1e3:.2f} ms")
