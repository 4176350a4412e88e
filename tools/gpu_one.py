#!/usr/bin/env python
"""One distance launch for profiling: python tools/gpu_one.py <n> <L> <engine popc|umma> <cg> [heavy_n] [simplex]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from btb_phylo_b200 import _capi, device, synth
n, L, eng, cg = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
heavy = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
simplex = int(sys.argv[6]) if len(sys.argv) > 6 else 0
variant = int(sys.argv[7]) if len(sys.argv) > 7 else 0
lib = _capi.lib()
_capi.check(lib.btb_set_option(b"umma_cta_group", cg))
_capi.check(lib.btb_set_option(b"umma_dense_simplex", simplex))
_capi.check(lib.btb_set_option(b"umma_cg2_variant", variant))
seqs = synth.snp_alignment(n, L, seed=4, heavy_n=heavy)
stride = (L + 15) // 16 * 16
host = np.full((n, stride), ord("."), dtype=np.uint8); host[:, :L] = seqs
d = torch.from_numpy(host).cuda()
plan = device.DistancePlan(d, length=L, engine={"popc": 0, "umma": 1}[eng])
plan.encode()
out = plan.alloc_matrix()
for _ in range(3):
    plan.compute(out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); plan.compute(out); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
ok = ""
if os.environ.get("BTB_CHECK"):
    ref = device.DistancePlan(d, length=L, engine=0); ref.encode(); o2 = ref.alloc_matrix(); ref.compute(o2); torch.cuda.synchronize()
    ok = " match_popc=%s" % bool((out == o2).all().item())
print("n=%d L=%d engine=%s cg=%d sigma=%d simplex=%d variant=%d%s ms=%.3f pairs/s=%.3e tops_sigma4=%.0f" % (n, L, eng, cg, plan.sigma, simplex, variant, ok, ms, n*(n-1)/2/(ms*1e-3), 8*L*n*(n-1)/2/(ms*1e-3)/1e12))
