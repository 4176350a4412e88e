#!/usr/bin/env python
"""How much of the distance kernel's time is the mirrored half of the epilogue?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from btb_phylo_b200 import _capi, device
import bench
n, L = 50000, 30000
d = bench.make_alignment_cuda(n, L, torch.device("cuda", 0))
plan = device.DistancePlan(d, length=L, engine=1)
plan.encode()
out = plan.alloc_matrix()
for mirror in (True, False, True, False):
    plan.compute(out, mirror=mirror); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.compute(out, mirror=mirror); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("mirror=%s ms=%s" % (mirror, ["%.2f" % t for t in ts]), flush=True)
