#!/usr/bin/env python
"""Kernel time against L at fixed n: separates the per-unit overhead from the K loop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from btb_phylo_b200 import _capi, device
import bench
lib = _capi.lib()
n = 50000
for L in (3750, 7500, 15000, 30000):
    d = bench.make_alignment_cuda(n, L, torch.device("cuda", 0))
    for pair in (0, 1):
        _capi.check(lib.btb_set_option(b"umma_fp4_pair", pair))
        plan = device.DistancePlan(d, length=L, engine=1)
        plan.encode()
        out = plan.alloc_matrix()
        plan.compute(out); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.compute(out); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("L=%d pair=%d ms=%s" % (L, pair, ["%.2f" % t for t in ts]), flush=True)
        del out, plan
    del d
