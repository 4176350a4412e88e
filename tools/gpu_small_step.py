#!/usr/bin/env python
"""Per-step device times of encode + distance kernel on a small alignment (is a small launch slow?).
    python tools/gpu_small_step.py [n] [L]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from btb_phylo_b200 import _capi, build, device as dev  # noqa: E402

build.build()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6400
L = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
sys.path.insert(0, ROOT)
import bench  # noqa: E402

seqs = bench.make_alignment_cuda(n, L, torch.device("cuda", 0))
plan = dev.DistancePlan(seqs, length=L, engine=_capi.ENGINE_UMMA)
out = plan.alloc_matrix()
T = dev.block_rows(n)
for it in range(12):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t0 = time.perf_counter()
    e[0].record(); plan.encode(); e[1].record(); plan.compute_rows(out, 0, T); e[2].record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("iter %2d  encode %.3f ms  kernel %.3f ms  host enqueue %.3f ms" % (it, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), (t1 - t0) * 1e3))
