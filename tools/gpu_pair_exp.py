#!/usr/bin/env python
"""cta_group::2 FP4 kernel: correctness against the single-CTA kernel, then C4 timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from btb_phylo_b200 import _capi, device, synth
import bench
lib = _capi.lib()

def run(d, L, pair, reps=0):
    _capi.check(lib.btb_set_option(b"umma_fp4_pair", pair))
    plan = device.DistancePlan(d, length=L, engine=1)
    plan.encode()
    out = plan.alloc_matrix()
    plan.compute(out); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.compute(out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return out, ts

for n, L, heavy in ((6400, 777, 0.0), (6657, 1500, 0.0), (7000, 900, 0.1)):
    seqs = synth.snp_alignment(n, L, seed=14, heavy_n=heavy)
    stride = (L + 15) // 16 * 16
    host = np.full((n, stride), ord("."), dtype=np.uint8); host[:, :L] = seqs
    d = torch.from_numpy(host).cuda()
    a, _ = run(d, L, 0)
    b, _ = run(d, L, 1)
    bad = (a[:, :n] != b[:, :n])
    print("n=%d L=%d heavy=%.1f mismatches=%d" % (n, L, heavy, int(bad.sum().item())), flush=True)
    if bad.any():
        idx = bad.nonzero()
        print("  first", idx[0].tolist(), "last", idx[-1].tolist(), "got", int(b[tuple(idx[0].tolist())]), "want", int(a[tuple(idx[0].tolist())]), flush=True)

n, L = 50000, 30000
d = bench.make_alignment_cuda(n, L, torch.device("cuda", 0))
ref = None
for pair in (0, 1, 0, 1):
    out, ts = run(d, L, pair, reps=3)
    chk = int(out[::997, ::991].to(torch.int64).sum().item())
    ref = chk if ref is None else ref
    print("C4 pair=%d ms=%s same=%s" % (pair, ["%.2f" % t for t in ts], chk == ref), flush=True)
    del out
