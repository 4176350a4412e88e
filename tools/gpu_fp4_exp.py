#!/usr/bin/env python
"""Experiments on the FP4 distance kernels: (1) does the negate-A bit of the block-scaled instruction
descriptor work for kind::mxf4?  (2) lockstep slack sweep at C4."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from btb_phylo_b200 import _capi, device, synth
import bench

lib = _capi.lib()


def matrix(d, L, n):
    plan = device.DistancePlan(d, length=L, engine=1)
    plan.encode()
    out = plan.alloc_matrix()
    plan.compute(out)
    torch.cuda.synchronize()
    return out[:, :n].cpu().numpy().astype(np.int64)


for n, L in ((600, 1000), (40000, 1000)):
    seqs = synth.snp_alignment(n, L, seed=4, heavy_n=0.0)
    seqs = np.where(np.isin(seqs, list(b"ACGT")), seqs, ord("A")).astype(np.uint8)
    stride = (L + 15) // 16 * 16
    host = np.full((n, stride), ord("."), dtype=np.uint8); host[:, :L] = seqs
    d = torch.from_numpy(host).cuda()
    normal = matrix(d, L, n)
    _capi.check(lib.btb_set_option(b"umma_dbg_negate", 1))
    neg = matrix(d, L, n)
    _capi.check(lib.btb_set_option(b"umma_dbg_negate", 0))
    verdict = "no effect" if (neg == normal).all() else "changed"
    for sym in b"ACGT":
        c = (seqs != sym).sum(axis=1).astype(np.int64)
        if ((neg + normal) == 2 * (c[:, None] + c[None, :]))[np.triu_indices(n, 1)].all():
            verdict = "negate works (zero-coded symbol %s)" % chr(sym)
    print("negate probe n=%d L=%d: %s; sample normal %s neg %s" % (n, L, verdict, normal[0, 1:4], neg[0, 1:4]), flush=True)

n, L = 50000, 30000
d = bench.make_alignment_cuda(n, L, torch.device("cuda", 0))
plan = device.DistancePlan(d, length=L, engine=1)
plan.encode()
out = plan.alloc_matrix()
ref = None
for key, values in ((b"umma_tile_bk", (128, 64, 128, 64)),):
    for v in values:
        _capi.check(lib.btb_set_option(key, v))
        plan.compute(out); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.compute(out); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        chk = int(out[::997, ::991].to(torch.int64).sum().item())
        ref = chk if ref is None else ref
        print("%s=%d ms=%s same=%s" % (key.decode(), v, ["%.2f" % t for t in ts], chk == ref), flush=True)
