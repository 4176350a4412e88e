#!/usr/bin/env python
"""BASELINE config 5 at full size on one B200: 100,000 samples (= snp-dists MAX_SEQ) x 60,000 sites,
20-30 % of cells N / '-' (shared blocks + private runs), 0.1 % IUPAC, a few lower-case bytes.
Runs the distance stage in `-a` mode and in default mode, times it, and checks the result by
size-independent properties (symmetry, zero diagonal, engine agreement on a row block) and by the
CPU oracle on a handful of full rows.

    python tools/bench_c5.py [n] [L] [--rows R]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from btb_phylo_b200 import _capi, device as dev


def make_c5(n, L, device, seed=5, row_chunk=2048):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    base_i = torch.randint(0, 4, (L,), generator=g, device=device)
    minor_i = (base_i + torch.randint(1, 4, (L,), generator=g, device=device)) % 4
    base, minor = acgt[base_i], acgt[minor_i]
    u = torch.rand(L, generator=g, device=device, dtype=torch.float64)
    size = torch.clamp((float(n) ** u * 0.5 + 0.5).to(torch.int64), 1, max(1, n - 1))
    a = (torch.rand(L, generator=g, device=device, dtype=torch.float64) * (n - size + 1).to(torch.float64)).to(torch.int64)
    b = a + size
    order = torch.randperm(n, generator=g, device=device)
    # shared masked blocks: 15 % of the columns, each sample carries a block with p = 0.9
    n_blocks = max(1, L // 400)
    block_of = torch.full((L,), -1, dtype=torch.int64, device=device)
    starts = torch.randint(0, max(1, L - 60), (n_blocks,), generator=g, device=device)
    for k in range(60):
        block_of[torch.clamp(starts + k, max=L - 1)] = torch.arange(n_blocks, device=device)
    iupac = torch.tensor(list(b"RYKMSW"), dtype=torch.uint8, device=device)
    stride = (L + 15) // 16 * 16
    out = torch.full((n, stride), ord("."), dtype=torch.uint8, device=device)
    for r0 in range(0, n, row_chunk):
        r1 = min(n, r0 + row_chunk)
        pos = order[r0:r1, None]
        rows = torch.where((pos >= a[None, :]) & (pos < b[None, :]), minor[None, :], base[None, :])
        has_block = torch.rand((r1 - r0, n_blocks), generator=g, device=device) < 0.9
        in_block = (block_of >= 0)[None, :] & torch.gather(has_block, 1, torch.clamp(block_of, min=0)[None, :].expand(r1 - r0, L))
        rows = torch.where(in_block, torch.tensor(ord("N"), dtype=torch.uint8, device=device), rows)
        r = torch.rand((r1 - r0, L), generator=g, device=device)
        rows = torch.where(r < 0.10, torch.tensor(ord("-"), dtype=torch.uint8, device=device), rows)      # private gaps
        rows = torch.where((r >= 0.10) & (r < 0.101), iupac[torch.randint(0, 6, (r1 - r0, L), generator=g, device=device)], rows)
        lower = (r >= 0.101) & (r < 0.1012)
        rows = torch.where(lower, rows | 0x20, rows)
        out[r0:r1, :L] = rows
        del rows, has_block, in_block, r, lower
    return out


def run_mode(seqs, n, L, allchars, oracle_rows, threads):
    res = {}
    t0 = time.perf_counter()
    plan = dev.DistancePlan(seqs, length=L, allchars=allchars, engine=_capi.ENGINE_POPC)
    # what ENGINE_AUTO does in the level-1/2 calls: the tensor-core engine when its operands fit in HBM
    need = _capi.lib().btb_dist_operand_bytes(_capi.ENGINE_UMMA, n, L, plan.sigma) + 2 * n * n * 2 + (8 << 30)
    umma = need < torch.cuda.mem_get_info()[0] + plan.operands.numel()
    res["engine"] = "umma" if umma else "popc"
    if umma:
        del plan
        torch.cuda.empty_cache()
        plan = dev.DistancePlan(seqs, length=L, allchars=allchars, engine=_capi.ENGINE_UMMA)
    res["sigma"] = plan.sigma
    res["operand_GB"] = plan.operands.numel() / 1e9
    out = plan.alloc_matrix()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); plan.encode(); e[1].record(); plan.compute(out); e[2].record()
    torch.cuda.synchronize()
    res["encode_ms"], res["distance_ms"] = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    pairs = n * (n - 1) / 2
    res["pairs_per_s"] = pairs / ((res["encode_ms"] + res["distance_ms"]) * 1e-3)
    sig = max(1, abs(plan.sigma))
    res["tops_algorithmic"] = 2.0 * sig * L * pairs / (res["distance_ms"] * 1e-3) / 1e12
    # properties over the whole matrix, in row chunks (the matrix is 20 GB at n = 100,000)
    m = out[:, :n]
    sym, diag0, top = True, True, 0
    for r0 in range(0, n, 4096):
        r1 = min(n, r0 + 4096)
        sym &= bool((m[r0:r1, :] == m[:, r0:r1].T).all().item())
        diag0 &= bool((torch.diagonal(m[r0:r1, r0:r1]) == 0).all().item())
        top = max(top, int(m[r0:r1, :].to(torch.int32).max().item()))
    res["symmetric"], res["zero_diagonal"], res["max"] = sym, diag0, top
    # engine agreement: the other engine on the first / middle / last block rows
    del plan.operands
    torch.cuda.empty_cache()
    if umma:
        p2 = dev.DistancePlan(seqs, length=L, allchars=allchars, engine=_capi.ENGINE_POPC)
        p2.encode()
        o2 = torch.zeros_like(out)
        T = dev.block_rows(n)
        agree = True
        for lo, hi in ((0, 1), (T // 2, T // 2 + 1), (T - 1, T)):
            p2.compute_rows(o2, lo, hi, mirror=False)
            torch.cuda.synchronize()
            r0, r1 = lo * 256, min(n, hi * 256)
            agree &= bool((o2[r0:r1, r0:n] == out[r0:r1, r0:n]).all().item())
        res["popc_rows_agree"] = agree
        del o2, p2
        torch.cuda.empty_cache()
    if oracle_rows:
        from oracle import oracle
        h = seqs[:, :L].cpu().numpy()
        picks = sorted(set([0, n - 1] + list(np.random.default_rng(1).integers(0, n, size=max(0, oracle_rows - 2)))))
        ok = True
        t1 = time.perf_counter()
        for r in picks:
            want = oracle.snp_dists_rows(h, allchars=allchars, keepcase=False, threads=threads, row0=int(r), row1=int(r) + 1)[0]
            ok &= bool((want == m[r].cpu().numpy().astype(np.uint32)).all())
        res["oracle_rows_checked"], res["oracle_rows_match"] = len(picks), ok
        res["oracle_s_per_row"] = (time.perf_counter() - t1) / len(picks)
    res["wall_s"] = time.perf_counter() - t0
    del out, m
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("n", nargs="?", type=int, default=100000)
    ap.add_argument("L", nargs="?", type=int, default=60000)
    ap.add_argument("--rows", type=int, default=6)
    args = ap.parse_args()
    from btb_phylo_b200 import build
    build.build()
    device = torch.device("cuda", 0)
    t0 = time.perf_counter()
    seqs = make_c5(args.n, args.L, device)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    h = seqs[:64, :args.L].cpu().numpy()
    frac_nonacgt = float(np.isin(h, list(b"ACGTacgt"), invert=True).mean())
    line = {"workload": "C5: %d samples x %d sites, %.0f %% non-ACGT cells" % (args.n, args.L, 100 * frac_nonacgt),
            "generation_s": gen_s, "threads": os.cpu_count()}
    for name, allchars in (("allchars(-a)", True), ("default", False)):
        line[name] = run_mode(seqs, args.n, args.L, allchars, args.rows, os.cpu_count() or 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
