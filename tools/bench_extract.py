#!/usr/bin/env python
"""Extraction-stage measurement (BASELINE config 2 shape): k1 pack, k2 column
mask, k3 select/compact against the HBM roofline, and the wall time of the reference-facing
snp_sites() on a synthetic multi-FASTA next to the CPU oracle's snp-sites.

    python tools/bench_extract.py [--samples 2000] [--genome 4349904] [--snps 20000] [--cpu]

Writes one JSON line.  Not the driver's bench (bench.py measures the distance stage the metric is
quoted on); this is the second half of the north_star's reporting."""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_multifasta(path, n, G, n_snps, wrap=60, seed=2):
    """n records of G bases, 60-column lines, clade SNPs, ~1 % N in shared masked blocks; written
    straight from GPU-generated rows.  Returns the file size."""
    import torch
    from btb_phylo_b200 import synth
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ref_i = torch.randint(0, 4, (G,), generator=g, device=dev)
    cols = torch.randperm(G, generator=g, device=dev)[:n_snps].sort().values
    minor_i = (ref_i[cols] + torch.randint(1, 4, (n_snps,), generator=g, device=dev)) % 4
    u = torch.rand(n_snps, generator=g, device=dev, dtype=torch.float64)
    size = torch.clamp((float(n) ** u * 0.5 + 0.5).to(torch.int64), 1, max(1, n - 1))
    a = (torch.rand(n_snps, generator=g, device=dev, dtype=torch.float64) * (n - size + 1).to(torch.float64)).to(torch.int64)
    b = a + size
    order = torch.randperm(n, generator=g, device=dev)
    nblk = max(1, int(G * 0.01 / 110))
    starts = torch.randint(0, G - 110, (nblk,), generator=g, device=dev)
    blk_mask = torch.zeros(G, dtype=torch.bool, device=dev)
    for k in range(110):
        blk_mask[starts + k] = True
    blk_id = torch.zeros(G, dtype=torch.int64, device=dev)
    blk_id[starts] = 1
    lines = (G + wrap - 1) // wrap
    text_len = G + lines
    # column -> text offset inside the sequence part
    colpos = torch.arange(G, device=dev)
    textpos = colpos + colpos // wrap
    names = synth.sample_names(n)
    ref = acgt[ref_i]
    total = 0
    with open(path, "wb") as f:
        for i in range(n):
            row = ref.clone()
            carries = (order[i] >= a) & (order[i] < b)
            row[cols[carries]] = acgt[minor_i[carries]]
            has = torch.rand(nblk, generator=g, device=dev) < 0.9
            drop = torch.zeros(G, dtype=torch.bool, device=dev)
            for k in range(110):
                drop[starts[has] + k] = True
            row[drop] = ord("N")
            text = torch.full((text_len,), 10, dtype=torch.uint8, device=dev)
            text[textpos] = row
            head = (">" + names[i] + "\n").encode()
            f.write(head)
            f.write(text.cpu().numpy().tobytes())
            total += len(head) + text_len
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=400)
    ap.add_argument("--genome", type=int, default=4349904)
    ap.add_argument("--snps", type=int, default=20000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--cpu", action="store_true", help="also time the oracle's snp-sites on the same file")
    ap.add_argument("--keep", default=None)
    args = ap.parse_args()
    import numpy as np
    import torch
    from btb_phylo_b200 import _capi, build, phylogeny
    build.build()
    lib = _capi.lib()
    n, G = args.samples, args.genome
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    tmp = args.keep or tempfile.mkdtemp(prefix="btb_extract_")
    os.makedirs(tmp, exist_ok=True)
    mf = os.path.join(tmp, "multi_fasta.fas")
    t0 = time.perf_counter()
    size = make_multifasta(mf, n, G, args.snps)
    t_gen = time.perf_counter() - t0

    # ---- kernels on device-resident text (one chunk of records), CUDA events ----------------------
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    chunk_n = min(n, 48)
    raw = np.fromfile(mf, dtype=np.uint8, count=int(size * chunk_n / n) + 64)
    offs, pos = [], 0
    for i in range(chunk_n):
        nl = int(np.argmax(raw[pos:pos + 256] == 10))
        offs.append(pos + nl + 1)
        pos = offs[-1] + G + (G + 59) // 60
    text_bytes = pos
    d_text = torch.from_numpy(raw[: text_bytes + 64].copy()).cuda()
    d_off = torch.tensor(offs, dtype=torch.int64).cuda()
    d_lw = torch.full((chunk_n,), 60, dtype=torch.int32).cuda()
    d_el = torch.ones(chunk_n, dtype=torch.int32).cuda()
    words = lib.btb_plane_words(G)
    n_res = min(n, 2000)
    planes = torch.zeros(3 * n_res * words, dtype=torch.int32, device="cuda")
    flags = torch.zeros(n_res, dtype=torch.int32, device="cuda")

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best * 1e-3

    def pack_all():
        for first in range(0, n_res, chunk_n):
            m = min(chunk_n, n_res - first)
            _capi.check(lib.btb_pack_planes(P(d_text), text_bytes, P(d_off), P(d_lw), P(d_el), m, first, n_res, G, P(planes), P(flags[first:]), st))

    t_pack = timed(pack_all, reps=3)
    pack_bytes = n_res * (G + G // 60) + n_res * 3 * words * 4
    state = torch.zeros(5 * words, dtype=torch.int32, device="cuda")

    def pack_all_fused():                                  # what btb_snp_sites runs: k1 with the column state (k2) fused in
        state.zero_()
        for first in range(0, n_res, chunk_n):
            m = min(chunk_n, n_res - first)
            _capi.check(lib.btb_pack_planes_state(P(d_text), text_bytes, P(d_off), P(d_lw), P(d_el), m, first, n_res, G, P(planes), P(flags[first:]), P(state), st))

    t_fused = timed(pack_all_fused, reps=3)
    state_fused = state.clone()
    t_mask = timed(lambda: _capi.check(lib.btb_column_mask(P(planes), n_res, n_res, G, P(state), 0, st)))
    mask_bytes = n_res * 3 * words * 4
    # the fused state leaves the padding bits (columns >= G) alone, k2 marks them invalid: compare the real columns
    real = torch.zeros(words * 32, dtype=torch.bool, device="cuda")
    real[:G] = True
    real_words = (real.view(words, 32).to(torch.int64) << torch.arange(32, device="cuda")[None, :]).sum(dim=1).to(torch.int32)
    fused_equals_k2 = bool(((((state_fused ^ state).view(5, words)) & real_words[None, :]) == 0).all().item())
    keep = torch.zeros(words, dtype=torch.int32, device="cuda")
    idx = torch.zeros(words * 32, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    t_sel = timed(lambda: _capi.check(lib.btb_select_columns(P(state), G, 1, P(keep), P(idx), P(cnt), st)))
    L = int(cnt[0].item())
    ld = (max(L, 1) + 15) // 16 * 16
    snps = torch.zeros(n_res * ld, dtype=torch.uint8, device="cuda")
    t_cmp = timed(lambda: _capi.check(lib.btb_compact_columns(P(planes), 0, n_res, n_res, G, P(idx), L, P(snps), ld, st))) if L else 0.0
    irregular = int(flags.sum().item())
    del planes, snps, d_text
    torch.cuda.empty_cache()

    # ---- the reference-facing call, wall clock ------------------------------------------------------
    sf = os.path.join(tmp, "snps.fas")
    t0 = time.perf_counter()
    meta = phylogeny.snp_sites(sf, mf, threads=args.threads)
    t_e2e = time.perf_counter() - t0
    t0 = time.perf_counter()
    meta = phylogeny.snp_sites(sf, mf, threads=args.threads)           # page cache warm
    t_e2e_warm = time.perf_counter() - t0
    out = {
        "stage": "extraction (snp-sites -c)", "n_samples": n, "genome": G, "fasta_bytes": size, "number_of_snps": meta["number_of_snps"],
        "resident_samples_for_kernel_timing": n_res, "irregular_flags": irregular,
        "k1_pack_with_column_state": {"seconds": t_fused, "algorithmic_bytes": pack_bytes, "GBps": pack_bytes / t_fused / 1e9,
                                      "frac_of_hbm_peak": pack_bytes / t_fused / 1e9 / hbm, "records_per_launch": chunk_n,
                                      "state_equals_separate_k2": fused_equals_k2},
        "k1_pack": {"seconds": t_pack, "algorithmic_bytes": pack_bytes, "GBps": pack_bytes / t_pack / 1e9, "frac_of_hbm_peak": pack_bytes / t_pack / 1e9 / hbm},
        "k2_column_mask": {"seconds": t_mask, "algorithmic_bytes": mask_bytes, "GBps": mask_bytes / t_mask / 1e9, "frac_of_hbm_peak": mask_bytes / t_mask / 1e9 / hbm},
        "k3_select": {"seconds": t_sel}, "k3_compact": {"seconds": t_cmp, "L": L},
        "hbm_peak_GBps": hbm,
        "snp_sites_wall_s": {"first_call": t_e2e, "page_cache_warm": t_e2e_warm, "host_threads": args.threads,
                             "bytes_per_s_warm": size / t_e2e_warm},
        "synthetic_generation_s": t_gen,
    }
    if args.cpu:
        from oracle import oracle
        osf = os.path.join(tmp, "o.fas")
        t0 = time.perf_counter()
        r = oracle.run_snp_sites(mf, osf)
        out["cpu_snp_sites_wall_s"] = time.perf_counter() - t0
        out["cpu_kind"] = "port (restated snp-sites 2.5.1 -c, single thread like upstream)"
        out["snps_fas_identical"] = r.returncode == 0 and open(sf, "rb").read() == open(osf, "rb").read()
    print(json.dumps(out), flush=True)
    if not args.keep:
        for fn in os.listdir(tmp):
            os.remove(os.path.join(tmp, fn))
        os.rmdir(tmp)


if __name__ == "__main__":
    main()
