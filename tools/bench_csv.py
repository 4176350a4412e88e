#!/usr/bin/env python
"""snps.csv wall time of the reference-facing build_snp_matrix() (file in, file out) on a
synthetic snps.fas, next to the CPU oracle's snp-dists on a row sample.

    python tools/bench_csv.py [--samples 20000] [--sites 5000] [--threads N] [--out-dir DIR]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=20000)
    ap.add_argument("--sites", type=int, default=5000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out-dir", default=None)
    args = ap.parse_args()
    from btb_phylo_b200 import build, phylogeny, synth
    build.build()
    n, L = args.samples, args.sites
    tmp = args.out_dir or tempfile.mkdtemp(prefix="btb_csv_")
    os.makedirs(tmp, exist_ok=True)
    sf, sc = os.path.join(tmp, "snps.fas"), os.path.join(tmp, "snps.csv")
    t0 = time.perf_counter()
    seqs = synth.snp_alignment(n, L, seed=4)
    synth.write_fasta(sf, synth.sample_names(n), seqs, wrap=0)
    t_gen = time.perf_counter() - t0
    res = {}
    for label in ("first_call", "second_call"):
        t0 = time.perf_counter()
        phylogeny.build_snp_matrix(sc, sf, threads=args.threads)
        res[label] = time.perf_counter() - t0
    size = os.path.getsize(sc)
    out = {"stage": "build_snp_matrix (snp-dists -c -j T), file to file", "n_samples": n, "n_sites": L,
           "host_threads": args.threads, "snps_csv_bytes": size, "cells": n * n, "wall_s": res,
           "cells_per_s": n * n / res["second_call"], "csv_bytes_per_s": size / res["second_call"],
           "out_dir": tmp, "synthetic_generation_s": t_gen}
    print(json.dumps(out), flush=True)
    if not args.out_dir:
        for fn in os.listdir(tmp):
            os.remove(os.path.join(tmp, fn))
        os.rmdir(tmp)


if __name__ == "__main__":
    main()
