#!/usr/bin/env python
"""BASELINE metric 2 for configs 2 / 3: wall clock from entering snp_sites() to snps.csv closed,
through the reference-facing entry points, next to the CPU tools (the oracle's programs) on the same
synthetic multi-FASTA.

    python tools/bench_pipeline.py [--samples 2000] [--genome 4349904] [--snps 20000] [--cpu]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=2000)
    ap.add_argument("--genome", type=int, default=4349904)
    ap.add_argument("--snps", type=int, default=20000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    from btb_phylo_b200 import build, phylogeny
    import bench_extract
    build.build()
    tmp = tempfile.mkdtemp(prefix="btb_pipe_")
    mf = os.path.join(tmp, "multi_fasta.fas")
    t0 = time.perf_counter()
    size = bench_extract.make_multifasta(mf, args.samples, args.genome, args.snps)
    gen_s = time.perf_counter() - t0
    runs = []
    for label in ("first_call", "warm"):
        sf, sc = os.path.join(tmp, "snps.fas"), os.path.join(tmp, "snps.csv")
        for f in (sf, sc):
            if os.path.exists(f):
                os.remove(f)
        t0 = time.perf_counter()
        meta = phylogeny.snp_sites(sf, mf, threads=args.threads)
        t1 = time.perf_counter()
        phylogeny.build_snp_matrix(sc, sf, args.threads, post_process=False)
        t2 = time.perf_counter()
        runs.append({"label": label, "snp_sites_s": t1 - t0, "build_snp_matrix_s": t2 - t1, "snps_csv_wall_s": t2 - t0})
    out = {"metric": "snps.csv wall time (snp_sites entry -> snps.csv closed)", "n_samples": args.samples, "genome": args.genome,
           "fasta_bytes": size, "number_of_snps": meta["number_of_snps"], "snps_csv_bytes": os.path.getsize(sc),
           "host_threads": args.threads, "runs": runs, "synthetic_generation_s": gen_s}
    if args.cpu:
        from oracle import oracle
        osf, osc = os.path.join(tmp, "o.fas"), os.path.join(tmp, "o.csv")
        t0 = time.perf_counter()
        assert oracle.run_snp_sites(mf, osf).returncode == 0
        t1 = time.perf_counter()
        assert oracle.run_snp_dists(osf, osc, threads=args.threads).returncode == 0
        t2 = time.perf_counter()
        out["cpu"] = {"kind": "port (restated snp-sites 2.5.1 single-threaded like upstream; snp-dists 0.8.2 -j %d)" % args.threads,
                      "snp_sites_s": t1 - t0, "snp_dists_s": t2 - t1, "snps_csv_wall_s": t2 - t0,
                      "identical_snps_fas": open(osf, "rb").read() == open(sf, "rb").read(),
                      "identical_snps_csv": open(osc, "rb").read() == open(sc, "rb").read()}
    print(json.dumps(out), flush=True)
    for fn in os.listdir(tmp):
        os.remove(os.path.join(tmp, fn))
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
