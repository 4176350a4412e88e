#!/usr/bin/env python
"""Incremental snp-sites stage (btb_phylo_b200/plane_store.py): wall time of snp_sites() over N
per-sample consensus files -- plain, with an empty plane store, with every sample in the store, and for a "weekly"
run in which 1 % of the samples are new -- and the byte-identity of the four snps.fas files.

    python tools/bench_store.py [--samples 2000] [--genome 4349904]
"""
import argparse
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=2000)
    ap.add_argument("--genome", type=int, default=4349904)
    ap.add_argument("--snps", type=int, default=20000)
    ap.add_argument("--threads", type=int, default=min(64, os.cpu_count() or 1))
    ap.add_argument("--dir", default="/dev/shm/btb_store_bench")
    ap.add_argument("--store-dir", default=None, help="where the plane store goes (default: under --dir)")
    a = ap.parse_args()
    import torch
    import bench_northstar as ns
    from btb_phylo_b200 import build, phylogeny
    build.build()
    os.makedirs(a.dir, exist_ok=True)
    g = argparse.Namespace(samples=a.samples, genome=a.genome, snps=a.snps, files=a.samples, writers=8, dir=a.dir, seed=5,
                           n_frac=0.01, private_n=None)
    paths, names, _, nbytes, gen_s = ns.generate(g, torch, torch.device("cuda", 0))
    store = os.path.join(a.store_dir or a.dir, "store")
    out = {"n_samples": a.samples, "genome": a.genome, "fasta_bytes": nbytes, "files": len(paths), "host_threads": a.threads,
           "inputs_dir": a.dir, "store_dir": store, "generate_s": gen_s}
    res = {}

    def run(label, **kw):
        sf = os.path.join(a.dir, "snps_%s.fas" % label)
        t0 = time.perf_counter()
        meta = phylogeny.snp_sites(sf, paths, threads=a.threads, **kw)
        res[label] = {"snp_sites_s": time.perf_counter() - t0, "number_of_snps": meta["number_of_snps"]}
        return open(sf, "rb").read()
    try:
        run("warmup")                                              # CUDA context, library load
        plain = run("plain")
        cold = run("store_empty", plane_store=store)
        res["store_empty"]["store_bytes"] = sum(os.path.getsize(os.path.join(store, f)) for f in os.listdir(store) if f.endswith(".planes"))
        warm = run("store_full", plane_store=store)
        for f in paths[:: 100]:                                    # 1 % of the samples are "new this week": another mtime
            st = os.stat(f)
            os.utime(f, ns=(st.st_atime_ns, st.st_mtime_ns + 1000000000))
        weekly = run("store_1pct_new", plane_store=store)
        out["runs"] = res
        out["identical_snps_fas"] = plain == cold == warm == weekly
    finally:
        shutil.rmtree(a.dir, ignore_errors=True)
        shutil.rmtree(store, ignore_errors=True)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
