#!/usr/bin/env python
"""Development probe for the GPU box: runs each distance-engine configuration in its own process
under a timeout (a wrong mbarrier protocol hangs rather than fails) and logs parity against the
CPU oracle plus a first timing.  Output: gpurun_out/probe.log.  Not part of the product."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(cfg):
    import numpy as np
    import torch
    from btb_phylo_b200 import _capi, device, synth
    from oracle import oracle
    n, L = cfg["n"], cfg["L"]
    eng = {"popc": _capi.ENGINE_POPC, "umma": _capi.ENGINE_UMMA}[cfg["engine"]]
    if cfg.get("cg"):
        _capi.check(_capi.lib().btb_set_option(b"umma_cta_group", cfg["cg"]))
    seqs = synth.snp_alignment(n, L, seed=cfg.get("seed", 4), heavy_n=cfg.get("heavy_n", 0.0),
                               iupac=cfg.get("iupac", 0.0), lower=cfg.get("lower", 0.0))
    stride = (L + 15) // 16 * 16
    host = np.full((n, stride), ord("."), dtype=np.uint8)
    host[:, :L] = seqs
    d = torch.from_numpy(host).cuda()
    plan = device.DistancePlan(d, length=L, allchars=cfg.get("allchars", False), engine=eng)
    plan.encode()
    out = plan.alloc_matrix()
    torch.cuda.synchronize()
    plan.compute(out)
    torch.cuda.synchronize()
    res = {"cfg": cfg, "sigma": plan.sigma}
    if cfg.get("check", True):
        rows = min(n, cfg.get("check_rows", 64))
        sel = np.unique(np.concatenate([np.arange(min(rows // 2, n)), np.linspace(0, n - 1, rows).astype(np.int64)]))
        got = out[:, :n].cpu().numpy().astype(np.int64)
        bad = 0
        for r in sel:
            ref = oracle.snp_dists_rows(seqs, allchars=cfg.get("allchars", False), threads=8, row0=int(r), row1=int(r) + 1)[0].astype(np.int64)
            diff = np.nonzero(ref != got[r])[0]
            if len(diff):
                if bad < 3:
                    res.setdefault("examples", []).append({"row": int(r), "cols": diff[:8].tolist(), "ref": ref[diff[:8]].tolist(), "got": got[r][diff[:8]].tolist(), "ndiff": int(len(diff))})
                bad += 1
        res["rows_checked"] = int(len(sel))
        res["rows_bad"] = bad
        res["symmetric"] = bool((got == got.T).all())
        res["diag_zero"] = bool((np.diag(got) == 0).all())
    if cfg.get("time", 0):
        reps = cfg["time"]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        plan.compute(out)
        ev[0].record()
        for i in range(reps):
            plan.compute(out)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
        pairs = n * (n - 1) / 2
        res["ms"] = ms
        res["pairs_per_s"] = pairs / (min(ms) * 1e-3)
        res["tops_sigma4"] = 2 * 4 * L * pairs / (min(ms) * 1e-3) / 1e12
    print("RESULT " + json.dumps(res))


CONFIGS = [
    {"n": 300, "L": 1000, "engine": "popc"},
    {"n": 300, "L": 1000, "engine": "umma", "cg": 1},
    {"n": 300, "L": 1000, "engine": "umma", "cg": 2},
    {"n": 1, "L": 7, "engine": "popc"},
    {"n": 1, "L": 7, "engine": "umma", "cg": 2},
    {"n": 777, "L": 831, "engine": "umma", "cg": 2, "heavy_n": 0.2, "lower": 0.01},
    {"n": 777, "L": 831, "engine": "popc", "heavy_n": 0.2, "lower": 0.01},
    {"n": 515, "L": 2049, "engine": "popc", "allchars": True, "heavy_n": 0.25, "iupac": 0.01, "lower": 0.01},
    {"n": 515, "L": 2049, "engine": "umma", "cg": 2, "allchars": True, "heavy_n": 0.25, "iupac": 0.01, "lower": 0.01},
    {"n": 515, "L": 2049, "engine": "umma", "cg": 1, "allchars": True, "heavy_n": 0.25, "iupac": 0.01, "lower": 0.01},
    {"n": 4096, "L": 30000, "engine": "umma", "cg": 2, "check_rows": 16, "time": 3},
    {"n": 4096, "L": 30000, "engine": "umma", "cg": 1, "check_rows": 16, "time": 3},
    {"n": 4096, "L": 30000, "engine": "popc", "check_rows": 16, "time": 3},
    {"n": 16384, "L": 30000, "engine": "umma", "cg": 2, "check_rows": 8, "time": 3},
    {"n": 16384, "L": 30000, "engine": "umma", "cg": 1, "check_rows": 8, "time": 2},
    {"n": 16384, "L": 30000, "engine": "popc", "check_rows": 8, "time": 2},
]


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(json.loads(sys.argv[2]))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "probe.log"), "a")
    cfgs = CONFIGS
    if len(sys.argv) > 1:
        cfgs = json.loads(sys.argv[1])
    for cfg in cfgs:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--child", json.dumps(cfg)], capture_output=True, text=True,
                               timeout=cfg.get("timeout", 240))
            if r.stderr.strip():
                with open(os.path.join(ROOT, "gpurun_out", "probe_stderr.log"), "a") as ef:
                    ef.write("==== %s\n%s\n" % (json.dumps(cfg), r.stderr[-200000:]))
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            msg = lines[-1] if lines else "NO RESULT rc=%d\n%s\n%s" % (r.returncode, r.stdout[-2000:], r.stderr[-3000:])
        except subprocess.TimeoutExpired as e:
            msg = "TIMEOUT %s" % json.dumps(cfg)
        line = "[%6.1fs] %s" % (time.time() - t0, msg)
        print(line, flush=True)
        log.write(line + "\n")
        log.flush()


if __name__ == "__main__":
    main()
