#!/usr/bin/env python
"""bench.py -- pairwise SNP distances/s on BASELINE.json config 4 (50,000 samples x 30,000 SNP
sites, distance stage), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over the whole alignment: SNP alignment bytes resident in HBM
-> operand encode -> tcgen05 kind::mxf4 distance GEMM (exact 0/1 products) over this rank's block rows
-> distances in HBM.  `value` = unique pairs N(N-1)/2 of the whole job / max-over-ranks device
time.  `e2e` = the same through the C-ABI's host-buffer call: page-locked host alignment in, H2D,
encode, GEMM, D2H, page-locked host matrix out (N = 1: btb_dist_matrix_host; N > 1:
btb_dist_matrix_host_comm -- sharded upload, operand shards GPU to GPU, ONE shared host matrix).
`snps_csv` = wall clock of build_snp_matrix() file to file (BASELINE metric, second half).
`--impl reference` times the CPU arm (the restated snp-dists 0.8.2 loop nest, all host threads)
on a bounded sample of the same workload.  SURVEY.md section 8d defines work and peaks.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = int(os.environ.get("BTB_BENCH_N", 50000))
N_SITES = int(os.environ.get("BTB_BENCH_L", 30000))
SIGMA = 4
WORKLOAD = "C4: %d samples x %d SNP sites (synthetic clade model, ACGT), snp-dists default mode, full matrix" % (N_SAMPLES, N_SITES)


def unique_pairs(n):
    return n * (n - 1) // 2


# ------------------------------------------------------------------------------ synthetic data
def make_alignment_cuda(n, L, device, seed=4, row_chunk=4096):
    """The clade model of btb_phylo_b200/synth.py, generated on the GPU (1.5 GB in < 1 s)."""
    import torch
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
    stride = (L + 15) // 16 * 16
    out = torch.full((n, stride), ord("."), dtype=torch.uint8, device=device)
    for r0 in range(0, n, row_chunk):
        pos = order[r0:r0 + row_chunk, None]
        m = (pos >= a[None, :]) & (pos < b[None, :])
        out[r0:r0 + row_chunk, :L] = torch.where(m, minor[None, :], base[None, :])
    return out


def make_alignment_host(n, L, seed=4):
    from btb_phylo_b200 import synth
    return synth.snp_alignment(n, L, seed=seed)


# ------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_lo=None, t_hi=None):
        """Summary of the samples taken in [t_lo, t_hi] (the timed region).  A region shorter than the
        sampling period may hold none: then the samples since start() (warm-up + timed steps, the same
        work) are used and the summary says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = "timed region"
        rows = [r for t, r in self.rows if (t_lo is None or t >= t_lo) and (t_hi is None or t <= t_hi)]
        if not rows:
            rows, window = [r for _, r in self.rows], "warm-up + timed region (timed region shorter than the sampling period)"
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons), "window": window}


# ------------------------------------------------------------------------------ CPU arm
class CpuArm:
    """The oracle's restatement of snp-dists 0.8.2, timed the way the tool runs: the alignment is loaded
    and prepared ONCE (upper-casing, non-ACGT -> '.'; outside every timed region, as in snp-dists' main),
    then matrix rows are computed by the loop nest (outer row serial, inner OpenMP parallel-for over all
    host threads).  Cost is exactly linear in rows, so R rows are timed and the whole job is R -> n."""

    MIN_ROWS = 200

    def __init__(self, seqs_host, threads):
        from oracle import oracle
        self.n, self.L = seqs_host.shape
        self.threads = threads
        self.prep = oracle.PreparedAlignment(seqs_host)
        self.prep.rows(0, min(2, self.n), threads=threads)          # OpenMP pool start-up, page touching
        t0 = time.perf_counter()
        self.prep.rows(0, min(8, self.n), threads=threads)
        self.t_row_probe = (time.perf_counter() - t0) / min(8, self.n)

    def rows_for(self, budget_s):
        return int(min(self.n, max(self.MIN_ROWS, budget_s / max(self.t_row_probe, 1e-9))))

    def time_rows(self, rows, row0=0):
        """-> (seconds, uint32 [rows, n]) for matrix rows [row0, row0 + rows)"""
        t0 = time.perf_counter()
        got = self.prep.rows(row0, row0 + rows, threads=self.threads)
        return time.perf_counter() - t0, got

    def sample_text(self, rows, t_row):
        return ("%d of %d matrix rows (each row = %d full-length compares; cost linear in rows; alignment prepared once "
                "outside the timed region, as snp-dists does at load); whole job = x%.0f = %.0f s for the full square "
                "snp-dists computes; restated snp-dists 0.8.2 loop nest, gcc -O3 -fopenmp, %d threads"
                % (rows, self.n, self.n, self.n / rows, t_row * self.n, self.threads))


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n, L = N_SAMPLES, N_SITES
    arm = CpuArm(make_alignment_host(n, L), threads)
    # every step = the same bounded sample; the whole run stays within a few minutes
    per_step_budget = max(2.0, min(20.0, 200.0 / max(1, args.steps + args.warmup)))
    rows = arm.rows_for(per_step_budget)
    times = []
    for s in range(args.warmup + args.steps):
        dt, _ = arm.time_rows(rows, row0=(s * rows) % max(1, n - rows))
        if s >= args.warmup:
            times.append(dt / rows)
    t_row = statistics.mean(times)
    value = unique_pairs(n) / (t_row * n)
    line = {
        "impl": "reference", "metric": "pairwise SNP distances/sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_row * rows * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_samples": n, "n_sites": L},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": "per step: " + arm.sample_text(rows, t_row),
                         "t_row_s": t_row, "t_row_spread": (max(times) - min(times)) / t_row if times else None},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers
    of the end-to-end leg are allocated next to its PCIe root (first-touch policy).  Returns a short
    description for the JSON line; does nothing when the topology cannot be read or the container's
    cpuset has no CPU on that node."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return "gpu %s: no NUMA node reported" % bdf
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return "gpu %s on node %d: none of this process's %d cpus is on that node" % (bdf, node, len(os.sched_getaffinity(0)))
        os.sched_setaffinity(0, allowed)
        return "gpu %s on node %d: bound to %d of its cpus" % (bdf, node, len(allowed))
    except Exception as e:                                           # noqa: BLE001 -- best effort, never fatal
        return "unavailable (%s)" % type(e).__name__


# ------------------------------------------------------------------------------ GPU arm
def _shared_array(path, shape, dtype, create):
    """A file under /dev/shm mapped by every rank: ONE alignment and ONE matrix for the whole job."""
    import numpy as np
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if create:
        with open(path, "wb") as f:
            f.truncate(nbytes)
    return np.memmap(path, dtype=dtype, mode="r+", shape=shape)


def _page_lock(arr):
    import torch
    rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
    if int(rc) != 0:
        raise RuntimeError("cudaHostRegister failed: %s" % rc)


def _page_unlock(arr):
    import torch
    torch.cuda.cudart().cudaHostUnregister(arr.ctypes.data)


def e2e_single(lib, _capi, torch, seqs, plan, n, L, local, steps):
    """N = 1: btb_dist_matrix_host, pinned host alignment in, pinned host matrix out."""
    h_seqs = torch.empty((n, seqs.shape[1]), dtype=torch.uint8, pin_memory=True)
    h_seqs.copy_(seqs)
    ld = (n + 7) // 8 * 8
    h_out = torch.zeros((n, ld), dtype=plan.out_dtype, pin_memory=True)

    def call():
        _capi.check(lib.btb_dist_matrix_host(ctypes.c_void_p(h_seqs.data_ptr()), n, L, h_seqs.stride(0), 0, 0,
                                             _capi.ENGINE_UMMA, local, 0, 1, ctypes.c_void_p(h_out.data_ptr()),
                                             plan.elem, ld), "btb_dist_matrix_host")
    call()                                                         # warm-up (allocations, page touching)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / steps
    info = {"call": "btb_dist_matrix_host (C-ABI level 2): pinned host alignment -> H2D -> encode -> distance kernels -> D2H band by band -> pinned host matrix",
            "h2d_bytes_per_step": int(n * L), "d2h_bytes_per_step": int(n * n * plan.elem)}
    # the same call delivering every unique pair once (upper triangle only): half the D2H bytes.  A second figure,
    # not the headline: `e2e` above is the full square snp-dists prints.
    h_up = torch.zeros((n, ld), dtype=plan.out_dtype, pin_memory=True)

    def call_upper():
        _capi.check(lib.btb_dist_matrix_host_upper(ctypes.c_void_p(h_seqs.data_ptr()), n, L, h_seqs.stride(0), 0, 0,
                                                   _capi.ENGINE_UMMA, local, 0, 1, ctypes.c_void_p(h_up.data_ptr()),
                                                   plan.elem, ld), "btb_dist_matrix_host_upper")
    call_upper()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        call_upper()
    torch.cuda.synchronize()
    up_s = (time.perf_counter() - t0) / steps
    import numpy as np
    rows = slice(0, min(n, 2048))
    same = bool((np.triu(h_up.numpy()[rows, :n]) == np.triu(h_out.numpy()[rows, :n])).all())
    info["upper_triangle_only"] = {"call": "btb_dist_matrix_host_upper: cells (r, c >= r) only", "seconds_per_step": up_s,
                                   "value": unique_pairs(n) / up_s, "unit": "pairs/s", "d2h_bytes_per_step": int(n * (n + 1) // 2 * plan.elem),
                                   "equals_full_matrix_on_first_rows": same}
    del h_up
    return e2e_s, info, h_seqs.numpy()[:, :L], h_out.numpy()[:, :n], (h_seqs, h_out)


def e2e_multi(lib, _capi, torch, dist, host_pg, seqs, plan, n, L, rank, world, local, steps):
    """N > 1: btb_dist_matrix_host_comm.  ONE page-locked alignment and ONE page-locked matrix in shared
    memory; rank r uploads and encodes samples [n r / N, n (r+1) / N), the operand shards are exchanged
    GPU to GPU (CUDA IPC peer copies), every rank copies its block rows into the one matrix."""
    import numpy as np
    tag = [None]
    if rank == 0:
        tag[0] = "btbphylo_bench_%d_%s" % (os.getpid(), os.urandom(4).hex())
    dist.broadcast_object_list(tag, src=0, group=host_pg)
    tag = tag[0]
    stride = seqs.shape[1]
    p_seqs, p_mat = "/dev/shm/%s.seqs" % tag, "/dev/shm/%s.mat" % tag
    out_np = np.uint16 if plan.elem == 2 else np.uint32
    if rank == 0:
        a = _shared_array(p_seqs, (n, stride), np.uint8, True)
        a[:] = seqs.cpu().numpy()
        _shared_array(p_mat, (n, n), out_np, True)
        del a
    dist.barrier(group=host_pg)
    h_seqs = _shared_array(p_seqs, (n, stride), np.uint8, False)
    h_out = _shared_array(p_mat, (n, n), out_np, False)
    dist.barrier(group=host_pg)
    if rank == 0:                       # every rank has its mappings: the names can go now (nothing is left behind if a step fails)
        os.unlink(p_seqs)
        os.unlink(p_mat)
    lo, hi = n * rank // world, n * (rank + 1) // world
    h_out[lo:hi] = 0                                               # first touch of "my" rows next to my GPU
    _page_lock(h_seqs)
    _page_lock(h_out)
    comm = ctypes.c_void_p()
    _capi.check(lib.btb_comm_create(tag.encode(), rank, world, local, ctypes.byref(comm)), "btb_comm_create")

    def call():
        _capi.check(lib.btb_dist_matrix_host_comm(comm, ctypes.c_void_p(h_seqs.ctypes.data), n, L, stride, 0, 0, _capi.ENGINE_UMMA,
                                                  ctypes.c_void_p(h_out.ctypes.data), plan.elem, n), "btb_dist_matrix_host_comm")
    call()
    torch.cuda.synchronize()
    # what the box gives when every rank copies its share of the matrix to the host at the same time (plain
    # cudaMemcpy from a device buffer into the shared page-locked matrix): the ceiling of the D2H phase
    share = n * n * plan.elem // world
    d_probe = torch.empty(share, dtype=torch.uint8, device=seqs.device)
    h_probe = torch.from_numpy(h_out.reshape(-1).view(np.uint8)[rank * share:(rank + 1) * share])
    torch.cuda.synchronize()
    dist.barrier(group=host_pg)
    tc = time.perf_counter()
    h_probe.copy_(d_probe, non_blocking=True)
    torch.cuda.synchronize()
    copy_s = torch.tensor([time.perf_counter() - tc], dtype=torch.float64)
    dist.all_reduce(copy_s, op=dist.ReduceOp.MAX, group=host_pg)
    d2h_all = n * n * plan.elem / float(copy_s[0]) / 1e9
    del d_probe, h_probe
    dist.barrier(group=host_pg)
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / steps
    phases = (ctypes.c_double * 8)()
    lib.btb_comm_last_timings(comm, phases, 8)
    path = ctypes.c_int64(-1)
    lib.btb_query(b"comm_last_path", ctypes.byref(path))
    dist.barrier(group=host_pg)                                                  # every rank's cells are in the matrix
    _capi.check(lib.btb_comm_destroy(comm), "btb_comm_destroy")
    names = ["upload_shard+presence", "barrier_A", "buffers+site_counts", "barrier_B", "encode_shard", "wait+nvlink_pull", "kernels+d2h"]
    info = {"call": "btb_dist_matrix_host_comm (C-ABI level 2, one process per GPU): ONE page-locked alignment and ONE page-locked matrix in "
                    "shared memory; sharded upload + encode, operand shards GPU to GPU over NVLink (CUDA IPC peer copies, no NCCL), "
                    "block rows per rank, D2H band by band",
            "h2d_bytes_per_step": int(n * L), "d2h_bytes_per_step": int(n * n * plan.elem),
            "bytes_are": "summed over all ranks (each rank moves about 1/N of both)",
            "path": "sharded" if path.value == 1 else "replicated-upload fallback",
            "plain_d2h_all_ranks_GBps": d2h_all,
            "rank0_phase_ms": {k: round(phases[i] * 1e3, 3) for i, k in enumerate(names)}}

    def cleanup():
        _page_unlock(h_seqs)
        _page_unlock(h_out)
        dist.barrier(group=host_pg)
    return e2e_s, info, h_seqs[:, :L], h_out, cleanup


def csv_wall(seqs_host, n_gpus, threads):
    """BASELINE metric, second half: wall clock of build_snp_matrix() (the reference's entry point,
    btbphylo/phylogeny.py:144-150), snps.fas file in -> snps.csv file out, on tmpfs and on the box's disk."""
    from btb_phylo_b200 import phylogeny, synth
    n, L = seqs_host.shape
    res = {"entry_point": "btb_phylo_b200.phylogeny.build_snp_matrix(snps.csv, snps.fas, threads=%d), BTBPHYLO_GPUS=%d" % (threads, n_gpus),
           "n_gpus": n_gpus, "host_threads": threads}
    os.environ["BTBPHYLO_GPUS"] = str(n_gpus)
    names = synth.sample_names(n)
    for label, base in (("tmpfs", "/dev/shm"), ("disk", os.environ.get("BTB_BENCH_DISK_DIR", os.path.join(ROOT, "gpurun_out")))):
        d = os.path.join(base, "btb_bench_csv_%d" % os.getpid())
        try:
            os.makedirs(d, exist_ok=True)
            sf, sc = os.path.join(d, "snps.fas"), os.path.join(d, "snps.csv")
            synth.write_fasta(sf, names, seqs_host, wrap=0)
            walls = []
            for _ in range(2):                                      # first call of a process pays CUDA context + library load
                if os.path.exists(sc):
                    os.unlink(sc)
                t0 = time.perf_counter()
                phylogeny.build_snp_matrix(sc, sf, threads=threads)
                walls.append(time.perf_counter() - t0)
            size = os.path.getsize(sc)
            res[label] = {"csv_wall_s": min(walls), "first_call_s": walls[0], "snps_csv_bytes": size,
                          "csv_GBps": size / min(walls) / 1e9, "cells_per_s": n * n / min(walls), "dir": base}
        except Exception as exc:                                     # noqa: BLE001 -- e.g. no space: report, do not lose the run
            res[label] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        finally:
            for fn in ("snps.fas", "snps.csv"):
                try:
                    os.unlink(os.path.join(d, fn))
                except OSError:
                    pass
            try:
                os.rmdir(d)
            except OSError:
                pass
    return res


def gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from btb_phylo_b200 import _capi, build, device as dev, multirank

    build.build()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)
    host_pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        host_pg = dist.new_group(backend="gloo")                    # host-side barriers that keep the GPUs free
    lib = _capi.lib()
    n, L = N_SAMPLES, N_SITES

    seqs = make_alignment_cuda(n, L, device)                      # identical on every rank (operands are replicated)
    plan = dev.DistancePlan(seqs, length=L, engine=_capi.ENGINE_UMMA)
    # The path shards by block rows of the matrix: each rank owns a contiguous range of 256-row blocks; no
    # data-path collective.  A rank only encodes the rows its block rows meet (from its first row on), and the
    # ranges are cut so that  tiles + encode share  is the same for every rank (btb_dist_row_share_weighted).
    T = dev.block_rows(n)
    out = plan.alloc_matrix()
    encode_tiles = 0.0
    if world > 1:
        def _ms(fn, reps=3):
            best = 1e30
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best
        plan.encode()
        probe_lo = max(0, T - 64) // 2 * 2
        plan.compute_rows(out, probe_lo, T)
        torch.cuda.synchronize()
        t_enc = _ms(plan.encode)
        t_tile = _ms(lambda: plan.compute_rows(out, probe_lo, T)) / max(1, sum(T - i for i in range(probe_lo, T)))
        ratio = torch.tensor([t_enc / max(t_tile, 1e-9)], dtype=torch.float64, device=device)
        dist.broadcast(ratio, src=0)                                 # every rank must cut the same ranges
        encode_tiles = float(ratio[0])
    if os.environ.get("BTB_BENCH_SHARE") == "plain":                 # (diagnosis) equal tile counts, every rank encodes every row
        row_lo, row_hi = dev.row_share(n, rank, world)
        enc_row0 = 0
    else:
        row_lo, row_hi = dev.row_share_weighted(n, rank, world, encode_tiles) if world > 1 else (0, T)
        enc_row0 = max(0, row_lo * 256 - 256)                        # the 240-column units of the first block row start up to 239 rows early
    my_tiles = sum(T - i for i in range(row_lo, row_hi))

    def step():
        plan.encode(enc_row0, n)
        plan.compute_rows(out, row_lo, row_hi)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    step()                                                          # (lists, lazy allocations) before the sampler has anything to see
    sync_all()
    time.sleep(0.15)                                                # nvidia-smi needs ~0.1 s to deliver its first sample
    for _ in range(args.warmup):
        step()
    sync_all()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    t_timed0 = time.perf_counter()
    ev[0].record()
    for s in range(args.steps):
        plan.encode(enc_row0, n)
        ev[2 * s + 1].record()
        plan.compute_rows(out, row_lo, row_hi)
        ev[2 * s + 2].record()
    sync_all()
    clocks = sampler.stop(t_timed0, time.perf_counter())
    total_ms = ev[0].elapsed_time(ev[2 * args.steps])
    kern_ms = [ev[2 * s + 1].elapsed_time(ev[2 * s + 2]) for s in range(args.steps)]
    if os.environ.get("BTB_BENCH_VERBOSE"):
        enc_ms = [ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(args.steps)]
        sys.stderr.write("[bench] rank %d block rows [%d, %d) tiles %d: encode %.3f ms, kernel %.3f ms, step %.3f ms\n"
                         % (rank, row_lo, row_hi, my_tiles, statistics.mean(enc_ms), statistics.mean(kern_ms), total_ms / args.steps))
    t = torch.tensor([total_ms, statistics.mean(kern_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = unique_pairs(n) / (ms_per_step * 1e-3)
    del out
    torch.cuda.empty_cache()

    # ---- end to end through the C-ABI host call: page-locked host alignment in, page-locked host matrix out ----
    e2e_steps = max(1, min(args.steps, 5))
    if world == 1:
        e2e_s, e2e_info, h_seqs_np, h_out_np, keep = e2e_single(lib, _capi, torch, seqs, plan, n, L, local, e2e_steps)
        cleanup = None
    else:
        e2e_s, e2e_info, h_seqs_np, h_out_np, cleanup = e2e_multi(lib, _capi, torch, dist, host_pg, seqs, plan, n, L, rank, world, local, e2e_steps)
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])

    # what the box's PCIe link gives a plain pinned copy (context for e2e)
    def _copy_rate(dst, src):
        best = 0.0
        for _ in range(2):
            torch.cuda.synchronize()
            t0c = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            best = max(best, src.numel() * src.element_size() / (time.perf_counter() - t0c) / 1e9)
        return best
    pcie = None
    if rank == 0:
        try:
            probe_host = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
            probe_dev = torch.empty(1 << 30, dtype=torch.uint8, device=device)
            pcie = {"h2d_GBps": _copy_rate(probe_dev, probe_host), "d2h_GBps": _copy_rate(probe_host, probe_dev)}
            del probe_dev, probe_host
        except Exception as exc:                                     # noqa: BLE001 -- context only, never fatal
            pcie = {"unavailable": type(exc).__name__}
    e2e = {"value": unique_pairs(n) / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": e2e_info.pop("h2d_bytes_per_step"),
           "d2h_bytes_per_step": e2e_info.pop("d2h_bytes_per_step"), "seconds_per_step": e2e_s, "steps": e2e_steps,
           "host_numa": numa, "pcie_plain_copy_one_gpu": pcie}
    e2e.update(e2e_info)
    e2e["pcie_GBps_all_ranks"] = (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) / e2e_s / 1e9

    if rank != 0:
        if cleanup:
            cleanup()
        del plan, seqs
        lib.btb_release_workspace()
        torch.cuda.empty_cache()
        dist.barrier(group=host_pg)                                 # rank 0: CPU baseline + snps.csv wall (uses every GPU itself)
        dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the distance GEMM), measured live ---------------------
    fp4 = bool(_query(lib, b"umma_fp4")) and plan.sigma == _capi.SIGMA_DENSE4 and _query(lib, b"umma_dense_simplex") == 2
    pair = bool(_query(lib, b"umma_fp4_pair")) and fp4
    dense_code = _query(lib, b"umma_dense_simplex")
    planes_executed = 3 if (plan.sigma == _capi.SIGMA_DENSE4 and dense_code in (1, 2)) else 4
    share = my_tiles / multirank.tile_count(n)                       # rank 0's launch
    useful_ops = 2.0 * planes_executed * L * unique_pairs(n) * share
    algorithmic_ops = 2.0 * SIGMA * L * unique_pairs(n) * share
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    sm_mhz = clocks.get("sm_mhz") or 0.0
    # tensor-pipe issue rate of the MMA shape the kernel uses, measured on this GPU with a kernel that
    # issues the same instructions without data movement (btb_query probe_cg<cta_group>_e1: clk x 100 per
    # pipeline stage of 8 MMAs 128 x 240 x 64 per SM): flop / clk / SM
    probe = _query(lib, b"probe_cg2_e1" if pair else b"probe_cg1_e1") / 100.0 if fp4 else 0.0
    stage_flop = 8 * 2.0 * 128 * 240 * 64
    flop_per_clk_sm = stage_flop / probe if probe > 0 else (32768.0 if fp4 else 16384.0)
    peak = flop_per_clk_sm * sms * sm_mhz * 1e6 / 1e12 if sm_mhz else None
    achieved = useful_ops / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    bf16 = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0
    roofline = {"bound": "tensor",
                "kernel": ("dist_umma_fp4_pair_kernel (tcgen05.mma cta_group::2 kind::mxf4, exact 0/1 products)" if pair else
                           "dist_umma_fp4_tile_kernel (tcgen05.mma kind::mxf4, exact 0/1 products)") if fp4 else "dist_umma_tile_kernel (tcgen05.mma kind::i8)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": _ncu_traffic(n, L, world), "kernel_ms": kernel_ms,
                "achieved_is": "EXECUTED useful flop of this launch = 2 x %d elements per site x L x unique pairs of this rank's tiles / kernel time "
                               "(CUDA events on the launching stream); padding of L to the K step, diagonal units and cells below the diagonal are not counted" % planes_executed,
                "peak_is": "measured kind::mxf4 issue rate (%.0f flop/clk/SM from the data-free probe kernel: %.1f clk per 8-MMA stage) x %d SMs x the SM clock "
                           "sampled by nvidia-smi inside the timed region (%.0f MHz)" % (flop_per_clk_sm, probe, sms, sm_mhz),
                "algorithmic": {"ops_per_launch": algorithmic_ops, "tflops": algorithmic_ops / (kernel_ms * 1e-3) / 1e12,
                                "note": "SURVEY.md 8d counts sigma = 4 planes per site; the dense tetrahedron code executes 3, so executed / algorithmic = %.2f" % (planes_executed / 4.0),
                                "frac_int8_basis": algorithmic_ops / (kernel_ms * 1e-3) / 1e12 / (2.0 * bf16),
                                "int8_basis_peak": 2.0 * bf16,
                                "int8_basis_is": "cross-reference only: SURVEY.md 8d's provisional peak (2 x measured bf16) was written for a kind::i8 one-hot GEMM; "
                                                 "this kernel multiplies e2m1 nibbles (kind::mxf4, 4 x the bf16 rate) and executes 3/4 of the ops, so this quotient exceeds 1 by construction"},
                "peak_nominal_dense": 9000.0 if fp4 else 4500.0,
                "frac_of_nominal": achieved / (9000.0 if fp4 else 4500.0)}

    # ---- CPU baseline on this box's host cores (bounded sample), also the parity spot-check -----
    os.sched_setaffinity(0, all_cpus)                                # the CPU arm gets every core again
    threads = os.cpu_count() or 1
    arm = CpuArm(np.ascontiguousarray(h_seqs_np), threads)
    rows = arm.rows_for(15.0)
    dt, cpu_rows = arm.time_rows(rows)
    gpu_rows = np.asarray(h_out_np[:rows, :n]).astype(np.uint32)
    match = bool((gpu_rows == cpu_rows).all())                      # whole matrix rows: every rank's cells are in the one matrix
    t_row = dt / rows
    cpu = {"value": unique_pairs(n) / (t_row * n), "unit": "pairs/s", "cores": threads, "kind": "port",
           "sample": "first " + arm.sample_text(rows, t_row), "t_row_s": t_row,
           "rows_match_gpu": match, "cells_compared": int(rows * n)}
    seqs_host = np.ascontiguousarray(h_seqs_np)
    del arm, cpu_rows, gpu_rows
    if cleanup:
        cleanup()
    keep = None                                                       # noqa: F841 -- the pinned e2e buffers go before the file stage

    # ---- BASELINE metric, second half: snps.csv wall time, file to file, through the reference's entry point ----
    csv = None
    if not os.environ.get("BTB_BENCH_NO_CSV"):
        lib.btb_release_workspace()
        del plan, seqs
        torch.cuda.empty_cache()
        csv = csv_wall(seqs_host, world, threads)

    # kernels of this library per timed step: operand encode (+ the two per-site majority kernels of the dense
    # FP4 code) and the distance kernel
    launches_per_step = 2 + (2 if (fp4 and _query(lib, b"umma_major_zero")) else 0)
    line = {
        "metric": "pairwise SNP distances/sec", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "e2m1 x e2m1 -> fp32 (0/1 operands, exact integers)" if _query(lib, b"umma_fp4") else "int8 x int8 -> int32 (exact)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_samples": n, "n_sites": L},
        "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "snps_csv": csv,
        "csv_wall_s": (csv or {}).get("tmpfs", {}).get("csv_wall_s"), "csv_wall_s_disk": (csv or {}).get("disk", {}).get("csv_wall_s"),
        "details": {"engine": "tcgen05 %s, cta_group::%d" % ("kind::mxf4 (e2m1 0/1 operands)" if _query(lib, b"umma_fp4") else "kind::i8", 2 if pair else _query(lib, b"umma_cta_group")),
                    "options": os.environ.get("BTB_OPTIONS", ""), "tiles_total": multirank.tile_count(n), "tiles_this_rank": my_tiles,
                    "block_rows_this_rank": [row_lo, row_hi],
                    "sharding": "contiguous 256-row blocks per rank, cut so that tiles + operand encode (rows from the rank's first row on) are equal; "
                                "no collective on the compute path", "encode_cost_in_tiles": encode_tiles, "rows_encoded_this_rank": [enc_row0, n],
                    "l2_policy": "operands (2.25 GB) and output (5 GB) exceed the 126 MB L2; no flush needed",
                    "timed_region": "operand encode + distance kernels per step, CUDA events on the launching stream, max over ranks"},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(group=host_pg)
        dist.destroy_process_group()


def _ncu_traffic(n, L, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of one dist_umma_kernel launch, from the committed
    `ncu --set full` capture of this very command (profiles/umma_traffic.json); None if the capture
    was taken on another configuration."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "umma_traffic.json")))
        if t.get("n") == n and t.get("L") == L and t.get("n_gpus") == world:
            return t["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    return None


def _query(lib, key):
    v = ctypes.c_int64(0)
    lib.btb_query(key, ctypes.byref(v))
    return v.value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
