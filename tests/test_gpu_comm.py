"""GPU (-m gpu): the multi-rank distance stage with a sharded upload (btb_comm, csrc/dist_comm.cu):
one process per rank, operand shards exchanged GPU to GPU through CUDA IPC, one shared host matrix.
The ranks are spread over the GPUs the box has -- on a one-GPU box they share device 0 (CUDA IPC
between two processes works on one device as well), so the exchange protocol is exercised
everywhere; the in-process (thread per device) form is covered through btb_snp_dists(n_gpus > 1)."""
import os
import subprocess
import sys
import uuid

import numpy as np
import pytest
import torch

from btb_phylo_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "comm_worker.py")


def run_ranks(seqs, world, flags="", calls=1, tmp_path=None):
    n = seqs.shape[0]
    elem = 4 if "4" in flags else 2
    tag = uuid.uuid4().hex[:12]
    out_path = "/dev/shm/btbphylo_test_%s.mat" % tag
    seqs_path = str(tmp_path / "seqs.npy")
    np.save(seqs_path, np.ascontiguousarray(seqs))
    with open(out_path, "wb") as f:
        f.truncate(n * n * elem)
    n_dev = torch.cuda.device_count()
    try:
        procs = [subprocess.Popen([sys.executable, WORKER, "btbphylo_test_" + tag, str(r), str(world), str(r % n_dev),
                                   seqs_path, out_path, flags or "-", str(calls)],
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                                  env=dict(os.environ, BTB_COMM_TIMEOUT_S="120"))
                 for r in range(world)]
        outs = [p.communicate(timeout=600) for p in procs]
        for p, (so, se) in zip(procs, outs):
            assert p.returncode == 0, "rank failed:\n%s\n%s" % (so, se)
        paths = [int(so.split("path")[1].split()[0]) for so, _ in outs]
        mat = np.array(np.memmap(out_path, dtype=np.uint32 if elem == 4 else np.uint16, mode="r", shape=(n, n)))
    finally:
        os.unlink(out_path)
    return mat.astype(np.int64), paths


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_upload_dense_alignment(capi, oracle, tmp_path, world):
    seqs = synth.snp_alignment(2300, 1500, seed=31)                    # ACGT only: the dense tetrahedron code + row sums
    got, paths = run_ranks(seqs, world, calls=2, tmp_path=tmp_path)   # two epochs: buffers are reused
    assert paths == [1] * world
    assert (got == oracle.snp_dists_rows(seqs, threads=8).astype(np.int64)).all()


def test_sharded_upload_ignored_bytes_in_one_shard_only(capi, oracle, tmp_path):
    # N only in the LAST rows: rank 0's shard alone looks dense, the OR of the presence maps must decide
    seqs = synth.snp_alignment(1800, 900, seed=32)
    seqs[1700:, 100:300] = ord("N")
    seqs[1750, 5] = ord("-")
    got, paths = run_ranks(seqs, 2, tmp_path=tmp_path)
    assert paths == [1, 1]
    assert (got == oracle.snp_dists_rows(seqs, threads=8).astype(np.int64)).all()


def test_sharded_upload_allchars_alphabet(capi, oracle, tmp_path):
    seqs = synth.snp_alignment(1500, 1100, seed=33, heavy_n=0.2, iupac=0.01, lower=0.02)
    seqs[7, 10:40] = ord(".")
    got, paths = run_ranks(seqs, 2, flags="a", tmp_path=tmp_path)
    assert paths == [1, 1]
    assert (got == oracle.snp_dists_rows(seqs, allchars=True, threads=8).astype(np.int64)).all()


def test_popc_engine_takes_the_replicated_fallback(capi, oracle, tmp_path):
    seqs = synth.snp_alignment(900, 400, seed=34, heavy_n=0.1)
    got, paths = run_ranks(seqs, 2, flags="p", tmp_path=tmp_path)
    assert paths == [0, 0]
    assert (got == oracle.snp_dists_rows(seqs, threads=8).astype(np.int64)).all()


def test_cta_pair_kernel_shares_on_four_ranks(capi, oracle, tmp_path):
    # enough rows for the 512 x 240 CTA-pair units in every rank's share
    seqs = synth.snp_alignment(12800, 256, seed=35)
    got, paths = run_ranks(seqs, 4, tmp_path=tmp_path)
    assert paths == [1] * 4
    want = oracle.snp_dists_rows(seqs, threads=os.cpu_count() or 8).astype(np.int64)
    assert (got == want).all()


def test_a_failed_rank_does_not_hang_the_others(capi, tmp_path):
    # rank 1 never shows up: rank 0 must give up with BTB_ECOMM instead of spinning forever
    seqs = synth.snp_alignment(300, 64, seed=36)
    np.save(str(tmp_path / "s.npy"), seqs)
    out = "/dev/shm/btbphylo_test_%s.mat" % uuid.uuid4().hex[:12]
    with open(out, "wb") as f:
        f.truncate(300 * 300 * 2)
    try:
        p = subprocess.run([sys.executable, WORKER, "btbphylo_test_alone_" + uuid.uuid4().hex[:8], "1", "2", "0",
                            str(tmp_path / "s.npy"), out, "-", "1"], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, BTB_COMM_TIMEOUT_S="3"))
    finally:
        os.unlink(out)
    assert p.returncode != 0 and "btb_comm_create" in p.stderr
