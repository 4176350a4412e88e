"""CPU, world_size 2 over gloo: the N>1 path of the distance stage is sharding (tile ranges, or
block rows with equal tile counts) with no data-path collective -- each rank produces its share,
the host gathers.  Here the
per-rank tiles come from the oracle (no GPU in this container); the sharding, the gather and the
max-over-ranks timing reduction that bench.py uses are the code under test."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, L, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from btb_phylo_b200 import multirank, synth
    from oracle import oracle
    seqs = synth.snp_alignment(n, L, seed=11, heavy_n=0.1)          # operands replicated on every rank
    lo, hi = multirank.tile_share(n, rank, world)
    full = oracle.snp_dists_rows(seqs, threads=2)
    tiles = np.zeros((hi - lo, 256, 256), dtype=np.uint32)
    for t in range(lo, hi):                                          # stands in for btb_distance_tiles(out_mode=1)
        I, J = multirank.tile_coords(n, t)
        blk = full[I * 256:(I + 1) * 256, J * 256:(J + 1) * 256]
        tiles[t - lo, : blk.shape[0], : blk.shape[1]] = blk
    # host gather (what the CSV writer consumes) + bench-style max-over-ranks time
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, tiles))
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    pairs = torch.tensor([float(sum(1 for t in range(lo, hi)))], dtype=torch.float64)
    dist.all_reduce(pairs, op=dist.ReduceOp.SUM)
    if rank == 0:
        shares = [(g[0], g[1]) for g in gathered]
        mat = multirank.assemble(n, shares, [g[2] for g in gathered])
        ret["equal"] = bool((mat == full).all())
        ret["max_ms"] = float(ms.item())
        ret["tiles"] = float(pairs.item())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tile_sharding_assembles_full_matrix():
    n, L, world = 700, 150, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, L, ret), nprocs=world, join=True)
    from btb_phylo_b200 import multirank
    assert ret["equal"] and ret["max_ms"] == 11.0 and ret["tiles"] == multirank.tile_count(n)


def _row_worker(rank, world, port, n, L, ret, encode_tiles=None):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from btb_phylo_b200 import device, multirank, synth
    from oracle import oracle
    seqs = synth.snp_alignment(n, L, seed=12, heavy_n=0.05)
    if encode_tiles is None:
        lo, hi = device.row_share(n, rank, world)                    # C-ABI: btb_dist_row_share (host only)
        assert (lo, hi) == multirank.row_share(n, rank, world)
    else:                                                            # shares with the operand encode counted in
        lo, hi = device.row_share_weighted(n, rank, world, encode_tiles)
        assert (lo, hi) == multirank.row_share_weighted(n, rank, world, encode_tiles)
    full = oracle.snp_dists_rows(seqs, threads=2)
    r0, r1 = min(n, lo * 256), min(n, hi * 256)
    mine = np.zeros((n, n), dtype=np.uint32)                         # stands in for btb_distance_rows(mirror=1)
    mine[r0:r1, r0:] = full[r0:r1, r0:]
    mine[r1:, r0:r1] = full[r1:, r0:r1]
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, mine))
    if rank == 0:
        mat = multirank.assemble_rows(n, [(g[0], g[1]) for g in gathered], [g[2] for g in gathered])
        ret["equal"] = bool((mat == full).all())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_block_row_sharding_assembles_full_matrix():
    n, L, world = 1100, 90, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_row_worker, args=(world, _free_port(), n, L, ret), nprocs=world, join=True)
    assert ret["equal"]


def test_two_rank_weighted_block_rows_assemble_full_matrix():
    n, L, world = 2300, 60, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_row_worker, args=(world, _free_port(), n, L, ret, 12.5), nprocs=world, join=True)
    assert ret["equal"]


@pytest.mark.parametrize("n", [300, 6400, 50000, 100000])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("encode_tiles", [0.0, 40.0, 675.0, 1e5])
def test_weighted_row_shares_partition_the_block_rows(n, world, encode_tiles):
    """btb_dist_row_share_weighted (C) == multirank.row_share_weighted (numpy mirror); the ranges are contiguous,
    cover every block row once, and the costliest rank is no worse than with the unweighted rule."""
    sys.path.insert(0, ROOT)
    from btb_phylo_b200 import device, multirank
    T = multirank.tiles_per_dim(n)
    shares = [device.row_share_weighted(n, r, world, encode_tiles) for r in range(world)]
    assert shares == [multirank.row_share_weighted(n, r, world, encode_tiles) for r in range(world)]
    assert shares[0][0] == 0 and shares[-1][1] == T and all(shares[r][1] == shares[r + 1][0] for r in range(world - 1))

    def cost(a, b):
        return sum(T - i for i in range(a, b)) + encode_tiles * (T - a) / T
    plain = [multirank.row_share(n, r, world) for r in range(world)]
    if T >= 4 * world:
        assert max(cost(a, b) for a, b in shares) <= max(cost(a, b) for a, b in plain) + (T + 1)    # within one block row of work
