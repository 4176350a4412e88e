"""CPU: the C-ABI library loads, exports every symbol include/btb_phylo.h declares, fails loudly
without a GPU (no CPU fallback), and its host-only pieces (tile order, CSV formatting) are right."""
import ctypes
import os
import re

import numpy as np
import pytest

from btb_phylo_b200 import multirank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "btb_phylo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(btb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(capi):
    lib = ctypes.CDLL(capi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "libbtbphylo.so does not export " + s
    assert set(syms) == set(capi.SIGNATURES), "ctypes binding and header disagree"


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "btb_phylo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("no oracle", ""), f + " mentions the oracle"


def test_no_gpu_means_loud_failure(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from btb_phylo_b200 import device, phylogeny
    with pytest.raises(capi.BtbError) as e:
        device.distance_matrix_host(np.zeros((4, 8), dtype=np.uint8))
    assert e.value.code == capi.ENODEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(Exception) as e2:
        phylogeny.build_snp_matrix("/tmp/never.csv", os.path.join(ROOT, "tests/golden/kat3_snps.fas"), threads="4")
    assert "snp-dists -c -j 4" in str(e2.value) and "exit code 1" in str(e2.value)
    with pytest.raises(Exception) as e3:
        phylogeny.snp_sites("/tmp/never.fas", os.path.join(ROOT, "tests/golden/kat3_multi_fasta.fas"))
    assert "snp-sites" in str(e3.value) and not os.path.exists("/tmp/never.fas")


@pytest.mark.parametrize("n", [1, 255, 256, 257, 2049, 5000])
def test_tile_order_covers_upper_triangle_once(capi, n):
    from btb_phylo_b200 import device
    T = (n + 255) // 256
    cnt = device.tile_count(n)
    assert cnt == T * (T + 1) // 2 == multirank.tile_count(n)
    seen = [device.tile_coords(n, t) for t in range(cnt)]
    assert len(set(seen)) == cnt and all(0 <= i <= j < T for i, j in seen)
    assert seen == [multirank.tile_coords(n, t) for t in range(cnt)]
    # band locality: any 74 consecutive tiles touch few distinct panels
    if cnt > 74:
        worst = max(len({i for i, _ in seen[k:k + 74]}) + len({j for _, j in seen[k:k + 74]}) for k in range(0, cnt - 74, 37))
        assert worst <= 40


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_tile_shares_partition(world):
    n = 50000
    shares = [multirank.tile_share(n, r, world) for r in range(world)]
    assert shares[0][0] == 0 and shares[-1][1] == multirank.tile_count(n)
    assert all(a[1] == b[0] for a, b in zip(shares, shares[1:]))
    sizes = [hi - lo for lo, hi in shares]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n", [1, 256, 257, 5000, 50000])
@pytest.mark.parametrize("world", [1, 2, 3, 8, 300])
def test_row_shares_partition_block_rows(capi, n, world):
    from btb_phylo_b200 import device, multirank
    T = device.block_rows(n)
    assert T == (n + 255) // 256
    total = multirank.tile_count(n)
    prev, counts = 0, []
    for r in range(world):
        lo, hi = device.row_share(n, r, world)
        assert (lo, hi) == multirank.row_share(n, r, world)
        assert lo == prev and hi >= lo
        counts.append(sum(T - i for i in range(lo, hi)))
        prev = hi
    assert prev == T and sum(counts) == total
    if world <= 8 and T >= 16 * world:                               # balanced within one block row of tiles
        assert max(counts) - min(counts) <= 2 * T


@pytest.mark.parametrize("elem", [2, 4])
def test_format_csv_rows_matches_snp_dists_layout(capi, oracle, elem):
    rng = np.random.default_rng(5)
    n = 301
    dt = np.uint16 if elem == 2 else np.uint32
    hi = 65535 if elem == 2 else 4_000_000_000
    mat = rng.integers(0, hi, size=(n, n), dtype=np.uint64).astype(dt)
    mat[0, :12] = [0, 9, 10, 99, 100, 999, 1000, 9999, 10000, 65535 if elem == 2 else 99999999, 7, hi - 1]
    names = ["AF-12-%05d-22_consensus" % i for i in range(n)]
    arr = (ctypes.c_char_p * n)(*[s.encode() for s in names])
    buf = ctypes.create_string_buffer(n * (n * 11 + 64))
    w = ctypes.c_size_t(0)
    for threads in (1, 4):
        for sep in (b",", b"\t"):
            capi.check(capi.lib().btb_format_csv_rows(mat.ctypes.data, elem, n, n, 0, n, arr, sep, threads, buf, len(buf), ctypes.byref(w)))
            want = oracle.format_matrix(names, mat, sep=sep.decode()).split(b"\n", 1)[1]
            assert buf.raw[: w.value] == want
    # row window
    capi.check(capi.lib().btb_format_csv_rows(mat[10:].ctypes.data, elem, n, n, 10, 20, arr, b",", 2, buf, len(buf), ctypes.byref(w)))
    want = b"".join(oracle.format_matrix(names, mat).split(b"\n")[11:21][k] + b"\n" for k in range(10))
    assert buf.raw[: w.value] == want


def test_assemble_from_packed_tiles(oracle):
    from btb_phylo_b200 import synth
    n = 600
    seqs = synth.snp_alignment(n, 200, seed=3, heavy_n=0.1)
    full = oracle.snp_dists_rows(seqs, threads=4)
    shares = [multirank.tile_share(n, r, 3) for r in range(3)]
    packed = []
    for lo, hi in shares:
        tiles = np.zeros((hi - lo, 256, 256), dtype=np.uint32)
        for t in range(lo, hi):
            I, J = multirank.tile_coords(n, t)
            blk = full[I * 256:(I + 1) * 256, J * 256:(J + 1) * 256]
            tiles[t - lo, : blk.shape[0], : blk.shape[1]] = blk
        packed.append(tiles)
    assert (multirank.assemble(n, shares, packed) == full).all()
