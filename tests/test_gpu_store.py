"""GPU (-m gpu): the incremental snp-sites stage (btb_phylo_b200/plane_store.py).  With a plane
store the call reads the text of new or changed consensus files only and takes the bit-planes of all others
from the store; snps.fas must be what the uncached call -- and the oracle -- write, byte for byte."""
import os
import time

import numpy as np
import pytest

from btb_phylo_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.fixture
def ph(capi):
    from btb_phylo_b200 import phylogeny
    return phylogeny


def write_samples(tmp_path, names, aln, wrap=60):
    paths = []
    for nm, row in zip(names, aln):
        p = tmp_path / (nm + ".fas")
        p.write_bytes(synth.fasta_bytes([nm], row[None, :], wrap=wrap))
        paths.append(str(p))
    return paths


def blobs_of(store):
    return sorted(f for f in os.listdir(store) if f.endswith(".planes"))


def oracle_snps(oracle, tmp_path, paths):
    mf = tmp_path / "multi.fas"
    with open(mf, "wb") as out:
        for p in paths:
            out.write(open(p, "rb").read())
    o = tmp_path / "oracle.fas"
    assert oracle.run_snp_sites(str(mf), str(o)).returncode == 0
    return o.read_bytes()


def test_weekly_growth_same_bytes_as_the_uncached_call(ph, oracle, tmp_path):
    n, G = 60, 123457
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 1500, seed=77, n_frac=0.02)
    paths = write_samples(tmp_path, names, aln)
    store = tmp_path / "store"
    out = tmp_path / "snps.fas"
    # week 1: 40 samples, empty store
    meta = ph.snp_sites(str(out), paths[:40], plane_store=str(store))
    want = oracle_snps(oracle, tmp_path, paths[:40])
    assert out.read_bytes() == want and meta["number_of_snps"] == len(want.split(b"\n")[1])
    assert len(blobs_of(store)) == 40
    # week 2: 20 more samples -- 40 come from the store.  The new samples change the kept-column set.
    before = {f: os.stat(store / f).st_mtime_ns for f in blobs_of(store)}
    meta = ph.snp_sites(str(out), paths, plane_store=str(store))
    want = oracle_snps(oracle, tmp_path, paths)
    assert out.read_bytes() == want
    assert len(blobs_of(store)) == 60 and all(os.stat(store / f).st_mtime_ns == t for f, t in before.items())
    plain = tmp_path / "plain.fas"
    ph.snp_sites(str(plain), paths)
    assert plain.read_bytes() == want
    # a sample is re-sequenced: its file changes (size or mtime) and is read again; another order of the files
    aln[7, 1000:1100] = ord("N")
    time.sleep(0.01)
    open(paths[7], "wb").write(synth.fasta_bytes([names[7]], aln[7][None, :], wrap=60))
    order = list(reversed(range(n)))
    meta = ph.snp_sites(str(out), [paths[i] for i in order], plane_store=str(store))
    assert out.read_bytes() == oracle_snps(oracle, tmp_path, [paths[i] for i in order])
    # everything from the store
    meta2 = ph.snp_sites(str(out), [paths[i] for i in order], plane_store=str(store))
    assert meta2 == meta and out.read_bytes() == oracle_snps(oracle, tmp_path, [paths[i] for i in order])


def test_store_is_ignored_for_inputs_that_need_the_exact_parser(ph, oracle, tmp_path):
    n, G = 12, 5003
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 300, seed=78, n_frac=0.0)
    paths = write_samples(tmp_path, names, aln)
    # one file with two wrap widths inside its record (irregular) and one with two records
    open(paths[3], "wb").write(b">" + names[3].encode() + b"\n" + aln[3][:100].tobytes() + b"\n" + aln[3][100:].tobytes() + b"\n")
    both = synth.fasta_bytes(names[5:7], aln[5:7], wrap=60)
    open(paths[5], "wb").write(both)
    files = paths[:6] + paths[7:]
    out = tmp_path / "snps.fas"
    ph.snp_sites(str(out), files, plane_store=str(tmp_path / "store"))
    assert out.read_bytes() == oracle_snps(oracle, tmp_path, files)


def test_no_snps_raises_like_the_tool(ph, tmp_path):
    names = synth.sample_names(3)
    aln = np.repeat(np.frombuffer(b"ACGT" * 50, dtype=np.uint8)[None, :], 3, axis=0)
    paths = write_samples(tmp_path, names, aln, wrap=0)
    with pytest.raises(Exception, match="No SNPs were detected"):
        ph.snp_sites(str(tmp_path / "snps.fas"), paths, plane_store=str(tmp_path / "store"))
    assert not (tmp_path / "snps.fas").exists()


def test_damaged_or_foreign_blobs_are_packed_again(ph, oracle, tmp_path):
    n, G = 16, 40003
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 700, seed=79, n_frac=0.01)
    paths = write_samples(tmp_path, names, aln)
    store = tmp_path / "store"
    out = tmp_path / "snps.fas"
    ph.snp_sites(str(out), paths, plane_store=str(store))
    want = oracle_snps(oracle, tmp_path, paths)
    assert out.read_bytes() == want
    blobs = blobs_of(store)
    assert len(blobs) == n
    raw = (store / blobs[0]).read_bytes()
    (store / blobs[0]).write_bytes(raw[:-8])                          # truncated
    (store / blobs[1]).write_bytes(b"BTBPLN01" + raw[8:])             # another format version
    (store / blobs[2]).write_bytes(b"")                               # empty
    flipped = bytearray((store / blobs[3]).read_bytes())              # size and header intact, source "changed" (mtime field)
    flipped[32] ^= 0xFF
    (store / blobs[3]).write_bytes(bytes(flipped))
    ph.snp_sites(str(out), paths, plane_store=str(store))
    assert out.read_bytes() == want
    assert all(len((store / b).read_bytes()) == len(raw) for b in blobs)   # all four were written again


def test_crlf_files_and_a_flagged_record(ph, oracle, tmp_path):
    n, G = 10, 9001
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 400, seed=80, n_frac=0.03)

    def write(nm, row, wrap=70, crlf=True):
        p = tmp_path / (nm + ".fas")
        text = synth.fasta_bytes([nm], row[None, :], wrap=wrap)
        p.write_bytes(text.replace(b"\n", b"\r\n") if crlf else text)
        return str(p)
    paths = [write(nm, row) for nm, row in zip(names, aln)]
    store = tmp_path / "store"
    out = tmp_path / "snps.fas"
    ph.snp_sites(str(out), paths[:7], plane_store=str(store))
    assert out.read_bytes() == oracle_snps(oracle, tmp_path, paths[:7])
    assert len(blobs_of(store)) == 7
    ph.snp_sites(str(out), paths, plane_store=str(store))
    assert out.read_bytes() == oracle_snps(oracle, tmp_path, paths)
    assert len(blobs_of(store)) == n
    # a new file with a tab among its bases: the kernel flags the record, the call goes through the exact parser
    extra = aln[0].copy()
    extra[1234] = 9
    files = paths + [write("tabbed", extra)]
    ph.snp_sites(str(out), files, plane_store=str(store))
    assert out.read_bytes() == oracle_snps(oracle, tmp_path, files)
    assert len(blobs_of(store)) == n                                # nothing was stored for the fallback call
    # new files of another layout than the first new one (other wrap, LF): not what the predicted framing takes
    # -- same bytes through the exact parser, whatever the store holds
    more = [write("w60", aln[1], wrap=60, crlf=False), write("w0", aln[2], wrap=0, crlf=False), write("w80", aln[3], wrap=80)]
    files = paths + more
    ph.snp_sites(str(out), files, plane_store=str(store))
    assert out.read_bytes() == oracle_snps(oracle, tmp_path, files)


def test_a_file_listed_twice_and_a_name_that_is_not_utf8(ph, oracle, tmp_path):
    n, G = 6, 7001
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 200, seed=81, n_frac=0.0)
    paths = write_samples(tmp_path, names, aln)
    odd = tmp_path / "odd.fas"
    odd.write_bytes(synth.fasta_bytes(["x"], aln[0][None, :], wrap=60).replace(b">x\n", b">s\xff\xfe 1\n", 1))
    files = paths + [str(odd), paths[2]]
    store = tmp_path / "store"
    out = tmp_path / "snps.fas"
    for _ in range(2):                                                 # empty store, then everything from it
        ph.snp_sites(str(out), files, plane_store=str(store))
        assert out.read_bytes() == oracle_snps(oracle, tmp_path, files)
    assert len(blobs_of(store)) == n + 1
