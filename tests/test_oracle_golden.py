"""CPU: the oracle against every golden vector held for this path (tests/golden/, produced by
make_golden.py) and against an independent numpy restatement.  The reference's own tests pin
nothing for snp_sites()/build_snp_matrix() (tests/phylogeny_test.py never calls them), so the
literals here -- upstream README examples and hand-derived edge cases -- are the pin."""
import gzip
import json
import os

import numpy as np
import pytest

from btb_phylo_b200 import synth
from tests.conftest import golden


def run_sites(oracle, tmp_path, data, pure=True, name="in.fa"):
    p = tmp_path / name
    p.write_bytes(data)
    out = tmp_path / "out.fas"
    r = oracle.run_snp_sites(str(p), str(out), pure=pure)
    return r, (out.read_bytes() if out.exists() else None)


def run_dists(oracle, tmp_path, data, flags=("-c",), threads=2, name="in.fa"):
    p = tmp_path / name
    p.write_bytes(data)
    out = tmp_path / "out.txt"
    r = oracle.run_snp_dists(str(p), str(out), threads=threads, flags=flags)
    return r, out.read_bytes()


def test_kat1_snp_dists_readme(oracle, tmp_path):
    r, out = run_dists(oracle, tmp_path, golden("kat1_in.fa"), flags=())
    assert r.returncode == 0 and out == golden("kat1_expected.tsv")
    assert r.stderr == b"This is snp-dists 0.8.2\nWill use 2 threads.\nRead 4 sequences of length 8\n"
    r, out = run_dists(oracle, tmp_path, golden("kat1_in.fa"), flags=("-c", "-m"))
    assert out == golden("kat1_expected_molten.csv")
    r, out = run_dists(oracle, tmp_path, golden("kat1_in.fa"), flags=("-b", "-q"))
    assert out == golden("kat1_expected.tsv").replace(b"snp-dists 0.8.2", b"", 1) and r.stderr == b""


def test_kat2_snp_sites_readme(oracle, tmp_path):
    r, out = run_sites(oracle, tmp_path, golden("kat2_in.fa"), pure=False)
    assert r.returncode == 0 and out == golden("kat2_expected_default.fas")
    r, out = run_sites(oracle, tmp_path, golden("kat2_in.fa"), pure=True)
    assert r.returncode == 0 and out == golden("kat2_expected_pure.fas")


@pytest.mark.parametrize("tag", ["kat3", "c1"])
def test_boundary_fixtures(oracle, tmp_path, tag):
    """multi_fasta -> snps.fas -> snps.csv as recorded through the unmodified reference functions."""
    r, fas = run_sites(oracle, tmp_path, golden(tag + "_multi_fasta.fas"))
    assert r.returncode == 0 and fas == golden(tag + "_snps.fas")
    r, csv = run_dists(oracle, tmp_path, fas, flags=("-c",), threads=4)
    assert r.returncode == 0 and csv == golden(tag + "_snps.csv")
    meta = json.loads(golden(tag + "_meta.json"))
    assert meta["number_of_snps"] == len(fas.split(b"\n")[1])


def test_edge_modes_snp_dists(oracle, tmp_path):
    """default / -a / -k / -a -k on lower case, N, '-', literal '.', IUPAC (derived by hand in make_golden.py)."""
    exp = json.loads(golden("edge_dists_expected.json"))
    for mode, flags in (("default", ("-c",)), ("a", ("-c", "-a")), ("k", ("-c", "-k")), ("ak", ("-c", "-a", "-k"))):
        r, out = run_dists(oracle, tmp_path, golden("edge_dists_in.fa"), flags=flags)
        assert out == oracle.format_matrix(["s1", "s2", "s3"], exp[mode]), mode


def test_edge_snp_sites_crlf_wrap_lowercase(oracle, tmp_path):
    r, out = run_sites(oracle, tmp_path, golden("edge_sites_in.fa"), pure=True)
    assert out == golden("edge_sites_expected_pure.fas")
    r, out = run_sites(oracle, tmp_path, golden("edge_sites_in.fa"), pure=False)
    assert out == golden("edge_sites_expected_default.fas")


def test_error_cases(oracle, tmp_path):
    r, out = run_sites(oracle, tmp_path, b">a\nACGT\n>b\nACGT\n")
    assert r.returncode == 1 and out is None and b"No SNPs were detected" in r.stderr
    r, out = run_sites(oracle, tmp_path, b">a\nACGT\n>b\nACG\n")
    assert r.returncode == 1 and b"unequal length. Expected length is 4 but got 3 in sequence b" in r.stderr
    r, out = run_dists(oracle, tmp_path, b">a\nACGT\n>b\nACG\n")
    assert r.returncode == 1 and b"ERROR: sequence #2 'b' has length 3 but expected 4" in r.stderr
    r, out = run_dists(oracle, tmp_path, b"")
    assert r.returncode == 1 and b"file contained no sequences" in r.stderr
    r = oracle.run_snp_dists(str(tmp_path / "missing.fa"), str(tmp_path / "o"), flags=())
    assert r.returncode == 1 and b"Could not open filename" in r.stderr


def test_max_seq_boundary(oracle, tmp_path):
    ok = b"".join(b">s%d\nA\n" % i for i in range(100000))
    r, out = run_dists(oracle, tmp_path, ok[: ok.find(b">s1000\n")])      # cheap sanity run
    assert r.returncode == 0
    p = tmp_path / "big.fa"
    p.write_bytes(ok + b">extra\nA\n")
    import subprocess
    r = subprocess.run([oracle.SNP_DISTS, "-q", str(p)], capture_output=True)
    assert r.returncode == 1 and b"can only handle 100000 sequences" in r.stderr


def test_gzip_input(oracle, tmp_path):
    r, out = run_dists(oracle, tmp_path, gzip.compress(golden("kat1_in.fa")), flags=(), name="in.fa.gz")
    assert out == golden("kat1_expected.tsv")


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_c_oracle_matches_numpy_restatement(oracle, seed):
    seqs = synth.snp_alignment(37, 501, seed=seed, heavy_n=0.2, iupac=0.02, lower=0.05)
    seqs[3, 10:20] = ord(".")
    for allchars in (False, True):
        for keepcase in (False, True):
            a = oracle.snp_dists_rows(seqs, allchars=allchars, keepcase=keepcase, threads=2)
            b = oracle.np_snp_dists(seqs, allchars=allchars, keepcase=keepcase)
            assert (a == b).all()
    aln = synth.genome_alignment(23, 3000, 200, seed=seed, n_frac=0.05, lower=0.02)
    for pure in (True, False):
        assert (oracle.snp_sites_mask(aln, pure=pure) == oracle.np_snp_sites_mask(aln, pure=pure)).all()


def test_oracle_file_and_memory_forms_agree(oracle, tmp_path):
    names = synth.sample_names(12)
    aln = synth.genome_alignment(12, 5000, 300, seed=9, n_frac=0.03)
    for wrap, eol in ((60, b"\n"), (70, b"\r\n"), (0, b"\n"), ([60, 80, 33], b"\n")):
        r, fas = run_sites(oracle, tmp_path, synth.fasta_bytes(names, aln, wrap=wrap, eol=eol))
        keep = oracle.snp_sites_mask(aln)
        want = synth.fasta_bytes(names, aln[:, keep], wrap=0)
        assert r.returncode == 0 and fas == want
    r, csv = run_dists(oracle, tmp_path, fas, flags=("-c",))
    assert csv == oracle.format_matrix(names, oracle.snp_dists_rows(aln[:, keep]))


def test_oracle_matches_third_party_hamming_on_dense_alignments(oracle):
    """An implementation nobody here wrote: on an alignment of A/C/G/T only, snp-dists' distance (SURVEY.md App. A.3:
    positions where both symbols count and differ) is the plain Hamming distance -- scipy's pdist x L.  With N, gaps
    and lower case the tools' rule is Hamming over the positions where both rows hold one of ACGT (after upper-casing),
    which scipy gives per pair on the masked columns."""
    from scipy.spatial.distance import hamming, pdist, squareform
    n, L = 41, 777
    dense = synth.snp_alignment(n, L, seed=11)
    assert set(np.unique(dense)) <= set(b"ACGT")
    want = np.rint(squareform(pdist(dense, metric="hamming")) * L).astype(np.int64)
    assert (oracle.snp_dists_rows(dense, threads=2) == want).all()
    mixed = synth.snp_alignment(19, 400, seed=12, heavy_n=0.15, iupac=0.03, lower=0.1)
    up = mixed.copy()
    up[(up >= 97) & (up <= 122)] -= 32
    counts = np.isin(up, np.frombuffer(b"ACGT", dtype=np.uint8))
    got = oracle.snp_dists_rows(mixed, threads=2)
    for i in range(len(mixed)):
        for j in range(len(mixed)):
            both = counts[i] & counts[j]
            d = int(round(hamming(up[i][both], up[j][both]) * both.sum())) if both.any() else 0
            assert got[i, j] == d, (i, j)
