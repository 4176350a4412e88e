"""GPU (-m gpu): the reference-facing entry points (same names / argument order as
btbphylo/phylogeny.py:128-150) produce byte-identical snps.fas / snps.csv: against the committed
golden fixtures (recorded through the unmodified reference functions) and against the oracle on
fresh synthetic inputs, including the edge cases the tools define (CRLF, wraps, lower case, gzip,
no SNPs, unequal lengths)."""
import ctypes
import gzip
import json
import os

import numpy as np
import pytest

from btb_phylo_b200 import synth
from tests.conftest import golden

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.fixture()
def ph(capi):
    from btb_phylo_b200 import phylogeny
    return phylogeny


@pytest.mark.parametrize("tag", ["kat3", "c1"])
def test_golden_pipeline(ph, tmp_path, tag):
    mf = tmp_path / "multi_fasta.fas"
    mf.write_bytes(golden(tag + "_multi_fasta.fas"))
    sf, sc = tmp_path / "snps.fas", tmp_path / "snps.csv"
    meta = ph.snp_sites(str(sf), str(mf))
    ph.build_snp_matrix(str(sc), str(sf), threads="4")
    assert meta == json.loads(golden(tag + "_meta.json"))
    assert sf.read_bytes() == golden(tag + "_snps.fas")
    assert sc.read_bytes() == golden(tag + "_snps.csv")


@pytest.mark.parametrize("tag", ["kat3", "c1", "labels"])
def test_fused_labels_equal_reference_post_processing(ph, tmp_path, tag):
    """build_snp_matrix(post_process=True) leaves the file that build_snp_matrix + the reference's
    post_process_snps_csv (phylogeny.py:172-187; golden recorded through it) would."""
    sf, sc = tmp_path / "snps.fas", tmp_path / "snps.csv"
    sf.write_bytes(golden(tag + "_snps.fas"))
    ph.build_snp_matrix(str(sc), str(sf), threads=2, post_process=True)
    assert sc.read_bytes() == golden(tag + "_snps_postprocessed.csv")
    ph.build_snp_matrix(str(sc), str(sf), threads=2)
    assert sc.read_bytes() == golden(tag + "_snps.csv")
    ph.post_process_snps_csv(str(sc))
    assert sc.read_bytes() == golden(tag + "_snps_postprocessed.csv")


def test_kat1_and_edge_modes_through_capi(capi, oracle, tmp_path):
    lib = capi.lib()
    src = tmp_path / "in.fa"
    out = tmp_path / "out.txt"
    src.write_bytes(golden("kat1_in.fa"))
    capi.check(lib.btb_snp_dists(str(src).encode(), str(out).encode(), 0, 0, 0, b"\t", 1, 1, 1, 2, -1, None, None))
    assert out.read_bytes() == golden("kat1_expected.tsv")
    capi.check(lib.btb_snp_dists(str(src).encode(), str(out).encode(), 0, 0, 1, b",", 1, 1, 1, 2, -1, None, None))
    assert out.read_bytes() == golden("kat1_expected_molten.csv")
    capi.check(lib.btb_snp_dists(str(src).encode(), str(out).encode(), 0, 0, 0, b"\t", 0, 1, 1, 1, -1, None, None))
    assert out.read_bytes() == golden("kat1_expected.tsv").replace(b"snp-dists 0.8.2", b"", 1)
    exp = json.loads(golden("edge_dists_expected.json"))
    src.write_bytes(golden("edge_dists_in.fa"))
    for mode, (a, k) in {"default": (0, 0), "a": (1, 0), "k": (0, 1), "ak": (1, 1)}.items():
        for engine in (0, 1):
            capi.check(lib.btb_snp_dists(str(src).encode(), str(out).encode(), a, k, 0, b",", 1, 1, 1, 1, engine, None, None))
            assert out.read_bytes() == oracle.format_matrix(["s1", "s2", "s3"], exp[mode]), (mode, engine)


def test_edge_snp_sites_and_gzip(ph, tmp_path):
    mf = tmp_path / "edge.fa"
    mf.write_bytes(golden("edge_sites_in.fa"))
    sf = tmp_path / "snps.fas"
    assert ph.snp_sites(str(sf), str(mf)) == {"number_of_snps": 2}
    assert sf.read_bytes() == golden("edge_sites_expected_pure.fas")
    gz = tmp_path / "c1.fa.gz"
    gz.write_bytes(gzip.compress(golden("c1_multi_fasta.fas")))
    assert ph.snp_sites(str(sf), str(gz)) == {"number_of_snps": 831}
    assert sf.read_bytes() == golden("c1_snps.fas")


def test_errors_raise_like_the_reference(ph, tmp_path):
    mf = tmp_path / "m.fa"
    sf = tmp_path / "s.fas"
    mf.write_bytes(b">a\nACGT\n>b\nACGT\n")
    with pytest.raises(Exception) as e:
        ph.snp_sites(str(sf), str(mf))
    assert "exit code 1" in str(e.value) and "No SNPs" in str(e.value) and not sf.exists()
    mf.write_bytes(b">a\nACGT\n>b\nACG\n")
    with pytest.raises(Exception) as e:
        ph.snp_sites(str(sf), str(mf))
    assert "unequal length. Expected length is 4 but got 3 in sequence b" in str(e.value)
    with pytest.raises(Exception) as e:
        ph.build_snp_matrix(str(tmp_path / "o.csv"), str(mf), 2)
    assert "sequence #2 'b' has length 3 but expected 4" in str(e.value) and not (tmp_path / "o.csv").exists()
    with pytest.raises(Exception) as e:
        ph.build_snp_matrix(str(tmp_path / "o.csv"), str(tmp_path / "missing.fas"))
    assert "Could not open filename" in str(e.value)
    mf.write_bytes(b"")
    with pytest.raises(Exception) as e:
        ph.build_snp_matrix(str(tmp_path / "o.csv"), str(mf))
    assert "no sequences" in str(e.value)


@pytest.mark.parametrize("wrap,eol", [(60, b"\n"), (70, b"\r\n"), (0, b"\n"), ([60, 80, 33, 7, 0], b"\n")])
def test_synthetic_pipeline_matches_oracle_bytes(ph, oracle, tmp_path, wrap, eol):
    n, G = 150, 60011
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 2500, seed=31, n_frac=0.02, lower=0.01)
    mf = tmp_path / "multi_fasta.fas"
    mf.write_bytes(synth.fasta_bytes(names, aln, wrap=wrap, eol=eol, comments=["len=%d" % G, ""]))
    sf, sc = tmp_path / "snps.fas", tmp_path / "snps.csv"
    meta = ph.snp_sites(str(sf), str(mf), threads=4)
    ph.build_snp_matrix(str(sc), str(sf), threads=4)
    o_sf, o_sc = tmp_path / "o.fas", tmp_path / "o.csv"
    assert oracle.run_snp_sites(str(mf), str(o_sf)).returncode == 0
    assert oracle.run_snp_dists(str(o_sf), str(o_sc), threads=8).returncode == 0
    assert sf.read_bytes() == o_sf.read_bytes()
    assert sc.read_bytes() == o_sc.read_bytes()
    assert meta["number_of_snps"] == len(o_sf.read_bytes().split(b"\n")[1])


def test_irregular_layouts_take_the_exact_walk(ph, oracle, tmp_path):
    """blank lines inside a record, a short line in the middle, '@' records, trailing junk lines"""
    data = (b"junk before\n>r1 first\nACGTAC\n\nGTAC\nGG\n>r2\nACGTACGTACGG\n@r3 fastq-style-marker\nACGAACGTAC\nTG\n"
            b">r4\r\nACGTAC\r\nGTACGG\r\n\r\n")
    mf = tmp_path / "m.fa"
    mf.write_bytes(data)
    sf, o_sf = tmp_path / "s.fas", tmp_path / "o.fas"
    r = oracle.run_snp_sites(str(mf), str(o_sf))
    assert r.returncode == 0
    ph.snp_sites(str(sf), str(mf))
    assert sf.read_bytes() == o_sf.read_bytes()


def test_full_genome_width_property(ph, oracle, tmp_path):
    """M. bovis genome length (4,349,904 columns), 24 samples: byte-identical with the oracle."""
    n, G = 24, 4349904
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 3000, seed=2, n_frac=0.01)
    mf = tmp_path / "multi_fasta.fas"
    synth.write_fasta(str(mf), names, aln, wrap=60)
    sf, o_sf = tmp_path / "snps.fas", tmp_path / "o.fas"
    ph.snp_sites(str(sf), str(mf), threads=8)
    assert oracle.run_snp_sites(str(mf), str(o_sf)).returncode == 0
    assert sf.read_bytes() == o_sf.read_bytes()


def test_multi_gpu_file_path_same_bytes(capi, oracle, tmp_path):
    """n_gpus = 2 (bands of block rows dealt to the devices, host assembly) writes the same snps.csv
    as one GPU and as the oracle.  Needs two visible devices (gpurun --gpus 2)."""
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n, L = 4500, 700
    names = synth.sample_names(n)
    seqs = synth.snp_alignment(n, L, seed=6, heavy_n=0.05)
    sf = tmp_path / "snps.fas"
    synth.write_fasta(str(sf), names, seqs, wrap=0)
    outs = []
    for gpus in (1, 2):
        out = tmp_path / ("snps_%d.csv" % gpus)
        capi.check(capi.lib().btb_snp_dists(str(sf).encode(), str(out).encode(), 0, 0, 0, b",", 1, 1, gpus, 8, -1, None, None))
        outs.append(out.read_bytes())
    assert outs[0] == outs[1]
    assert outs[0] == oracle.format_matrix(names, oracle.snp_dists_rows(seqs, threads=8))


def test_multi_gpu_snp_sites_same_bytes(capi, oracle, tmp_path):
    """n_gpus = 2: samples are dealt to the devices, the per-device column states are OR-ed on the
    host, every device compacts its own samples.  Same snps.fas as one GPU and as the oracle."""
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n, G = 37, 250007
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 3000, seed=41, n_frac=0.02, lower=0.01)
    mf = tmp_path / "multi_fasta.fas"
    mf.write_bytes(synth.fasta_bytes(names, aln, wrap=[60, 70, 0, 31], eol=b"\n"))
    outs = []
    for gpus in (1, 2):
        out = tmp_path / ("snps_%d.fas" % gpus)
        L = ctypes.c_int64(0)
        capi.check(capi.lib().btb_snp_sites(str(mf).encode(), str(out).encode(), 1, gpus, 8, None, None, ctypes.byref(L)))
        outs.append(out.read_bytes())
    o_sf = tmp_path / "o.fas"
    assert oracle.run_snp_sites(str(mf), str(o_sf)).returncode == 0
    assert outs[0] == outs[1] == o_sf.read_bytes()
