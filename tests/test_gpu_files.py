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
    # snp-dists reads gzip too (kseq over gzread)
    gz2, sc = tmp_path / "c1_snps.fas.gz", tmp_path / "snps.csv"
    gz2.write_bytes(gzip.compress(golden("c1_snps.fas")))
    ph.build_snp_matrix(str(sc), str(gz2), threads=2)
    assert sc.read_bytes() == golden("c1_snps.csv")


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


@pytest.fixture
def workers_share_devices(capi):
    """Multi-GPU level-1 calls on a box with fewer GPUs than workers: several workers per device
    (the sharding, exchange and assembly logic is the same; only the speed-up is missing)."""
    capi.check(capi.lib().btb_set_option(b"multi_gpu_oversubscribe", 1))
    yield
    capi.check(capi.lib().btb_set_option(b"multi_gpu_oversubscribe", 0))


@pytest.mark.parametrize("workers", [2, 3])
def test_multi_gpu_file_path_same_bytes(capi, oracle, tmp_path, workers_share_devices, workers):
    """n_gpus > 1 (one host thread per device, sharded upload, operand shards exchanged GPU to GPU,
    block rows per device, one host matrix) writes the same snps.csv as one GPU and as the oracle."""
    n, L = 6400, 700
    names = synth.sample_names(n)
    seqs = synth.snp_alignment(n, L, seed=6, heavy_n=0.05)
    sf = tmp_path / "snps.fas"
    synth.write_fasta(str(sf), names, seqs, wrap=0)
    outs = []
    for gpus in (1, workers):
        out = tmp_path / ("snps_%d.csv" % gpus)
        capi.check(capi.lib().btb_snp_dists(str(sf).encode(), str(out).encode(), 0, 0, 0, b",", 1, 1, gpus, 8, -1, None, None))
        outs.append(out.read_bytes())
    assert outs[0] == outs[1]
    assert outs[0] == oracle.format_matrix(names, oracle.snp_dists_rows(seqs, threads=8))
    v = ctypes.c_int64(-1)
    capi.check(capi.lib().btb_query(b"comm_last_path", ctypes.byref(v)))
    assert v.value == 1                      # the sharded path ran, not the replicated fallback


@pytest.mark.parametrize("workers", [2, 5])
def test_multi_gpu_snp_sites_same_bytes(capi, oracle, tmp_path, workers_share_devices, workers):
    """n_gpus > 1: samples are dealt to the devices, the per-device column states are OR-ed on the
    host, every device compacts its own samples.  Same snps.fas as one GPU and as the oracle."""
    n, G = 37, 250007
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 3000, seed=41, n_frac=0.02, lower=0.01)
    mf = tmp_path / "multi_fasta.fas"
    mf.write_bytes(synth.fasta_bytes(names, aln, wrap=[60, 70, 0, 31], eol=b"\n"))
    outs = []
    for gpus in (1, workers):
        out = tmp_path / ("snps_%d.fas" % gpus)
        L = ctypes.c_int64(0)
        capi.check(capi.lib().btb_snp_sites(str(mf).encode(), str(out).encode(), 1, gpus, 8, None, None, ctypes.byref(L)))
        outs.append(out.read_bytes())
    o_sf = tmp_path / "o.fas"
    assert oracle.run_snp_sites(str(mf), str(o_sf)).returncode == 0
    assert outs[0] == outs[1] == o_sf.read_bytes()


def _last_predicted(capi):
    v = ctypes.c_int64(-2)
    capi.check(capi.lib().btb_query(b"ingest_last_predicted", ctypes.byref(v)))
    return v.value


def test_predicted_framing_is_used_and_equals_exact(ph, capi, oracle, tmp_path):
    """Equal-layout records take the predicted framing (no byte scan for markers); switching it off
    gives the same bytes; so do the oracle tools."""
    n, G = 90, 30017
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 1200, seed=41, n_frac=0.02, lower=0.01)
    o_sf = tmp_path / "o.fas"
    for wrap, eol, tail in ((60, b"\n", True), (70, b"\r\n", True), (0, b"\n", True), (60, b"\n", False)):
        mf = tmp_path / "multi_fasta.fas"
        data = synth.fasta_bytes(names, aln, wrap=wrap, eol=eol, comments=["x=%d" % G, "", "a b c"])
        mf.write_bytes(data if tail else data[: -len(eol)])
        assert oracle.run_snp_sites(str(mf), str(o_sf)).returncode == 0
        out = {}
        for predict in (1, 0):
            capi.check(capi.lib().btb_set_option(b"ingest_predict", predict))
            sf = tmp_path / ("snps%d.fas" % predict)
            try:
                ph.snp_sites(str(sf), str(mf), threads=3)
            finally:
                capi.check(capi.lib().btb_set_option(b"ingest_predict", 1))
            assert _last_predicted(capi) == predict, (wrap, eol, tail)
            out[predict] = sf.read_bytes()
        assert out[1] == out[0] == o_sf.read_bytes(), (wrap, eol, tail)


def test_predicted_framing_traps_fall_back_to_the_exact_path(ph, capi, oracle, tmp_path):
    """Inputs that pass the host's boundary checks but are not what the prediction assumes: the pack
    kernel's flags catch them and the call is redone with exact framing."""
    rng = np.random.default_rng(3)
    G, W, n = 6000, 60, 12
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), G)
    rows = []
    for i in range(n):
        r = base.copy()
        r[rng.integers(0, G, 40)] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 40)
        rows.append(r.tobytes())

    def record(i, lines):
        return b">s%d\n" % i + b"".join(l + b"\n" for l in lines)

    def even(seq):
        return [seq[k:k + W] for k in range(0, G, W)]

    # (a) one record has a 61-base line followed by a 59-base line: same byte count, terminators shifted
    recs = [record(i, even(rows[i])) for i in range(n)]
    lines = even(rows[5])
    moved = lines[:10] + [lines[10] + lines[11][:1], lines[11][1:]] + lines[12:]
    recs[5] = record(5, moved)
    assert len(recs[5]) == len(recs[4])
    mf, sf, o_sf = tmp_path / "a.fa", tmp_path / "a.fas", tmp_path / "o.fas"
    mf.write_bytes(b"".join(recs))
    assert oracle.run_snp_sites(str(mf), str(o_sf)).returncode == 0
    ph.snp_sites(str(sf), str(mf))
    assert _last_predicted(capi) == 0 and sf.read_bytes() == o_sf.read_bytes()

    # (b) a line inside a record starts with '>': for the parser that is a new (short) record -> unequal lengths
    recs = [record(i, even(rows[i])) for i in range(n)]
    lines = even(rows[7])
    lines[20] = b">" + lines[20][1:]
    recs[7] = record(7, lines)
    mf.write_bytes(b"".join(recs))
    r = oracle.run_snp_sites(str(mf), str(o_sf))
    assert r.returncode != 0 and b"unequal length" in (r.stderr if isinstance(r.stderr, bytes) else r.stderr.encode())
    with pytest.raises(Exception) as e:
        ph.snp_sites(str(sf), str(mf))
    assert _last_predicted(capi) == 0 and "unequal length" in str(e.value)
    # ... unless the pieces happen to be records of the right length themselves
    half = [record(i, even(rows[i])) for i in range(4)]
    tricky = b">t\n" + b"".join(l + b"\n" for l in even(rows[4])[: G // W - 1])       # one line short ...
    tricky = tricky + b">" + b"u" * (W - 1) + b"\n"                                     # ... made up by a header-looking line
    assert len(tricky) == len(half[0]) + (len(b">t\n") - len(b">s0\n"))
    mf.write_bytes(b"".join(half) + tricky)
    r = oracle.run_snp_sites(str(mf), str(o_sf))
    with pytest.raises(Exception) as e:
        ph.snp_sites(str(sf), str(mf))
    assert r.returncode != 0 and _last_predicted(capi) == 0

    # (c) a control byte among the bases (a tab) keeps the prediction off but is plain data for the parser
    recs = [record(i, even(rows[i])) for i in range(n)]
    lines = even(rows[2])
    lines[3] = lines[3][:7] + b"\t" + lines[3][8:]
    recs[2] = record(2, lines)
    mf.write_bytes(b"".join(recs))
    assert oracle.run_snp_sites(str(mf), str(o_sf)).returncode == 0
    ph.snp_sites(str(sf), str(mf))
    assert _last_predicted(capi) == 0 and sf.read_bytes() == o_sf.read_bytes()


def test_consensus_file_lists_equal_the_concatenated_multi_fasta(ph, capi, oracle, tmp_path):
    """snp_sites() over per-sample consensus files == build_multi_fasta + snp_sites."""
    text = golden("c1_multi_fasta.fas")
    parts = [b">" + p for p in text.split(b">") if p]
    paths = []
    for i, p in enumerate(parts):
        f = tmp_path / ("s%d.fas" % i)
        f.write_bytes(p)
        paths.append(str(f))
    sf = tmp_path / "snps.fas"
    assert ph.snp_sites(str(sf), paths) == json.loads(golden("c1_meta.json"))
    assert sf.read_bytes() == golden("c1_snps.fas") and _last_predicted(capi) == 1
    # the reference's way: concatenate, then run
    mf = tmp_path / "multi_fasta.fas"
    ph.build_multi_fasta(str(mf), paths)
    assert mf.read_bytes() == text
    # gzip members and an empty file in the list; a file without its final newline in the middle
    gz = tmp_path / "s1.fas.gz"
    gz.write_bytes(gzip.compress(parts[1]))
    empty = tmp_path / "empty.fas"
    empty.write_bytes(b"")
    assert ph.snp_sites(str(sf), [paths[0], str(gz), str(empty)] + paths[2:]) == json.loads(golden("c1_meta.json"))
    assert sf.read_bytes() == golden("c1_snps.fas") and _last_predicted(capi) == 0
    cut = tmp_path / "cut.fas"
    cut.write_bytes(parts[1][:-1])
    odd = [paths[0], str(cut)] + paths[2:]
    ph.build_multi_fasta(str(mf), odd)
    o_sf = tmp_path / "o.fas"
    r = oracle.run_snp_sites(str(mf), str(o_sf))
    if r.returncode == 0:
        ph.snp_sites(str(sf), odd)
        assert sf.read_bytes() == o_sf.read_bytes()
    else:
        with pytest.raises(Exception):
            ph.snp_sites(str(sf), odd)
    with pytest.raises(Exception) as e:
        ph.snp_sites(str(sf), [paths[0], str(tmp_path / "missing.fas")])
    assert "Could not open filename" in str(e.value)


def test_path_shims_serve_the_reference_command_lines(tmp_path):
    """with tools/ first on PATH the exact command lines the reference builds
    (btbphylo/phylogeny.py:133 and :149, run through a shell like utils.run does) produce the golden
    files, and failures surface as a non-zero exit status."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PATH=os.path.join(root, "tools") + os.pathsep + os.path.dirname(sys.executable) + os.pathsep + os.environ["PATH"])
    mf, sf, sc = tmp_path / "multi_fasta.fas", tmp_path / "snps.fas", tmp_path / "snps.csv"
    mf.write_bytes(golden("c1_multi_fasta.fas"))
    r = subprocess.run("snp-sites %s -c -o %s" % (mf, sf), shell=True, env=env, capture_output=True)
    assert r.returncode == 0, r.stderr
    assert sf.read_bytes() == golden("c1_snps.fas")
    r = subprocess.run("snp-dists -c -j 4 %s > %s" % (sf, sc), shell=True, env=env, capture_output=True)
    assert r.returncode == 0, r.stderr
    assert sc.read_bytes() == golden("c1_snps.csv")
    assert b"This is snp-dists 0.8.2" in r.stderr and b"Read 4 sequences of length 831" in r.stderr
    mf.write_bytes(b">a\nACGT\n>b\nACGT\n")                     # no SNPs: snp-sites exits 1
    r = subprocess.run("snp-sites %s -c -o %s" % (mf, tmp_path / "none.fas"), shell=True, env=env, capture_output=True)
    assert r.returncode == 1 and b"No SNPs" in r.stderr and not (tmp_path / "none.fas").exists()
    mf.write_bytes(b">a\nACGT\n>b\nACG\n")                      # unequal lengths: snp-dists exits 1
    r = subprocess.run("snp-dists -c -j 1 %s > %s" % (mf, sc), shell=True, env=env, capture_output=True)
    assert r.returncode == 1 and b"has length 3 but expected 4" in r.stderr


def test_long_rows_uint32_cells_through_the_file_writer(ph, capi, oracle, tmp_path):
    """66,000-site rows need uint32 cells and print 5-digit distances: the level-1 writer (exact row
    lengths, in-place formatting) against the oracle's snp-dists; also through the file mapping."""
    n, L = 1030, 66000
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = acgt[rng.integers(0, 4, size=(n, L))]                      # unrelated rows: distances around 3 L / 4
    seqs[::3] = seqs[0]                                               # ... and some identical / close ones
    seqs[1::3, : L // 2] = seqs[1, : L // 2]
    names = synth.sample_names(n)
    sf, sc, oc = tmp_path / "snps.fas", tmp_path / "snps.csv", tmp_path / "o.csv"
    synth.write_fasta(str(sf), names, seqs, wrap=0)
    assert oracle.run_snp_dists(str(sf), str(oc), threads=8).returncode == 0
    want = oc.read_bytes()
    ph.build_snp_matrix(str(sc), str(sf), threads=4)
    assert sc.read_bytes() == want
    try:
        capi.check(capi.lib().btb_set_option(b"csv_mmap", 1))
        ph.build_snp_matrix(str(sc), str(sf), threads=3)
        assert sc.read_bytes() == want
    finally:
        capi.check(capi.lib().btb_set_option(b"csv_mmap", 0))
