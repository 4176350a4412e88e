#!/usr/bin/env python
"""Regenerates tests/golden/*.  Run in the BUILD container (needs /root/reference); the GPU box
only reads the committed fixtures.

Two kinds of fixture:

1. Known-answer vectors whose expected bytes are LITERALS in this file.  They come from the
   upstream READMEs of the two third-party tools the reference shells out to
   (btbphylo/phylogeny.py:133 snp-sites, :149 snp-dists) or are
   derived by hand from the behavioural spec.  The script asserts that the
   CPU oracle reproduces every one of them before writing anything.

2. Boundary fixtures: the UNMODIFIED reference entry points
   btbphylo.phylogeny.snp_sites / build_snp_matrix / post_process_snps_csv are imported from
   /root/reference and run with `snp-sites` / `snp-dists` on PATH resolving to the oracle
   executables (the real binaries do not exist offline).  This pins the call contract
   (argv, file names, return value, the pandas relabel) -- the arithmetic in these files is the
   oracle's, which is why DESIGN.md says "parity unpinned" for the arithmetic itself.
"""
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from btb_phylo_b200 import synth  # noqa: E402

REFERENCE = "/root/reference"

# ------------------------------------------------------------------ literal known answers
KAT1_IN = b">seq1\nAGTCAGTC\n>seq2\nAGGCAGTC\n>seq3\nAGTGAGTA\n>seq4\nTGTTAGAC\n"
KAT1_TSV = (b"snp-dists 0.8.2\tseq1\tseq2\tseq3\tseq4\n"
            b"seq1\t0\t1\t2\t3\nseq2\t1\t0\t3\t4\nseq3\t2\t3\t0\t4\nseq4\t3\t4\t4\t0\n")
KAT1_MOLTEN_CSV_B = b"".join(
    ("%s,%s,%d\n" % (a, b, d)).encode()
    for a, row in zip(["seq1", "seq2", "seq3", "seq4"],
                      [[0, 1, 2, 3], [1, 0, 3, 4], [2, 3, 0, 4], [3, 4, 4, 0]])
    for b, d in zip(["seq1", "seq2", "seq3", "seq4"], row))

KAT2_IN = b">sample1\nAGACACAGTCAC\n>sample2\nAGACAC----AC\n>sample3\nAAACGCATTCAN\n"
KAT2_DEFAULT = b">sample1\nGAG\n>sample2\nGA-\n>sample3\nAGT\n"
KAT2_PURE = b">sample1\nGA\n>sample2\nGA\n>sample3\nAG\n"

KAT3_IN = (b">A_consensus\nACGTACGTNA\nACGT\n>B_consensus\nACGAACGTAA\nACGT\n"
           b">C_consensus\nACGAACCTAA\nACGT\n>D_consensus\nTCGAACCTAA\nAC-T\n")
KAT3_FAS = b">A_consensus\nATG\n>B_consensus\nAAG\n>C_consensus\nAAC\n>D_consensus\nTAC\n"
KAT3_CSV = (b"snp-dists 0.8.2,A_consensus,B_consensus,C_consensus,D_consensus\n"
            b"A_consensus,0,1,2,3\nB_consensus,1,0,1,2\nC_consensus,2,1,0,1\nD_consensus,3,2,1,0\n")

# hand-derived edge vector for snp-dists modes (derivation in the comments of the test file)
EDGE_D_IN = b">s1\nACGTNacgt-\n>s2 comment here\nACGANACGTA\n>s3\nA.GTRACGTT\n"
EDGE_D = {
    "default": [[0, 1, 0], [1, 0, 2], [0, 2, 0]],
    "a": [[0, 2, 2], [2, 0, 3], [2, 3, 0]],
    "k": [[0, 1, 0], [1, 0, 2], [0, 2, 0]],
    "ak": [[0, 6, 6], [6, 0, 3], [6, 3, 0]],
}
# hand-derived edge vector for snp-sites: CRLF, wrap, lower case, header comment
EDGE_S_IN = b">s1 desc\r\nACgT\r\nAC\r\n>s2\r\nACGA\r\nNC\r\n>s3\r\nAAGA\r\nGC\r\n"
EDGE_S_PURE = b">s1\nCT\n>s2\nCA\n>s3\nAA\n"
EDGE_S_DEFAULT = b">s1\nCTA\n>s2\nCAN\n>s3\nAAG\n"


def w(name, data):
    with open(os.path.join(HERE, name), "wb") as f:
        f.write(data)


def check_literals(tmp):
    def dists(inp, flags):
        p = os.path.join(tmp, "in.fa")
        open(p, "wb").write(inp)
        o = os.path.join(tmp, "out.txt")
        r = oracle.run_snp_dists(p, o, threads=2, flags=flags)
        assert r.returncode == 0, r.stderr
        return open(o, "rb").read(), r.stderr

    def sites(inp, pure):
        p = os.path.join(tmp, "in.fa")
        open(p, "wb").write(inp)
        o = os.path.join(tmp, "out.fas")
        r = oracle.run_snp_sites(p, o, pure=pure)
        assert r.returncode == 0, r.stderr
        return open(o, "rb").read()

    out, err = dists(KAT1_IN, ())
    assert out == KAT1_TSV, out
    assert err == b"This is snp-dists 0.8.2\nWill use 2 threads.\nRead 4 sequences of length 8\n", err
    out, _ = dists(KAT1_IN, ("-c", "-m"))
    assert out == KAT1_MOLTEN_CSV_B, out
    assert sites(KAT2_IN, False) == KAT2_DEFAULT
    assert sites(KAT2_IN, True) == KAT2_PURE
    assert sites(KAT3_IN, True) == KAT3_FAS
    out, _ = dists(KAT3_FAS, ("-c",))
    assert out == KAT3_CSV
    names = ["s1", "s2", "s3"]
    for mode, flags in (("default", ("-c",)), ("a", ("-c", "-a")), ("k", ("-c", "-k")), ("ak", ("-c", "-a", "-k"))):
        out, _ = dists(EDGE_D_IN, flags)
        assert out == oracle.format_matrix(names, EDGE_D[mode]), (mode, out)
    assert sites(EDGE_S_IN, True) == EDGE_S_PURE
    assert sites(EDGE_S_IN, False) == EDGE_S_DEFAULT


def reference_phylogeny():
    """Import the unmodified reference module; boto3/botocore are absent here, stub them
    (btbphylo/utils.py:9-10 imports them at module top; the phylogeny path never calls them)."""
    for name in ("boto3", "botocore", "botocore.exceptions"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["botocore.exceptions"].ClientError = type("ClientError", (Exception,), {})
    sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
    sys.path.insert(0, REFERENCE)
    import btbphylo.phylogeny as ph
    return ph


def through_reference(ph, tmp, multi_fasta_bytes, tag, postprocess=True):
    """multi_fasta -> snps.fas -> snps.csv via the reference's own functions."""
    mf = os.path.join(tmp, "multi_fasta.fas")
    sf = os.path.join(tmp, "snps.fas")
    sc = os.path.join(tmp, "snps.csv")
    open(mf, "wb").write(multi_fasta_bytes)
    meta = ph.snp_sites(sf, mf)                      # btbphylo/phylogeny.py:128
    ph.build_snp_matrix(sc, sf, threads=4)           # btbphylo/phylogeny.py:144
    w(tag + "_multi_fasta.fas", multi_fasta_bytes)
    w(tag + "_snps.fas", open(sf, "rb").read())
    w(tag + "_snps.csv", open(sc, "rb").read())
    w(tag + "_meta.json", (json.dumps(meta, sort_keys=True) + "\n").encode())
    if postprocess:
        ph.post_process_snps_csv(sc)                 # btbphylo/phylogeny.py:172
        w(tag + "_snps_postprocessed.csv", open(sc, "rb").read())
    return meta


def label_fixtures(ph, tmp):
    """Sample-label vectors.  Inputs: the literal lists the reference's own unit
    tests hold (tests/phylogeny_test.py test_process_sample_name, tests/utils_test.py
    test_extract_submission_no -- read from those files with `ast`, nothing is executed) plus a few
    extra names; expected values: the reference test's own expectation AND what the unmodified
    reference functions return here.  Then a snps.csv with those names through the unmodified
    post_process_snps_csv (pandas read_csv -> rename -> to_csv)."""
    import ast

    def literal_lists(path, func):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.FunctionDef) and node.name == func:
                found = {}
                for st in ast.walk(node):
                    if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name) and isinstance(st.value, ast.List):
                        found[st.targets[0].id] = ast.literal_eval(st.value)
                return found
        raise KeyError(func)

    import btbphylo.utils as ref_utils
    pt = literal_lists(os.path.join(REFERENCE, "tests", "phylogeny_test.py"), "test_process_sample_name")
    ut = literal_lists(os.path.join(REFERENCE, "tests", "utils_test.py"), "test_extract_submission_no")
    extra = ["af-12-3456-89_consensus", "x_consensus_consensus", "_consensus", "consensus", "123-45678-901",
             "1-2345-67", "12-345-67", "12-345678-90", "99-1234-5612-34567-89", "AF-61-00123-19_consensus",
             "s12-3456-78.trimmed", "12-3456-7", "a12-34567-89b", "Sample_12-1234-12_consensus"]
    fixture = {
        "process_sample_name": [{"in": i, "ref_test_expects": o, "ref_returns": ph.process_sample_name(i)}
                                for i, o in zip(pt["test_input"], pt["test_output"])] +
                               [{"in": i, "ref_test_expects": None, "ref_returns": ph.process_sample_name(i)} for i in extra],
        "extract_submission_no": [{"in": i, "ref_test_expects": o, "ref_returns": ref_utils.extract_submission_no(i)}
                                  for i, o in zip(ut["test_input"], ut["test_output"])],
    }
    for group in fixture.values():
        for row in group:
            assert row["ref_test_expects"] in (None, row["ref_returns"]), row
    w("labels_expected.json", (json.dumps(fixture, indent=1, sort_keys=True) + "\n").encode())

    # a matrix whose names exercise the relabel (duplicates after mapping included)
    names = [n for n in dict.fromkeys(pt["test_input"] + extra) if n and not n.isdigit()]
    seqs = synth.snp_alignment(len(names), 120, seed=31, heavy_n=0.1)
    fas = os.path.join(tmp, "labels.fas")
    csv = os.path.join(tmp, "labels.csv")
    open(fas, "wb").write(synth.fasta_bytes(names, seqs, wrap=0))
    ph.build_snp_matrix(csv, fas, threads=2)
    w("labels_snps.fas", open(fas, "rb").read())
    w("labels_snps.csv", open(csv, "rb").read())
    ph.post_process_snps_csv(csv)
    w("labels_snps_postprocessed.csv", open(csv, "rb").read())


def main():
    oracle.build(force=True)
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        check_literals(tmp)
        w("kat1_in.fa", KAT1_IN); w("kat1_expected.tsv", KAT1_TSV); w("kat1_expected_molten.csv", KAT1_MOLTEN_CSV_B)
        w("kat2_in.fa", KAT2_IN); w("kat2_expected_default.fas", KAT2_DEFAULT); w("kat2_expected_pure.fas", KAT2_PURE)
        w("edge_dists_in.fa", EDGE_D_IN)
        w("edge_dists_expected.json", (json.dumps(EDGE_D, sort_keys=True) + "\n").encode())
        w("edge_sites_in.fa", EDGE_S_IN); w("edge_sites_expected_pure.fas", EDGE_S_PURE)
        w("edge_sites_expected_default.fas", EDGE_S_DEFAULT)

        if not os.path.isdir(REFERENCE):
            print("no /root/reference here: literal fixtures written, boundary fixtures left as committed")
            return
        os.environ["PATH"] = oracle.BUILD + os.pathsep + os.environ["PATH"]
        ph = reference_phylogeny()
        meta = through_reference(ph, tmp, KAT3_IN, "kat3")
        assert meta == {"number_of_snps": 3}, meta
        assert open(os.path.join(HERE, "kat3_snps.fas"), "rb").read() == KAT3_FAS
        assert open(os.path.join(HERE, "kat3_snps.csv"), "rb").read() == KAT3_CSV
        # BASELINE config 1: 4 samples, 831 SNP sites (mirrors Readme.md:64-66 of the reference)
        seqs = synth.c1_alignment(G=20000, L_target=831, seed=1)
        names = synth.sample_names(4)
        meta = through_reference(ph, tmp, synth.fasta_bytes(names, seqs, wrap=60), "c1")
        assert meta == {"number_of_snps": 831}, meta
        label_fixtures(ph, tmp)
        print("golden fixtures written to", HERE)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
