"""Sample labels for ViewBovine: process_sample_name / extract_submission_no /
post_process_snps_csv of btbphylo/phylogeny.py:172-216 and btbphylo/utils.py:110-119.

Pinned by the literal vectors of the reference's own unit tests (tests/phylogeny_test.py
test_process_sample_name, tests/utils_test.py test_extract_submission_no) and by files the
unmodified reference functions wrote (tests/golden/make_golden.py label_fixtures); the pandas
round trip the reference performs is replayed live as a second check.  Host-only: no GPU."""
import ctypes
import json
import random

import pytest

from tests.conftest import golden


@pytest.fixture()
def ph(capi):
    from btb_phylo_b200 import phylogeny
    return phylogeny


def c_label(capi, name):
    buf = ctypes.create_string_buffer(len(name.encode()) + 16)
    capi.check(capi.lib().btb_process_sample_name(name.encode(), buf, len(buf)))
    return buf.value.decode()


def test_reference_unit_test_vectors(ph, capi):
    exp = json.loads(golden("labels_expected.json"))
    assert len(exp["process_sample_name"]) > 40 and len(exp["extract_submission_no"]) > 20
    for row in exp["process_sample_name"]:
        assert ph.process_sample_name(row["in"]) == row["ref_returns"], row
        assert c_label(capi, row["in"]) == row["ref_returns"], row
        if row["ref_test_expects"] is not None:
            assert row["ref_returns"] == row["ref_test_expects"]
    for row in exp["extract_submission_no"]:
        assert ph.extract_submission_no(row["in"]) == row["ref_returns"] == row["ref_test_expects"], row


def test_c_and_python_rules_agree_on_random_ascii_names(ph, capi):
    rng = random.Random(5)
    alphabet = "0123456789-_AFafxz.consu"
    for _ in range(4000):
        name = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 24)))
        if rng.random() < 0.3:
            name += "_consensus"
        if rng.random() < 0.3:
            name = name[: rng.randrange(0, 6)] + "%02d-%0*d-%02d" % (rng.randrange(100), rng.choice((3, 4, 5, 6)), rng.randrange(1000), rng.randrange(100)) + name[6:]
        assert c_label(capi, name) == ph.process_sample_name(name), name


@pytest.mark.parametrize("tag", ["kat3", "c1", "labels"])
def test_post_process_snps_csv_matches_reference_output(ph, tmp_path, tag):
    sc = tmp_path / "snps.csv"
    sc.write_bytes(golden(tag + "_snps.csv"))
    ph.post_process_snps_csv(str(sc))
    assert sc.read_bytes() == golden(tag + "_snps_postprocessed.csv")
    assert not (tmp_path / "snps.csv.tmp").exists()


def pandas_round_trip(path):
    """What btbphylo/phylogeny.py:172-203 does to the file, replayed with pandas."""
    import pandas as pd
    from btb_phylo_b200.phylogeny import process_sample_name
    m = pd.read_csv(path, index_col=0)
    labels = m.index.map(process_sample_name)
    out = m.copy(deep=True)
    out.index = labels
    out.columns = labels
    out.to_csv(path)


def test_streaming_relabel_equals_pandas_round_trip(ph, oracle, tmp_path):
    import numpy as np
    rng = np.random.default_rng(9)
    n = 300
    names = ["AF-%02d-%05d-%02d_consensus" % (rng.integers(100), rng.integers(100000), rng.integers(100)) for _ in range(n - 4)]
    names += ["ref", "20-1234-20", "odd.name_consensus", 'quo"te']
    d = rng.integers(0, 70000, size=(n, n)).astype(np.uint32)
    d = np.triu(d, 1)
    d = d + d.T
    text = oracle.format_matrix(names, d)
    a, b = tmp_path / "a.csv", tmp_path / "b.csv"
    a.write_bytes(text)
    b.write_bytes(text)
    ph.post_process_snps_csv(str(a))
    pandas_round_trip(str(b))
    assert a.read_bytes() == b.read_bytes()


@pytest.mark.parametrize("names", [["NA", "b"], ["a", "null"], ["12345678", "42"], ["1.5", "2e3"], ["True", "false"]])
def test_inputs_the_pandas_round_trip_cannot_relabel_raise(ph, oracle, tmp_path, names):
    import numpy as np
    sc = tmp_path / "snps.csv"
    sc.write_bytes(oracle.format_matrix(names, np.zeros((2, 2), dtype=np.uint32)))
    before = sc.read_bytes()
    with pytest.raises(Exception):
        ph.post_process_snps_csv(str(sc))
    assert sc.read_bytes() == before
    with pytest.raises(Exception):                      # the reference fails on these too
        pandas_round_trip(str(sc))


def test_relabel_rejects_wrong_label_count(capi, tmp_path):
    sc = tmp_path / "snps.csv"
    sc.write_bytes(golden("kat3_snps.csv"))
    rc = capi.lib().btb_relabel_csv(str(sc).encode(), str(sc).encode(), b",", b"A\nB\nC", 3)
    assert rc != capi.OK and "Length mismatch" in capi.last_error()
    assert sc.read_bytes() == golden("kat3_snps.csv")


def test_fasta_and_csv_names(capi, ph):
    import os
    from tests.conftest import GOLDEN
    from btb_phylo_b200.phylogeny import _names_from
    names = _names_from(capi.lib().btb_fasta_names, os.path.join(GOLDEN, "kat3_snps.fas").encode(), 2)
    assert names == ["A_consensus", "B_consensus", "C_consensus", "D_consensus"]
    assert _names_from(capi.lib().btb_csv_row_names, os.path.join(GOLDEN, "kat3_snps.csv").encode(), b",") == names


def test_relabel_fuzz_against_pandas(ph, oracle, tmp_path):
    """Random row names from a hostile alphabet (digits only, NA-like words, quotes, dots, signs, mixed
    case, submission numbers): whenever the reference's pandas round trip succeeds, the streaming
    relabel writes the same bytes; whenever it fails, so does ours (and the file is left alone)."""
    import numpy as np
    rng = random.Random(77)
    words = ["NA", "nan", "NULL", "None", "n/a", "1e5", "12", "-3", "+7", "0.5", ".5", "inf", "True", "false", "TRUE",
             "AF-12-34567-89", "af-21-1234-21_consensus", "x_consensus", "12-3456-78", "99-12345-99x", "a.b", "a'b", 'a"b',
             "#c", "d#", "e;f", "g|h", "i:j", "k/l", "m=n", "o p", " q", "r ", "1_000", "0x1F", "1,5"]
    agree = fail = 0
    for case in range(160):
        n = rng.randint(1, 6)
        names = []
        for _ in range(n):
            if rng.random() < 0.6:
                names.append(rng.choice(words))
            else:
                names.append("".join(rng.choice("ABab019-_.x") for _ in range(rng.randint(1, 9))))
        if any("," in x for x in names) and rng.random() < 0.8:
            continue                                          # commas break the matrix shape for both; keep a few
        d = np.array([[abs(i - j) * 3 for j in range(n)] for i in range(n)], dtype=np.uint32)
        text = oracle.format_matrix(names, d)
        a, b = tmp_path / "a.csv", tmp_path / "b.csv"
        a.write_bytes(text)
        b.write_bytes(text)
        try:
            pandas_round_trip(str(b))
            ref_ok = True
        except Exception:
            ref_ok = False
        try:
            ph.post_process_snps_csv(str(a))
            ours_ok = True
        except Exception:
            ours_ok = False
        assert ours_ok == ref_ok, (case, names, ref_ok, ours_ok)
        if ref_ok:
            assert a.read_bytes() == b.read_bytes(), (case, names)
            agree += 1
        else:
            assert a.read_bytes() == text, (case, names)
            fail += 1
    assert agree > 40 and fail > 10
