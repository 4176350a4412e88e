"""CPU: btb_fasta_frame_files (include/btb_phylo.h) -- the predicted framing of a list of consensus files, the host
half of `build_multi_fasta` + `snp-sites` (btbphylo/phylogeny.py:48-88,133) that the plane store uses to feed k1.
No GPU is involved: the call only reads file headers and boundaries."""
import ctypes

import numpy as np
import pytest

from btb_phylo_b200 import synth


def frame(capi, paths, cap=None, names_cap=None):
    lib = capi.lib()
    n = len(paths)
    cap = n if cap is None else cap
    arr = (ctypes.c_char_p * n)(*[p.encode() for p in paths])
    out = (ctypes.c_int64 * (6 * max(1, cap)))()
    nbuf = ctypes.create_string_buffer(4096 * n + 16 if names_cap is None else names_cap)
    nrec = ctypes.c_int64(0)
    rc = lib.btb_fasta_frame_files(arr, n, out, cap, nbuf, len(nbuf), ctypes.byref(nrec))
    recs = np.frombuffer(out, dtype=np.int64).reshape(-1, 6)[: max(0, nrec.value)].copy()
    return rc, nrec.value, recs, nbuf.value


def write(tmp_path, names, aln, wrap=60, eol=b"\n", comments=None):
    paths = []
    for i, (nm, row) in enumerate(zip(names, aln)):
        p = tmp_path / (nm + ".fas")
        p.write_bytes(synth.fasta_bytes([nm], row[None, :], wrap=wrap, eol=eol, comments=[comments[i]] if comments else None))
        paths.append(str(p))
    return paths


@pytest.mark.parametrize("wrap,eol", [(60, b"\n"), (70, b"\r\n"), (0, b"\n"), (33, b"\n")])
def test_regular_files_are_framed_record_by_record(capi, tmp_path, wrap, eol):
    n, G = 7, 1234
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 60, seed=5, n_frac=0.02)
    comments = ["c" * (i % 3) for i in range(n)]                       # headers of different lengths
    paths = write(tmp_path, names, aln, wrap, eol, comments)
    rc, nrec, recs, nm = frame(capi, paths)
    assert rc == capi.OK and nrec == n
    assert nm.decode().split("\n") == names                            # the name ends at the first blank of the header
    for i in range(n):
        f, off, raw, bases, lw, el = (int(v) for v in recs[i])
        text = open(paths[i], "rb").read()
        assert f == i and bases == G
        assert off == text.index(eol) + len(eol) and off + raw <= len(text)
        assert lw == (wrap if wrap and wrap < G else 0)
        assert el == (len(eol) if lw else 0) or (lw == 0 and el in (0, len(eol)))
        body = text[off:off + raw].replace(b"\r", b"").replace(b"\n", b"")
        assert body == aln[i].tobytes()


def test_layouts_the_prediction_does_not_take_report_minus_one(capi, tmp_path):
    n, G = 4, 500
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 30, seed=6, n_frac=0.0)
    paths = write(tmp_path, names, aln)
    # another wrap in the second file: the records are not of one layout
    open(paths[1], "wb").write(synth.fasta_bytes([names[1]], aln[1][None, :], wrap=70))
    rc, nrec, _, _ = frame(capi, paths)
    assert rc == capi.OK and nrec == -1
    # a record of another length
    open(paths[1], "wb").write(synth.fasta_bytes([names[1]], aln[1][None, :G - 7], wrap=60))
    rc, nrec, _, _ = frame(capi, paths)
    assert rc == capi.OK and nrec == -1
    # gzip input
    import gzip
    open(paths[1], "wb").write(gzip.compress(synth.fasta_bytes([names[1]], aln[1][None, :], wrap=60)))
    rc, nrec, _, _ = frame(capi, paths)
    assert rc == capi.OK and nrec == -1


def test_buffers_that_are_too_small_are_an_error(capi, tmp_path):
    n, G = 3, 200
    names = synth.sample_names(n)
    aln = synth.genome_alignment(n, G, 10, seed=7, n_frac=0.0)
    paths = write(tmp_path, names, aln)
    rc, _, _, _ = frame(capi, paths, cap=2)
    assert rc == capi.EINVAL and "do not fit" in capi.last_error()
    rc, _, _, _ = frame(capi, paths, names_cap=4)
    assert rc == capi.EINVAL
    rc, _, _, _ = frame(capi, [str(tmp_path / "missing.fas")])
    assert rc != capi.OK
