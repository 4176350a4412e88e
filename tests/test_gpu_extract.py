"""GPU (-m gpu): the extraction stage (k1 pack, k2 column mask, k3 select + compact) through the
C-ABI against the CPU oracle -- identical column selection and identical SNP bytes."""
import ctypes

import numpy as np
import pytest
import torch

from btb_phylo_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def extract_host(capi, aln):
    n, G = aln.shape
    aln = np.ascontiguousarray(aln)
    keep = np.zeros(G, dtype=np.uint8)
    L = ctypes.c_int64(0)
    capi.check(capi.lib().btb_extract_host(aln.ctypes.data, n, G, aln.strides[0], 1, 0, keep.ctypes.data, ctypes.byref(L), None, 0))
    snps = np.zeros((n, max(1, L.value)), dtype=np.uint8)
    L2 = ctypes.c_int64(0)
    capi.check(capi.lib().btb_extract_host(aln.ctypes.data, n, G, aln.strides[0], 1, 0, None, ctypes.byref(L2), snps.ctypes.data, snps.strides[0]))
    assert L2.value == L.value
    return keep.astype(bool), snps[:, : L.value]


@pytest.mark.parametrize("n,G,snps", [(1, 40, 0), (2, 33, 5), (4, 20000, 900), (37, 8191, 400), (37, 8193, 400), (300, 100003, 3000)])
def test_extract_matches_oracle(capi, oracle, n, G, snps):
    aln = synth.genome_alignment(n, G, snps, seed=n + G, n_frac=0.02, lower=0.01)
    if G > 100:
        aln[0, 50] = ord("-"); aln[n - 1, 51] = ord("R"); aln[0, 52] = ord("\t")
    keep, got = extract_host(capi, aln)
    want = oracle.snp_sites_mask(aln, pure=True)
    assert (keep == want).all()
    up = aln[:, want].copy()
    up[(up >= 97) & (up <= 122)] -= 32
    assert (got == up).all()


def test_level3_pack_wrapped_text(capi, oracle):
    """k1 on real FASTA text: 60/70-column wrap, CRLF, unwrapped and ragged header lengths in one
    chunk; a record with a misplaced terminator must be flagged, never silently mis-packed."""
    n, G = 9, 12345
    aln = synth.genome_alignment(n, G, 500, seed=77, n_frac=0.03, lower=0.02)
    names = synth.sample_names(n)
    wraps = [60, 70, 0, 61, 32, 60, 1000, 33, 8]            # 8 < 32: too short for the GPU layout -> flagged
    eols = [b"\n", b"\r\n", b"\n", b"\n", b"\n", b"\r\n", b"\n", b"\n", b"\n"]
    text = bytearray()
    offs, lws, els = [], [], []
    for i in range(n):
        rec = synth.fasta_bytes([names[i] + "x" * i], aln[i:i + 1], wrap=wraps[i], eol=eols[i])
        head = rec.index(eols[i]) + len(eols[i])
        offs.append(len(text) + head)
        lws.append(wraps[i] if wraps[i] and wraps[i] < G else 0)
        els.append(len(eols[i]) if lws[-1] else 0)
        text += rec
    # damage record 7: swap a base and its line terminator
    p = offs[7] + 33
    text[p - 1], text[p] = text[p], text[p - 1]
    text += b"\0" * 64
    lib = capi.lib()
    words = lib.btb_plane_words(G)
    d_text = torch.frombuffer(bytes(text), dtype=torch.uint8).cuda()
    d_off = torch.tensor(offs, dtype=torch.int64).cuda()
    d_lw = torch.tensor(lws, dtype=torch.int32).cuda()
    d_el = torch.tensor(els, dtype=torch.int32).cuda()
    planes = torch.zeros(3 * n * words, dtype=torch.int32).cuda()
    flags = torch.zeros(n, dtype=torch.int32).cuda()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    capi.check(lib.btb_pack_planes(P(d_text), len(text) - 64, P(d_off), P(d_lw), P(d_el), n, 0, n, G, P(planes), P(flags), st))
    torch.cuda.synchronize()
    f = flags.cpu().numpy()
    assert f[7] == 1 and f[8] == 1 and (f[:7] == 0).all()
    pl = planes.cpu().numpy().view(np.uint32).reshape(3, n, words)
    order = np.arange(32, dtype=np.uint32)                                                   # plane bit of column cl
    bits = lambda w: ((w[:, None] >> order[None, :]) & 1).reshape(-1)[:G]
    for i in range(n):
        if i in (7, 8):
            continue
        up = aln[i] & 0xDF
        valid = np.isin(up, np.frombuffer(b"ACGT", dtype=np.uint8)) & np.isin(aln[i], np.frombuffer(b"ACGTacgt", dtype=np.uint8))
        code = (aln[i] >> 1) & 3
        assert (bits(pl[2, i]) == valid).all(), i
        assert (bits(pl[0, i]) == (valid & (code & 1).astype(bool))).all(), i
        assert (bits(pl[1, i]) == (valid & (code >> 1).astype(bool))).all(), i


@pytest.mark.parametrize("groups", [0, 1, 3])
def test_level3_pack_skips_short_line_records_anywhere(capi, groups):
    """Records whose lines are shorter than 32 columns are flagged and skipped by k1 (the host walks them); they
    may stand first, in a row, or last in a CTA's share of the samples.  The text of the next record that IS packed
    is requested a sample ahead -- by its table entry when it follows directly, by a scan otherwise -- and the
    column state fused into the same pass must equal k2's over the packed planes."""
    G = 9973
    wraps = [8, 60, 16, 16, 16, 70, 0, 24, 61, 60, 8, 8, 33, 60, 12]
    n = len(wraps)
    aln = synth.genome_alignment(n, G, 400, seed=91, n_frac=0.02, lower=0.02)
    names = synth.sample_names(n)
    text = bytearray()
    offs, lws, els = [], [], []
    for i in range(n):
        eol = b"\r\n" if i % 4 == 1 else b"\n"
        rec = synth.fasta_bytes([names[i]], aln[i:i + 1], wrap=wraps[i], eol=eol)
        offs.append(len(text) + rec.index(eol) + len(eol))
        lws.append(wraps[i] if wraps[i] and wraps[i] < G else 0)
        els.append(len(eol) if lws[-1] else 0)
        text += rec
    nbytes = len(text)
    text += b"\0" * 64
    lib = capi.lib()
    capi.check(lib.btb_set_option(b"pack_gy", groups))
    try:
        words = lib.btb_plane_words(G)
        d_text = torch.frombuffer(bytes(text), dtype=torch.uint8).cuda()
        d_off = torch.tensor(offs, dtype=torch.int64).cuda()
        d_lw = torch.tensor(lws, dtype=torch.int32).cuda()
        d_el = torch.tensor(els, dtype=torch.int32).cuda()
        planes = torch.full((3 * n * words,), -1, dtype=torch.int32).cuda()
        flags = torch.zeros(n, dtype=torch.int32).cuda()
        state = torch.zeros(5 * words, dtype=torch.int32).cuda()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        capi.check(lib.btb_pack_planes_state(P(d_text), nbytes, P(d_off), P(d_lw), P(d_el), n, 0, n, G, P(planes), P(flags),
                                             P(state), st))
        torch.cuda.synchronize()
    finally:
        capi.check(lib.btb_set_option(b"pack_gy", 0))
    short = np.array([0 < w < 32 for w in wraps])
    assert (flags.cpu().numpy() == short.astype(np.int32)).all()
    pl = planes.cpu().numpy().view(np.uint32).reshape(3, n, words)
    order = np.arange(32, dtype=np.uint32)
    bits = lambda w: ((w[:, None] >> order[None, :]) & 1).reshape(-1)[:G]
    for i in range(n):
        if short[i]:
            assert not pl[:, i].any(), i                               # skipped records leave all-zero planes
            continue
        valid = np.isin(aln[i], np.frombuffer(b"ACGTacgt", dtype=np.uint8))
        code = (aln[i] >> 1) & 3
        assert (bits(pl[2, i]) == valid).all(), i
        assert (bits(pl[0, i]) == (valid & (code & 1).astype(bool))).all(), i
        assert (bits(pl[1, i]) == (valid & (code >> 1).astype(bool))).all(), i
    # k2 over the same planes (the skipped records' zero planes count as invalid everywhere, in both)
    ref = torch.zeros(5 * words, dtype=torch.int32).cuda()
    capi.check(lib.btb_column_mask(P(planes), n, n, G, P(ref), 0, st))
    torch.cuda.synchronize()
    got, want = state.cpu().numpy().view(np.uint32).reshape(5, words), ref.cpu().numpy().view(np.uint32).reshape(5, words)
    tail = np.uint32((1 << (G % 32)) - 1) if G % 32 else np.uint32(0xFFFFFFFF)
    last = (G - 1) // 32
    for plane in range(4):
        assert (got[plane, :last] == want[plane, :last]).all() and (got[plane, last] & tail) == (want[plane, last] & tail)
