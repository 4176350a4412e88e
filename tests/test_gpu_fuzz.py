"""GPU (-m gpu): randomised parity of the two file-level entry points against the CPU oracle on small
FASTA-ish inputs with the framing oddities the shared parser defines: mixed line widths (below and above
the GPU's 32-base minimum), LF / CRLF, blank lines, header comments, lower case, N / - / IUPAC, a missing
final newline, lines that start with '>' or '@' inside a sequence, unequal lengths.  Success / failure
must agree, and on success the bytes."""
import random

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def random_fasta(rng, equal=True, markers=False):
    n = rng.randint(1, 7)
    G = rng.randint(1, 260)
    alphabet = "ACGT" * 6 + "acgt" + "N-" + ("RYKM?" if rng.random() < 0.3 else "")
    ref = [rng.choice("ACGT") for _ in range(G)]
    parts = []
    for i in range(n):
        g = G if equal or rng.random() < 0.7 else max(0, G + rng.randint(-3, 3))
        seq = [ref[k] if k < G and rng.random() < 0.85 else rng.choice(alphabet) for k in range(g)]
        seq = "".join(seq)
        eol = "\r\n" if rng.random() < 0.25 else "\n"
        w = rng.choice([0, 0, rng.randint(1, 31), rng.randint(32, 90), 60, 70])
        head = ">s%d" % i + (" some comment %d" % i if rng.random() < 0.3 else "") + ("\tx" if rng.random() < 0.1 else "")
        lines = [seq] if not w else [seq[k:k + w] for k in range(0, len(seq), w)] or [""]
        body = []
        for li, line in enumerate(lines):
            if markers and li > 0 and rng.random() < 0.15:
                line = rng.choice(">@") + line[1:] if line else line
            body.append(line)
            if rng.random() < 0.05:
                body.append("")                              # a blank line inside the record
        text = head + eol + eol.join(body) + eol
        parts.append(text)
    data = "".join(parts)
    if rng.random() < 0.2:
        data = data.rstrip("\r\n")                           # no final newline
    if rng.random() < 0.1:
        data = "junk line\n" + data
    return data.encode()


@pytest.mark.parametrize("seed", range(6))
def test_snp_sites_fuzz(capi, oracle, tmp_path, seed):
    from btb_phylo_b200 import phylogeny as ph
    rng = random.Random(1000 + seed)
    mf, sf, of = tmp_path / "m.fa", tmp_path / "s.fas", tmp_path / "o.fas"
    for case in range(40):
        data = random_fasta(rng, equal=rng.random() < 0.85, markers=rng.random() < 0.3)
        mf.write_bytes(data)
        for f in (sf, of):
            if f.exists():
                f.unlink()
        r = oracle.run_snp_sites(str(mf), str(of))
        try:
            ph.snp_sites(str(sf), str(mf), threads=2)
            ok = True
        except Exception:
            ok = False
        assert ok == (r.returncode == 0), (seed, case, data, r.stderr)
        if ok:
            assert sf.read_bytes() == of.read_bytes(), (seed, case, data)
        else:
            assert not sf.exists(), (seed, case)


@pytest.mark.parametrize("seed", range(4))
def test_snp_dists_fuzz(capi, oracle, tmp_path, seed):
    rng = random.Random(2000 + seed)
    lib = capi.lib()
    mf, sc, oc = tmp_path / "m.fa", tmp_path / "s.csv", tmp_path / "o.csv"
    for case in range(30):
        data = random_fasta(rng, equal=rng.random() < 0.9, markers=rng.random() < 0.2)
        mf.write_bytes(data)
        flags = [f for f in ("-a", "-k") if rng.random() < 0.4]
        sep_flag = rng.random() < 0.5
        r = oracle.run_snp_dists(str(mf), str(oc), threads=2, flags=tuple((["-c"] if sep_flag else []) + flags))
        for engine in (-1, 0, 1):
            rc = lib.btb_snp_dists(str(mf).encode(), str(sc).encode(), int("-a" in flags), int("-k" in flags), 0,
                                   b"," if sep_flag else b"\t", 1, 1, 1, 2, engine, None, None)
            assert (rc == 0) == (r.returncode == 0), (seed, case, engine, data, capi.last_error(), r.stderr)
            if rc == 0:
                assert sc.read_bytes() == oc.read_bytes(), (seed, case, engine, flags, data)


@pytest.mark.parametrize("seed", range(4))
def test_consensus_lists_cut_anywhere_equal_the_concatenation(capi, oracle, tmp_path, seed):
    """A multi-FASTA cut into files at arbitrary byte positions (even inside a line or a header) and
    passed as a list gives what the concatenated file gives."""
    from btb_phylo_b200 import phylogeny as ph
    rng = random.Random(3000 + seed)
    mf, sf, of = tmp_path / "m.fa", tmp_path / "s.fas", tmp_path / "o.fas"
    for case in range(25):
        data = random_fasta(rng, equal=rng.random() < 0.9, markers=rng.random() < 0.2)
        cuts = sorted(rng.sample(range(len(data) + 1), k=min(len(data) + 1, rng.randint(1, 5)))) if rng.random() < 0.6 else \
            [i for i in range(len(data)) if data[i:i + 1] == b">" and i > 0]          # ... or exactly at the records
        pieces, last = [], 0
        for c in cuts + [len(data)]:
            pieces.append(data[last:c])
            last = c
        paths = []
        for k, piece in enumerate(pieces):
            f = tmp_path / ("p%d_%d.fa" % (case, k))
            f.write_bytes(piece)
            paths.append(str(f))
        mf.write_bytes(data)
        for f in (sf, of):
            if f.exists():
                f.unlink()
        r = oracle.run_snp_sites(str(mf), str(of))
        try:
            ph.snp_sites(str(sf), paths, threads=2)
            ok = True
        except Exception:
            ok = False
        assert ok == (r.returncode == 0), (seed, case, data, cuts)
        if ok:
            assert sf.read_bytes() == of.read_bytes(), (seed, case, data, cuts)


@pytest.mark.parametrize("seed", range(4))
def test_distance_engines_fuzz_on_arbitrary_bytes(capi, oracle, seed):
    """Level 2 on raw byte matrices: any byte value may occur (control bytes, '.', high bytes), sizes
    straddle the tile and K-block boundaries, all engines, all modes."""
    import numpy as np
    from btb_phylo_b200 import device
    rng = np.random.default_rng(4000 + seed)
    for case in range(12):
        n = int(rng.choice([1, 2, 3, 31, 127, 128, 129, 255, 256, 257, 300, 511, 513, 700]))
        L = int(rng.choice([1, 2, 15, 16, 17, 31, 63, 64, 65, 255, 256, 257, 300, 1000]))
        kind = int(rng.integers(0, 4))
        if kind == 0:                                         # dense DNA
            pool = np.frombuffer(b"ACGT", dtype=np.uint8)
        elif kind == 1:                                       # DNA with gaps, IUPAC, lower case, dots
            pool = np.frombuffer(b"ACGTACGTACGTacgtNN--.RYKMSWnx?", dtype=np.uint8)
        elif kind == 2:                                       # a few arbitrary byte values
            pool = rng.integers(0, 256, size=int(rng.integers(2, 9))).astype(np.uint8)
        else:                                                 # anything
            pool = np.arange(256, dtype=np.uint8)
        seqs = pool[rng.integers(0, len(pool), size=(n, L))]
        base = seqs[rng.integers(0, n)]                       # clade-like: most rows close to one of them
        keep = rng.random((n, L)) < 0.8
        seqs = np.where(keep, base[None, :], seqs).astype(np.uint8)
        for allchars, keepcase in ((False, False), (True, False), (False, True), (True, True)):
            want = oracle.snp_dists_rows(seqs, allchars=allchars, keepcase=keepcase, threads=4)
            for engine in (capi.ENGINE_AUTO, capi.ENGINE_POPC, capi.ENGINE_UMMA):
                got = device.distance_matrix_host(seqs, allchars=allchars, keepcase=keepcase, engine=engine)
                assert (got == want).all(), (seed, case, n, L, kind, allchars, keepcase, engine)


@pytest.mark.parametrize("seed", range(5))
def test_snp_sites_fuzz_at_kernel_scale(capi, oracle, tmp_path, seed):
    """Genomes long enough for several CTAs per sample and many plane words, laid out like consensus
    files (one wrap width per file, so the predicted framing is tried first), with anomalies injected
    into a few records: a blank line, a line one base short followed by one a base long, a CR in the
    middle of a line, a tab, a line starting with '>', lower case, N runs, a record without SNPs."""
    import numpy as np
    from btb_phylo_b200 import phylogeny as ph
    rng = random.Random(5000 + seed)
    nrng = np.random.default_rng(5000 + seed)
    mf, sf, of = tmp_path / "m.fa", tmp_path / "s.fas", tmp_path / "o.fas"
    for case in range(10):
        n = rng.randint(3, 14)
        G = rng.choice([32768, 32769, 40000, 65536, 100003, 131071])
        w = rng.choice([60, 60, 70, 80, 33, 32, 0, 200])
        eol = b"\r\n" if rng.random() < 0.2 else b"\n"
        ref = nrng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), G)
        rows = []
        for i in range(n):
            r = ref.copy()
            k = rng.randint(0, 400)
            r[nrng.integers(0, G, k)] = nrng.choice(np.frombuffer(b"ACGTacgtN-", dtype=np.uint8), k)
            if rng.random() < 0.3:
                a = rng.randint(0, G - 500)
                r[a:a + rng.randint(1, 500)] = ord("N")
            rows.append(r.tobytes())
        recs = []
        for i, row in enumerate(rows):
            lines = [row] if not w else [row[k:k + w] for k in range(0, G, w)]
            anomaly = rng.random()
            if w and len(lines) > 4 and anomaly < 0.30:
                j = rng.randint(1, len(lines) - 3)
                kind = rng.randint(0, 4)
                if kind == 0:
                    lines.insert(j, b"")
                elif kind == 1:
                    lines[j], lines[j + 1] = lines[j][:-1], lines[j][-1:] + lines[j + 1]
                elif kind == 2:
                    lines[j] = lines[j][:5] + b"\r" + lines[j][6:]
                elif kind == 3:
                    lines[j] = lines[j][:9] + b"\t" + lines[j][10:]
                else:
                    lines[j] = b">" + lines[j][1:]
            recs.append(b">smp%d" % i + (b" len=%d" % G if i % 3 == 0 else b"") + eol + eol.join(lines) + eol)
        data = b"".join(recs)
        if rng.random() < 0.25:
            data = data[: -len(eol)]
        mf.write_bytes(data)
        for f in (sf, of):
            if f.exists():
                f.unlink()
        r = oracle.run_snp_sites(str(mf), str(of))
        try:
            ph.snp_sites(str(sf), str(mf), threads=4)
            ok = True
        except Exception:
            ok = False
        assert ok == (r.returncode == 0), (seed, case, n, G, w, eol, r.stderr[-300:])
        if ok:
            assert sf.read_bytes() == of.read_bytes(), (seed, case, n, G, w, eol)


@pytest.mark.parametrize("dense", [True, False])
def test_block_row_ranges_fuzz(capi, oracle, dense):
    """btb_distance_rows on random block-row ranges of a 7,000-sequence alignment: even and odd first
    rows, ranges small enough for the half-tile kernel and large enough for the CTA-pair and the
    single-CTA full-tile kernels, with and without mirroring.  Cells of the range equal the oracle,
    every other cell of the matrix keeps its sentinel."""
    import numpy as np
    import torch
    from btb_phylo_b200 import device, synth
    from tests.test_gpu_distance import to_device
    n, L = 7000, 300
    seqs = synth.snp_alignment(n, L, seed=61, heavy_n=0.0 if dense else 0.08)
    want = torch.from_numpy(oracle.snp_dists_rows(seqs, threads=8).astype(np.int32)).cuda()
    plan = device.DistancePlan(to_device(seqs), length=L, engine=capi.ENGINE_UMMA)
    assert (plan.sigma == capi.SIGMA_DENSE4) == dense
    plan.encode()
    T = device.block_rows(n)
    rng = random.Random(62)
    ranges = [(0, T), (1, T), (0, 1), (T - 1, T), (2, 3), (3, 4)]
    while len(ranges) < 22:
        lo = rng.randint(0, T - 1)
        ranges.append((lo, rng.randint(lo, T)))
    rows_idx = torch.arange(n, device="cuda")
    upper = rows_idx[None, :] >= rows_idx[:, None]
    for lo, hi in ranges:
        for mirror in (True, False):
            out = torch.full((n, (n + 7) // 8 * 8), 0xEEEE, dtype=torch.uint16, device="cuda")
            plan.compute_rows(out, lo, hi, mirror=mirror)
            torch.cuda.synchronize()
            got = out[:, :n].to(torch.int32)
            r0, r1 = min(n, lo * 256), min(n, hi * 256)
            in_rows = (rows_idx >= r0) & (rows_idx < r1)
            must = in_rows[:, None] & upper                                  # (r, c >= r), r in the range
            assert bool((got[must] == want[must]).all()), (dense, lo, hi, mirror)
            if mirror:
                must_t = must.T
                assert bool((got[must_t] == want[must_t]).all()), (dense, lo, hi, mirror)
            # lower halves of diagonal tiles / mirrored cells may or may not be written; everything else must not be
            same_block = (rows_idx[:, None] // 256) == (rows_idx[None, :] // 256)
            may = must | must.T | (in_rows[:, None] & same_block) | (in_rows[None, :] & same_block)
            if not mirror:
                may = must | (in_rows[:, None] & same_block)
            untouched = ~may
            assert bool((got[untouched] == 0xEEEE).all()), (dense, lo, hi, mirror)
            stray = may & ~must & ~(must.T if mirror else torch.zeros_like(must))
            assert bool(((got[stray] == want[stray]) | (got[stray] == 0xEEEE)).all()), (dense, lo, hi, mirror)
