"""Seeded synthetic inputs for the phylogeny hot path.

There is no network, so every test/bench input is generated here.  The model follows what the
reference's data looks like (M. bovis consensus sequences, btbphylo/phylogeny.py:26-88 builds the
multi-FASTA from them): one random reference genome, S "true SNP" columns whose minor allele is
carried by one contiguous clade of samples, `N` in *shared* masked blocks (i.i.d. N would leave
snp-sites -c with zero columns), sample names that
`process_sample_name` (btbphylo/phylogeny.py:206-216) can parse.
"""
import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def sample_names(n, suffix="_consensus"):
    """AF-12-00001-22_consensus style names (unique, parseable by the reference's regexes)."""
    return ["AF-12-%05d-%02d%s" % (i + 1, 10 + (i * 7) % 90, suffix) for i in range(n)]


def clade_ranges(n, s, rng):
    """For each of s SNP columns: the half-open sample range [a, b) that carries the minor allele.
    Heavy-tailed sizes: most SNPs are private to small clades, a few split the whole tree."""
    u = rng.random(s)
    size = np.maximum(1, np.minimum(n - 1, (n ** u * 0.5 + 0.5).astype(np.int64))) if n > 1 else np.ones(s, np.int64)
    a = (rng.random(s) * (n - size + 1)).astype(np.int64)
    return a, a + size


def snp_alignment(n, L, seed=4, heavy_n=0.0, iupac=0.0, lower=0.0, row_chunk=2048):
    """N x L upper-case SNP alignment bytes (what snps.fas holds), uint8 [n, L].

    heavy_n : fraction of cells replaced by N / - (half in shared column blocks, half private runs)
    iupac   : fraction of cells replaced by one of R Y K M S W
    lower   : fraction of cells lower-cased
    """
    rng = _rng(seed)
    base = BASES[rng.integers(0, 4, L)]
    minor = BASES[(np.searchsorted(BASES, base) + rng.integers(1, 4, L)) % 4]
    a, b = clade_ranges(n, L, rng)
    order = rng.permutation(n)
    out = np.empty((n, L), dtype=np.uint8)
    if heavy_n > 0:
        blk_cols = rng.random(L) < heavy_n  # columns inside shared masked blocks
        # make the blocks contiguous-ish: smooth with a run model
        runs = np.repeat(rng.random((L + 63) // 64) < heavy_n, 64)[:L]
        blk_cols = runs
    for r0 in range(0, n, row_chunk):
        r1 = min(n, r0 + row_chunk)
        pos = order[r0:r1, None]
        m = (pos >= a[None, :]) & (pos < b[None, :])
        blk = np.where(m, minor[None, :], base[None, :])
        if heavy_n > 0:
            crng = _rng(seed * 1000003 + r0)
            shared = blk_cols[None, :] & (crng.random((r1 - r0, 1)) < 0.9)
            blk[shared] = ord("N")
            private = crng.random(blk.shape) < heavy_n * 0.5
            blk[private] = ord("-")
        if iupac > 0:
            crng = _rng(seed * 7919 + r0 + 1)
            hit = crng.random(blk.shape) < iupac
            blk[hit] = np.frombuffer(b"RYKMSW", dtype=np.uint8)[crng.integers(0, 6, int(hit.sum()))]
        if lower > 0:
            crng = _rng(seed * 104729 + r0 + 2)
            hit = (crng.random(blk.shape) < lower) & (blk >= 65) & (blk <= 90)
            blk[hit] += 32
        out[r0:r1] = blk
    return out


def genome_alignment(n, G, n_snps, seed=1, n_frac=0.01, private_n=None, lower=0.0):
    """N x G consensus alignment, uint8 [n, G]: a shared reference with n_snps clade SNPs and
    ~n_frac of the genome inside shared masked blocks (each sample has each block as N with
    p = 0.9) plus optional private N at rate private_n per site (default ln2/n, so about half of
    the true SNP columns survive `-c`)."""
    rng = _rng(seed)
    ref = BASES[rng.integers(0, 4, G)]
    cols = np.sort(rng.choice(G, size=min(n_snps, G), replace=False))
    minor = BASES[(np.searchsorted(BASES, ref[cols]) + rng.integers(1, 4, len(cols))) % 4]
    a, b = clade_ranges(n, len(cols), rng)
    order = rng.permutation(n)
    out = np.repeat(ref[None, :], n, axis=0)
    sub = out[:, cols]
    pos = order[:, None]
    m = (pos >= a[None, :]) & (pos < b[None, :])
    sub[m] = np.broadcast_to(minor[None, :], sub.shape)[m]
    out[:, cols] = sub
    if n_frac > 0:
        blk_len = 110
        nblk = max(1, int(G * n_frac / blk_len))
        starts = rng.integers(0, max(1, G - blk_len), nblk)
        has = rng.random((n, nblk)) < 0.9
        for k, s0 in enumerate(starts):
            out[has[:, k], s0:s0 + blk_len] = ord("N")
    if private_n is None:
        private_n = np.log(2.0) / max(n, 1)
    if private_n > 0:
        cnt = rng.poisson(private_n * G, n)
        for i in range(n):
            if cnt[i]:
                out[i, rng.integers(0, G, cnt[i])] = ord("N")
    if lower > 0:
        hit = (rng.random(out.shape) < lower) & (out >= 65) & (out <= 90)
        out[hit] += 32
    return out


def fasta_bytes(names, seqs, wrap=60, eol=b"\n", comments=None):
    """Serialise to multi-FASTA text.  wrap=0 -> one line per record; wrap may be a list (per record)."""
    parts = []
    for i, name in enumerate(names):
        w = wrap[i % len(wrap)] if isinstance(wrap, (list, tuple)) else wrap
        head = b">" + name.encode()
        if comments:
            head += b" " + comments[i % len(comments)].encode()
        parts.append(head + eol)
        row = seqs[i].tobytes()
        if w and w > 0:
            for k in range(0, len(row), w):
                parts.append(row[k:k + w] + eol)
            if len(row) == 0:
                parts.append(eol)
        else:
            parts.append(row + eol)
    return b"".join(parts)


def write_fasta(path, names, seqs, wrap=60, eol=b"\n", comments=None, rows_per_write=256):
    with open(path, "wb") as f:
        for r0 in range(0, len(names), rows_per_write):
            r1 = min(len(names), r0 + rows_per_write)
            f.write(fasta_bytes(names[r0:r1], seqs[r0:r1], wrap, eol, comments))


def c1_alignment(G=20000, L_target=831, seed=1):
    """BASELINE config 1: 4 samples whose snp-sites -c output has exactly L_target columns."""
    n = 4
    rng = _rng(seed)
    ref = BASES[rng.integers(0, 4, G)]
    out = np.repeat(ref[None, :], n, axis=0)
    cols = np.sort(rng.choice(G, size=L_target + 40, replace=False))
    snp_cols, n_cols = cols[:L_target], cols[L_target:]
    for c in snp_cols:
        carriers = rng.integers(1, 2 ** n - 1)
        alt = BASES[(np.searchsorted(BASES, ref[c]) + rng.integers(1, 4)) % 4]
        for i in range(n):
            if carriers >> i & 1:
                out[i, c] = alt
    # a second, different SNP set that is hidden by an N in one sample (must be dropped by -c)
    for c in n_cols:
        out[rng.integers(0, n), c] = ord("N")
        j = rng.integers(0, n)
        if out[j, c] != ord("N"):
            out[j, c] = BASES[(np.searchsorted(BASES, ref[c]) + 1) % 4]
    return out
