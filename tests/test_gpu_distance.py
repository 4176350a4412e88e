"""GPU (-m gpu): the distance stage (k4) through the C-ABI against the CPU oracle, bit-exact.
Both engines (POPC bit-planes; tcgen05: kind::mxf4 on 0/1 nibble codes with CTA pairs (cg = 2), single
CTAs (cg = 1) and half-tile units, and the kind::i8 kernels), default and -a / -k modes, ragged sizes, tile-packed
output, rank shares by tiles and by block rows, uint16 / uint32 output, the level-2 host call."""
import ctypes

import numpy as np
import pytest
import torch

from btb_phylo_b200 import multirank, synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

ENGINES = [("popc", 0, 0), ("umma_pair", 1, 2), ("umma_single", 1, 1)]


def to_device(seqs):
    n, L = seqs.shape
    stride = max(16, (L + 15) // 16 * 16)
    host = np.full((n, stride), ord("."), dtype=np.uint8)
    host[:, :L] = seqs
    return torch.from_numpy(host).cuda()


def gpu_matrix(capi, seqs, engine, cg, allchars=False, keepcase=False, out_dtype=None):
    from btb_phylo_b200 import device
    if cg:
        capi.check(capi.lib().btb_set_option(b"umma_fp4_pair", 1 if cg == 2 else 0))
    plan = device.DistancePlan(to_device(seqs), length=seqs.shape[1], allchars=allchars, keepcase=keepcase,
                               engine=engine, out_dtype=out_dtype)
    plan.encode()
    out = plan.alloc_matrix()
    plan.compute(out)
    torch.cuda.synchronize()
    return plan, out[:, : seqs.shape[0]].cpu().numpy().astype(np.int64)


@pytest.mark.parametrize("name,engine,cg", ENGINES)
@pytest.mark.parametrize("n,L", [(1, 1), (2, 7), (4, 831), (255, 33), (256, 128), (257, 129), (700, 1000), (1300, 4097)])
def test_default_mode_bit_exact(capi, oracle, name, engine, cg, n, L):
    seqs = synth.snp_alignment(n, L, seed=n + L, heavy_n=0.15, lower=0.02)
    _, got = gpu_matrix(capi, seqs, engine, cg)
    want = oracle.snp_dists_rows(seqs, threads=8).astype(np.int64)
    assert (got == want).all()


@pytest.mark.parametrize("name,engine,cg", ENGINES)
@pytest.mark.parametrize("allchars,keepcase", [(True, False), (False, True), (True, True)])
def test_allchars_and_keepcase_modes(capi, oracle, name, engine, cg, allchars, keepcase):
    seqs = synth.snp_alignment(530, 2049, seed=21, heavy_n=0.25, iupac=0.01, lower=0.03)
    seqs[5, 100:140] = ord(".")           # a literal '.' never counts, even with -a
    plan, got = gpu_matrix(capi, seqs, engine, cg, allchars=allchars, keepcase=keepcase)
    want = oracle.snp_dists_rows(seqs, allchars=allchars, keepcase=keepcase, threads=8).astype(np.int64)
    assert (got == want).all()
    if allchars:
        assert plan.sigma > 4


@pytest.mark.parametrize("name,engine,cg", ENGINES)
@pytest.mark.parametrize("n,L,allchars", [(3, 5, False), (300, 1000, False), (513, 4099, False), (513, 4099, True), (1500, 30000, False)])
def test_dense_alignments_use_compact_encodings(capi, oracle, name, engine, cg, n, L, allchars):
    """No ignored byte anywhere (what snp-sites -c always writes): UMMA switches to the +-1 simplex
    operand (d = (3L - dot)/4), POPC drops the validity plane.  Same integers."""
    seqs = synth.snp_alignment(n, L, seed=n * 7 + L, lower=0.0 if allchars else 0.02)
    plan, got = gpu_matrix(capi, seqs, engine, cg, allchars=allchars)
    assert plan.sigma == capi.SIGMA_DENSE4
    assert (got == oracle.snp_dists_rows(seqs, allchars=allchars, threads=8)).all()
    # one ignored byte switches the general encodings back on
    seqs[n // 2, L // 2] = ord("N")
    plan, got = gpu_matrix(capi, seqs, engine, cg, allchars=False)
    assert plan.sigma == 4
    assert (got == oracle.snp_dists_rows(seqs, threads=8)).all()


def test_kernel_variants_agree(capi, oracle):
    """Every tuning variant of the tensor-core engine gives the same integers: dense codes
    (one-hot / +-1 simplex / 0-1 tetrahedron with row sums), CTA pairs or single CTAs, lockstep on or off."""
    lib = capi.lib()
    dense = synth.snp_alignment(700, 3001, seed=12)
    general = synth.snp_alignment(700, 3001, seed=13, heavy_n=0.1)
    want_d = oracle.snp_dists_rows(dense, threads=8)
    want_g = oracle.snp_dists_rows(general, threads=8)
    try:
        for fp4 in (1, 0):                                   # kind::mxf4 on packed nibbles / kind::i8
            capi.check(lib.btb_set_option(b"umma_fp4", fp4))
            for code in (0, 1, 2):
                capi.check(lib.btb_set_option(b"umma_dense_simplex", code))
                for cg in (1, 2):
                    plan, got = gpu_matrix(capi, dense, 1, cg)
                    assert plan.sigma == capi.SIGMA_DENSE4 and (got == want_d).all(), (fp4, code, cg)
        big = synth.snp_alignment(6400, 777, seed=14, heavy_n=0.05)    # 325 tiles >= 2 per SM: the full-tile kernels engage
        iupac = synth.snp_alignment(900, 2500, seed=15, heavy_n=0.2, lower=0.01)
        iupac[::7, ::5] = ord("R"); iupac[::11, ::3] = ord("y"); iupac[::13, ::17] = ord("-"); iupac[5::9, 1::4] = ord("K")
        want_big = oracle.snp_dists_rows(big, threads=8)
        big_dense = synth.snp_alignment(6400, 777, seed=14)
        want_big_dense = oracle.snp_dists_rows(big_dense, threads=8)
        for tile_mode in (0, 1):
            capi.check(lib.btb_set_option(b"umma_tile_mode", tile_mode))
            for fp4 in (1, 0):
                capi.check(lib.btb_set_option(b"umma_fp4", fp4))
                plan, got = gpu_matrix(capi, big_dense, 1, 1)
                assert plan.sigma == capi.SIGMA_DENSE4 and (got == want_big_dense).all(), (tile_mode, fp4)
        capi.check(lib.btb_set_option(b"umma_fp4", 1))
        want_iupac = oracle.snp_dists_rows(iupac, allchars=True, threads=8)
        for general_fp4 in (1, 0):                           # V - X nibble planes on kind::mxf4 / int8 one-hot x complement
            capi.check(lib.btb_set_option(b"umma_fp4_general", general_fp4))
            for tile_mode in (0, 1):                         # half-tile units / one CTA per full tile
                capi.check(lib.btb_set_option(b"umma_tile_mode", tile_mode))
                plan, got = gpu_matrix(capi, big, 1, 1)
                assert plan.sigma == 4 and (got == want_big).all(), (general_fp4, tile_mode)
                plan, got = gpu_matrix(capi, iupac, 1, 1, allchars=True)
                assert plan.sigma > 4 and (got == want_iupac).all(), (general_fp4, tile_mode, plan.sigma)
                _, got = gpu_matrix(capi, general, 1, 1)
                assert (got == want_g).all(), (general_fp4, tile_mode)
        capi.check(lib.btb_set_option(b"umma_tile_mode", 1))
        # tuning knobs of the FP4 full-tile kernels: CTA pairs (cta_group::2) or single CTAs, loose / no lockstep
        for key, values, reset in ((b"umma_fp4_pair", (0, 1), None), (b"umma_lockstep", (3, 0, 1), 1)):
            before = ctypes.c_int64(0)
            lib.btb_query(key, ctypes.byref(before))
            for v in values:
                capi.check(lib.btb_set_option(key, v))
                _, got = gpu_matrix(capi, big_dense, 1, 1)
                assert (got == want_big_dense).all(), (key, v)
                _, got = gpu_matrix(capi, big, 1, 1)
                assert (got == want_big).all(), (key, v)
            capi.check(lib.btb_set_option(key, reset if reset is not None else before.value))
    finally:
        capi.check(lib.btb_set_option(b"umma_dense_simplex", 2))
        capi.check(lib.btb_set_option(b"umma_fp4", 1))
        capi.check(lib.btb_set_option(b"umma_fp4_general", 1))
        capi.check(lib.btb_set_option(b"umma_tile_mode", 1))
        capi.check(lib.btb_set_option(b"umma_fp4_pair", 1))


def test_golden_c1_matrix(capi, oracle):
    from tests.conftest import golden
    lines = golden("c1_snps.fas").split(b"\n")
    seqs = np.stack([np.frombuffer(lines[2 * i + 1], dtype=np.uint8) for i in range(4)])
    want = np.array([[int(v) for v in row.split(b",")[1:]] for row in golden("c1_snps.csv").split(b"\n")[1:5]])
    for _, engine, cg in ENGINES:
        _, got = gpu_matrix(capi, seqs, engine, cg)
        assert (got == want).all()


@pytest.mark.parametrize("name,engine,cg", ENGINES)
def test_tile_packed_rank_shares_assemble(capi, oracle, name, engine, cg):
    """world = 3: each rank computes its contiguous tile range into tile-packed storage; the host
    assembly equals the single-GPU matrix and the oracle (multi-GPU = same bytes as 1 GPU)."""
    from btb_phylo_b200 import device
    n, L = 1100, 700
    seqs = synth.snp_alignment(n, L, seed=8, heavy_n=0.1)
    if cg:
        capi.check(capi.lib().btb_set_option(b"umma_fp4_pair", 1 if cg == 2 else 0))
    plan = device.DistancePlan(to_device(seqs), length=L, engine=engine)
    plan.encode()
    shares, packed = [], []
    for r in range(3):
        lo, hi = multirank.tile_share(n, r, 3)
        tiles = plan.alloc_tiles(lo, hi)
        plan.compute(tiles, lo, hi, packed=True)
        torch.cuda.synchronize()
        shares.append((lo, hi))
        packed.append(tiles.cpu().numpy().astype(np.uint32))
    mat = multirank.assemble(n, shares, packed)
    assert (mat == oracle.snp_dists_rows(seqs, threads=8)).all()


@pytest.mark.parametrize("name,engine,cg", ENGINES)
@pytest.mark.parametrize("n,L,dense,world", [(1100, 700, False, 3), (2900, 1200, True, 2), (2900, 1200, True, 5), (300, 90, True, 4)])
def test_block_row_shares_cover_the_matrix(capi, oracle, name, engine, cg, n, L, dense, world):
    """btb_distance_rows: each rank writes the cells (r, c >= r) of its block rows plus their mirrors
    into its own n x n matrix; merged by ownership they equal the oracle, and a rank never writes
    outside what it owns (the rest of its matrix stays zero)."""
    from btb_phylo_b200 import device
    seqs = synth.snp_alignment(n, L, seed=21, heavy_n=0.0 if dense else 0.1, lower=0.0 if dense else 0.01)
    if dense:
        seqs = np.where(np.isin(seqs, list(b"ACGT")), seqs, ord("A")).astype(np.uint8)
    if cg:
        capi.check(capi.lib().btb_set_option(b"umma_fp4_pair", 1 if cg == 2 else 0))
    plan = device.DistancePlan(to_device(seqs), length=L, engine=engine)
    assert (plan.sigma == capi.SIGMA_DENSE4) == dense
    plan.encode()
    want = oracle.snp_dists_rows(seqs, threads=8)
    shares, mats = [], []
    for r in range(world):
        lo, hi = device.row_share(n, r, world)
        out = plan.alloc_matrix()
        plan.compute_rows(out, lo, hi)
        torch.cuda.synchronize()
        m = out[:, :n].cpu().numpy().astype(np.uint32)
        r0, r1 = min(n, lo * 256), min(n, hi * 256)
        own = np.zeros((n, n), dtype=bool)
        own[r0:r1, r0:] = True
        own[r1:, r0:r1] = True
        assert (m[own] == want[own]).all()
        assert not m[~own].any()
        shares.append((lo, hi))
        mats.append(m)
    assert (multirank.assemble_rows(n, shares, mats) == want).all()


def test_block_rows_without_mirror_and_arbitrary_ranges(capi, oracle):
    from btb_phylo_b200 import device
    n, L = 1600, 500
    seqs = synth.snp_alignment(n, L, seed=22, heavy_n=0.0)
    seqs = np.where(np.isin(seqs, list(b"ACGT")), seqs, ord("C")).astype(np.uint8)
    want = oracle.snp_dists_rows(seqs, threads=8)
    for _, engine, cg in ENGINES:
        if cg:
            capi.check(capi.lib().btb_set_option(b"umma_fp4_pair", 1 if cg == 2 else 0))
        plan = device.DistancePlan(to_device(seqs), length=L, engine=engine)
        plan.encode()
        for lo, hi in [(0, 7), (2, 5), (6, 7), (3, 3)]:
            out = plan.alloc_matrix()
            plan.compute_rows(out, lo, hi, mirror=False)
            torch.cuda.synchronize()
            m = out[:, :n].cpu().numpy().astype(np.uint32)
            r0, r1 = min(n, lo * 256), min(n, hi * 256)
            upper = np.triu(np.ones((n, n), dtype=bool))
            own = np.zeros((n, n), dtype=bool)
            own[r0:r1, r0:] = True
            lower = own & ~upper                                   # lower halves of diagonal tiles: engine's choice
            own &= upper
            assert (m[own] == want[own]).all()
            assert ((m[lower] == want[lower]) | (m[lower] == 0)).all()
            outside = np.ones((n, n), dtype=bool)
            outside[r0:r1, r0:] = False
            assert not m[outside].any()


def test_uint32_output_and_long_rows(capi, oracle):
    seqs = synth.snp_alignment(300, 70001, seed=2, heavy_n=0.05)
    for _, engine, cg in ENGINES:
        plan, got = gpu_matrix(capi, seqs, engine, cg)
        assert plan.out_dtype == torch.uint32
        assert (got == oracle.snp_dists_rows(seqs, threads=8)).all()


def test_uint32_output_through_the_full_tile_kernel(capi, oracle):
    """Rows of 66,000 sites need uint32 cells; 6,400 sequences engage the single-CTA full-tile kernel
    (the CTA-pair kernel writes uint16 only).  Checked against the POPC engine everywhere and against
    the oracle on a few rows."""
    n, L = 6400, 66000
    seqs = synth.snp_alignment(n, L, seed=23)
    plan, a = gpu_matrix(capi, seqs, 1, 1)
    assert plan.out_dtype == torch.uint32 and plan.sigma == capi.SIGMA_DENSE4
    _, b = gpu_matrix(capi, seqs, 0, 0)
    assert (a == b).all() and (a == a.T).all() and (np.diag(a) == 0).all()
    for r in (0, 3333, n - 1):
        assert (a[r] == oracle.snp_dists_rows(seqs, threads=8, row0=r, row1=r + 1)[0]).all()


def test_level2_host_call(capi, oracle):
    from btb_phylo_b200 import device
    seqs = synth.snp_alignment(777, 831, seed=5, heavy_n=0.2, lower=0.01)
    want = oracle.snp_dists_rows(seqs, threads=8)
    for engine in (capi.ENGINE_AUTO, capi.ENGINE_POPC, capi.ENGINE_UMMA):
        assert (device.distance_matrix_host(seqs, engine=engine) == want).all()
        assert (device.distance_matrix_host(seqs, engine=engine, dtype=np.uint16) == want).all()
    # two ranks fill disjoint tile sets of the same host matrix
    out = np.zeros((777, 777), dtype=np.uint32)
    device.distance_matrix_host(seqs, rank=0, world=2, out=out)
    assert not (out == want).all()
    device.distance_matrix_host(seqs, rank=1, world=2, out=out)
    assert (out == want).all()


@pytest.mark.parametrize("dense", [True, False])
@pytest.mark.parametrize("world", [2, 3])
def test_partial_encode_and_weighted_shares(capi, oracle, dense, world):
    """A rank encodes only the rows from (one block before) its first block row on -- everything before stays
    garbage -- and the block-row ranges are cut with the encode cost counted in: the assembled matrix is exact."""
    from btb_phylo_b200 import device
    n, L = 5000, 700
    seqs = synth.snp_alignment(n, L, seed=61, heavy_n=0.0 if dense else 0.1)
    want = oracle.snp_dists_rows(seqs, threads=8).astype(np.int64)
    shares, mats = [], []
    for rank in range(world):
        lo, hi = device.row_share_weighted(n, rank, world, 40.0)
        assert (lo, hi) == multirank.row_share_weighted(n, rank, world, 40.0)
        plan = device.DistancePlan(to_device(seqs), length=L, engine=capi.ENGINE_UMMA)
        assert (plan.sigma == capi.SIGMA_DENSE4) == dense
        plan.operands.fill_(0xA7)                                     # rows that are not encoded hold garbage
        plan.encode(max(0, lo * 256 - 256), n)
        out = plan.alloc_matrix()
        plan.compute_rows(out, lo, hi)
        torch.cuda.synchronize()
        shares.append((lo, hi))
        mats.append(out[:, :n].cpu().numpy().astype(np.int64))
    assert shares[0][0] == 0 and shares[-1][1] == multirank.tiles_per_dim(n)
    assert (multirank.assemble_rows(n, shares, mats) == want).all()


def test_level2_upper_triangle_only(capi, oracle):
    """btb_dist_matrix_host_upper: every cell on or above the diagonal, nothing below it is touched --
    one GPU (band by band) and the two shares of a 2-rank job, dense and general alphabets, both engines."""
    from btb_phylo_b200 import device
    for seqs in (synth.snp_alignment(4700, 900, seed=52), synth.snp_alignment(2300, 1300, seed=53, heavy_n=0.1)):
        n = seqs.shape[0]
        want = np.triu(oracle.snp_dists_rows(seqs, threads=8))
        for engine in (capi.ENGINE_UMMA, capi.ENGINE_POPC):
            out = np.full((n, n), 0xFFFF, dtype=np.uint16)
            device.distance_matrix_host(seqs, engine=engine, out=out, upper_only=True)
            assert (np.triu(out) == want).all() and (np.tril(out, -1 * 256 * 8) == np.tril(np.full((n, n), 0xFFFF, np.uint16), -1 * 256 * 8)).all()
        out = np.zeros((n, n), dtype=np.uint32)
        device.distance_matrix_host(seqs, rank=0, world=2, out=out, upper_only=True)
        device.distance_matrix_host(seqs, rank=1, world=2, out=out, upper_only=True)
        assert (np.triu(out) == want).all()


def test_level2_streamed_path(capi, oracle):
    """n >= 16 block rows, default mode, dense input: the level-2 call uploads / computes / downloads
    band by band in reverse order; a non-dense band sends the call down the general path.  Same
    matrix either way, and with the streaming switched off."""
    from btb_phylo_b200 import device
    lib = capi.lib()

    def last():
        v = ctypes.c_int64(-2)
        capi.check(lib.btb_query(b"host_stream_last", ctypes.byref(v)))
        return v.value

    n, L = 4500, 333
    dense = synth.snp_alignment(n, L, seed=17)
    dense[::5, ::3] |= 0x20                                  # lower case is still dense in default mode
    want = oracle.snp_dists_rows(dense, threads=8)
    assert (device.distance_matrix_host(dense) == want).all() and last() == 0       # off by default
    try:
        capi.check(lib.btb_set_option(b"host_stream", 1))
        for rows in (8, 16, 64):
            capi.check(lib.btb_set_option(b"host_stream_rows", rows))
            for dtype in (np.uint16, np.uint32):
                out = np.full((n, n + 8), 0xFFFF, dtype=dtype)[:, :n]   # a padded host matrix (ld > n)
                device.distance_matrix_host(dense, out=out)
                bad = np.argwhere(out != want)
                assert last() == 1 and len(bad) == 0, (rows, dtype, len(bad), bad[:3], bad[-3:], out[tuple(bad[0])], want[tuple(bad[0])])
        capi.check(lib.btb_set_option(b"host_stream_rows", 8))
        for bad_row in (n - 1, 2100, 0):                     # the first / a middle / the last band to be checked
            mixed = dense.copy()
            mixed[bad_row, 7] = ord("N")
            got = device.distance_matrix_host(mixed)
            assert last() == 0 and (got == oracle.snp_dists_rows(mixed, threads=8)).all(), bad_row
        assert (device.distance_matrix_host(dense) == want).all() and last() == 1
        assert (device.distance_matrix_host(dense, allchars=True) == oracle.snp_dists_rows(dense, allchars=True, threads=8)).all() and last() == 0
    finally:
        capi.check(lib.btb_set_option(b"host_stream", 0))
        capi.check(lib.btb_set_option(b"host_stream_rows", 64))


def test_properties_at_scale(capi, oracle):
    """8192 x 30000 (BASELINE config 4 shape, cut to 8192 rows): symmetry, zero diagonal, engines
    agree everywhere, sampled rows equal the oracle."""
    n, L = 8192, 30000
    seqs = synth.snp_alignment(n, L, seed=4)
    _, a = gpu_matrix(capi, seqs, 1, 2)
    _, b = gpu_matrix(capi, seqs, 0, 0)
    assert (a == a.T).all() and (np.diag(a) == 0).all() and (a == b).all()
    for r in (0, 1, 4097, 8191):
        assert (a[r] == oracle.snp_dists_rows(seqs, threads=8, row0=r, row1=r + 1)[0]).all()
    assert a.max() < L


def test_allchars_heavy_gap_content_at_scale(capi, oracle):
    """BASELINE config 5 shape cut to 6,000 rows: 60,000 sites, 25 % N / gap, IUPAC codes, a few
    lower-case bases, snp-dists -a.  AUTO engine; sampled rows equal the oracle, engines agree."""
    from btb_phylo_b200 import device
    n, L = 6000, 60000
    seqs = synth.snp_alignment(n, L, seed=5, heavy_n=0.25, iupac=0.001, lower=0.001)
    plan, a = gpu_matrix(capi, seqs, 1, 1, allchars=True)
    assert plan.sigma >= 6
    _, b = gpu_matrix(capi, seqs, 0, 0, allchars=True)
    assert (a == b).all() and (a == a.T).all() and (np.diag(a) == 0).all()
    for r in (0, 2999, 5999):
        assert (a[r] == oracle.snp_dists_rows(seqs, allchars=True, threads=8, row0=r, row1=r + 1)[0]).all()
    # default mode on the same bytes (N and gaps ignored)
    _, c = gpu_matrix(capi, seqs, 1, 1, allchars=False)
    assert (c[17] == oracle.snp_dists_rows(seqs, threads=8, row0=17, row1=18)[0]).all()


def _rows_from_base(base, n, mutations, alphabet, seed):
    """n rows = `base` with `mutations` random substitutions each (torch, on the GPU)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    L = base.shape[0]
    rows = base[None, :].repeat(n, 1)
    pos = torch.randint(0, L, (n, mutations), generator=g, device="cuda")
    sym = torch.tensor(list(alphabet), dtype=torch.uint8, device="cuda")[torch.randint(0, len(alphabet), (n, mutations), generator=g, device="cuda")]
    rows.scatter_(1, pos, sym)
    return rows


def test_fp4_exact_at_the_dense_guard(capi, oracle):
    """The dense kind::mxf4 path is admitted while 3 * len < 2^24 (common.cuh umma_fp4_dense): fp32
    accumulators must then hold every partial sum exactly.  Rows at the guard whose dot products are as
    large as the code allows: near-identical rows over C/G/T only, majority-zero coding OFF, so every site
    carries two 1-nibbles and dot(i, j) ~ 2 * len = 1.1e7.  Checked against the oracle, bit for bit."""
    from btb_phylo_b200 import device
    L = ((1 << 24) - 512) // 3                       # 5,592,234: 3 * L = 2^24 - 514
    n = 96
    g = torch.Generator(device="cuda")
    g.manual_seed(71)
    base = torch.tensor(list(b"CGT"), dtype=torch.uint8, device="cuda")[torch.randint(0, 3, (L,), generator=g, device="cuda")]
    rows = _rows_from_base(base, n, 300, b"ACGT", 72)
    rows[5] = rows[4]                                # an identical pair: distance 0 out of a dot of ~1.1e7
    stride = (L + 15) // 16 * 16
    seqs = torch.full((n, stride), ord("."), dtype=torch.uint8, device="cuda")
    seqs[:, :L] = rows
    capi.check(capi.lib().btb_set_option(b"umma_major_zero", 0))
    try:
        plan = device.DistancePlan(seqs, length=L, engine=capi.ENGINE_UMMA)
        assert plan.sigma == capi.SIGMA_DENSE4
        assert capi.lib().btb_dist_operand_bytes(capi.ENGINE_UMMA, n, L, plan.sigma) == n * (((3 * L + 1) // 2 + 127) // 128 * 128) + 4 * n   # the nibble layout, not int8
        plan.encode()
        out = plan.alloc_matrix()
        plan.compute(out)
        torch.cuda.synchronize()
    finally:
        capi.check(capi.lib().btb_set_option(b"umma_major_zero", 1))
    got = out[:, :n].cpu().numpy().astype(np.int64)
    want = oracle.snp_dists_rows(rows.cpu().numpy(), threads=16).astype(np.int64)
    assert got[4, 5] == 0 and (got == want).all()
    # one site more and the guard sends the rows to the int8 kernels (int32 accumulators)
    assert capi.lib().btb_dist_operand_bytes(capi.ENGINE_UMMA, n, L + 200, plan.sigma) == n * (((L + 200) * 3 + 127) // 128 * 128) + 4 * n


def test_fp4_exact_at_the_general_guard(capi, oracle):
    """General alphabets (validity + one-hot nibble planes, V.V - X.X with the negate-A bit) are admitted
    while len < 2^24: V.V reaches len.  Near-identical rows of len = 2^24 - 256 with a few N."""
    from btb_phylo_b200 import device
    L = (1 << 24) - 256
    n = 40
    g = torch.Generator(device="cuda")
    g.manual_seed(73)
    base = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")[torch.randint(0, 4, (L,), generator=g, device="cuda")]
    rows = _rows_from_base(base, n, 500, b"ACGTN-", 74)
    rows[7] = rows[3]
    seqs = torch.full((n, L), ord("."), dtype=torch.uint8, device="cuda")       # L is a multiple of 16
    seqs[:, :L] = rows
    plan = device.DistancePlan(seqs, length=L, engine=capi.ENGINE_UMMA)
    assert plan.sigma == 4                            # ignored bytes present: not dense
    assert capi.lib().btb_dist_operand_bytes(capi.ENGINE_UMMA, n, L, plan.sigma) == n * 5 * ((L + 255) // 256 * 128) + 4 * n
    plan.encode()
    out = plan.alloc_matrix()
    plan.compute(out)
    torch.cuda.synchronize()
    got = out[:, :n].cpu().numpy().astype(np.int64)
    want = oracle.snp_dists_rows(rows.cpu().numpy(), threads=16).astype(np.int64)
    assert got[3, 7] == 0 and (got == want).all()
