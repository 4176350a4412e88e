"""Host-side logic of the multi-GPU distance stage: which upper-triangle
tiles a rank owns and how the per-rank tile-packed results become one symmetric matrix for the
CSV writer.  There is no collective on the compute path: operands are replicated, tiles are
independent, results are gathered on the host.  Pure numpy so that it is testable with gloo on
CPU; the tiles themselves come from btb_distance_tiles (out_mode 1) on each rank's GPU."""
import numpy as np

TILE = 256
BAND = 8          # must match kBand in csrc/common.cuh


def tiles_per_dim(n):
    return (n + TILE - 1) // TILE


def tile_count(n):
    t = tiles_per_dim(n)
    return t * (t + 1) // 2


def tile_coords(n, idx):
    """Same row-band order as csrc/common.cuh tile_coords (checked against the library in tests)."""
    T = tiles_per_dim(n)
    r0 = 0
    while r0 < T:
        w = min(BAND, T - r0)
        tri = w * (w + 1) // 2
        if idx < tri:
            for c in range(w):
                if idx < c + 1:
                    return r0 + idx, r0 + c
                idx -= c + 1
        idx -= tri
        rect = (T - r0 - w) * w
        if idx < rect:
            return r0 + idx % w, r0 + w + idx // w
        idx -= rect
        r0 += w
    raise IndexError("tile index out of range")


def tile_share(n, rank, world):
    """Tiles [lo, hi) owned by `rank`: equal counts (+-1), contiguous in band order."""
    t = tile_count(n)
    return t * rank // world, t * (rank + 1) // world


def row_share(n, rank, world):
    """Block rows [lo, hi) of `rank` (same rule as btb_dist_row_share): contiguous ranges holding
    (nearly) equal numbers of upper-triangle tiles -- the unit the multi-GPU paths shard by."""
    T = tiles_per_dim(n)
    total = tile_count(n)

    def boundary(r):
        if r <= 0:
            return 0
        if r >= world:
            return T
        want = total * r // world
        acc = i = 0
        while i < T and acc + (T - i) <= want:
            acc += T - i
            i += 1
        if i % 2 == 1 and i < T and T >= 4 * world:          # even starts: 512-row units of the CTA-pair kernel
            i += 1
        return i
    return boundary(rank), boundary(rank + 1)


def row_share_weighted(n, rank, world, encode_tiles):
    """Same rule as btb_dist_row_share_weighted: contiguous block-row ranges whose  tiles + encode share  are
    (nearly) equal, where a rank starting at block row lo pays encode_tiles * (T - lo) / T for its operands."""
    T = tiles_per_dim(n)

    def cost(a, b):
        return float(b - a) * float(T - a) - float(b - a) * float(b - a - 1) / 2.0 + encode_tiles * float(T - a) / float(T)

    def fits(cap):
        a, bounds = 0, []
        for _ in range(world):
            b = a
            while b < T and cost(a, b + 1) <= cap:
                b += 1
            if b == a and a < T:
                return None
            if b % 2 == 1 and b < T and T >= 4 * world and b - 1 > a:
                b -= 1
            bounds.append(a)
            a = b
        return bounds + [T] if a >= T else None

    lo_c, hi_c = 0.0, cost(0, T)
    for _ in range(60):
        mid = 0.5 * (lo_c + hi_c)
        if fits(mid) is not None:
            hi_c = mid
        else:
            lo_c = mid
    bounds = fits(hi_c)
    return bounds[rank], (T if rank + 1 == world else bounds[rank + 1])


def assemble_rows(n, shares, mats, out=None):
    """shares: list of block-row ranges (lo, hi); mats: per-rank n x n matrices in which the rank has
    written the cells (r, c >= r) of its rows and their mirrors.  Returns the merged matrix."""
    if out is None:
        out = np.zeros((n, n), dtype=mats[0].dtype)
    for (lo, hi), m in zip(shares, mats):
        r0, r1 = min(n, lo * TILE), min(n, hi * TILE)
        out[r0:r1, r0:] = m[r0:r1, r0:n]
        out[r1:, r0:r1] = m[r1:n, r0:r1]
    return out


def assemble(n, shares, packed, out=None):
    """shares: list of (lo, hi); packed: list of arrays [(hi-lo), TILE, TILE] in the same order.
    Returns the full symmetric n x n matrix (both triangles)."""
    dtype = packed[0].dtype
    if out is None:
        out = np.zeros((n, n), dtype=dtype)
    for (lo, hi), tiles in zip(shares, packed):
        for t in range(lo, hi):
            I, J = tile_coords(n, t)
            r0, c0 = I * TILE, J * TILE
            r1, c1 = min(n, r0 + TILE), min(n, c0 + TILE)
            blk = tiles[t - lo][: r1 - r0, : c1 - c0]
            out[r0:r1, c0:c1] = blk
            if I != J:
                out[c0:c1, r0:r1] = blk.T
    return out
