"""Buffer plumbing between PyTorch and the C-ABI's device-pointer level (include/btb_phylo.h,
"Level 3").  torch allocates HBM, owns the streams and (in bench.py) the process group; every
kernel is launched by libbtbphylo.so on the stream passed in.  Used by bench.py and the tests --
the reference-facing entry points are in phylogeny.py."""
import ctypes

import numpy as np
import torch

from . import _capi


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def tile_count(n):
    return int(_capi.lib().btb_dist_tile_count(n))


def tile_coords(n, tile):
    i, j = ctypes.c_int32(0), ctypes.c_int32(0)
    _capi.check(_capi.lib().btb_dist_tile_coords(n, tile, ctypes.byref(i), ctypes.byref(j)), "btb_dist_tile_coords")
    return i.value, j.value


def tile_share(n, rank, world):
    """Tiles [lo, hi) of the upper triangle owned by `rank` (contiguous in the band order)."""
    t = tile_count(n)
    return t * rank // world, t * (rank + 1) // world


def block_rows(n):
    return int(_capi.lib().btb_dist_block_rows(n))


def row_share(n, rank, world):
    """Block rows [lo, hi) of `rank`: contiguous, (nearly) equal numbers of upper-triangle tiles."""
    lo, hi = ctypes.c_int64(0), ctypes.c_int64(0)
    _capi.check(_capi.lib().btb_dist_row_share(n, rank, world, ctypes.byref(lo), ctypes.byref(hi)), "btb_dist_row_share")
    return lo.value, hi.value


def row_share_weighted(n, rank, world, encode_tiles):
    """row_share with the operand encode counted in (btb_dist_row_share_weighted)."""
    lo, hi = ctypes.c_int64(0), ctypes.c_int64(0)
    _capi.check(_capi.lib().btb_dist_row_share_weighted(n, rank, world, float(encode_tiles), ctypes.byref(lo), ctypes.byref(hi)),
                "btb_dist_row_share_weighted")
    return lo.value, hi.value


class DistancePlan:
    """Operands + output of the distance stage for one alignment resident in HBM.

    seqs : uint8 cuda tensor [n, stride] (row i = sequence i, `length` valid bytes)
    """

    def __init__(self, seqs, length=None, allchars=False, keepcase=False, engine=_capi.ENGINE_UMMA,
                 out_dtype=None):
        assert seqs.is_cuda and seqs.dtype == torch.uint8 and seqs.dim() == 2 and seqs.stride(1) == 1
        self.lib = _capi.lib()
        self.seqs = seqs
        self.n = seqs.shape[0]
        self.len = seqs.shape[1] if length is None else int(length)
        self.row_stride = seqs.stride(0)
        self.engine = engine
        table = (ctypes.c_uint8 * 256)()
        sigma = ctypes.c_int(0)
        _capi.check(self.lib.btb_dist_alphabet(_ptr(seqs), self.n, self.len, self.row_stride, int(allchars),
                                               int(keepcase), table, ctypes.byref(sigma), _stream()),
                    "btb_dist_alphabet")
        self.table, self.sigma = table, sigma.value
        nbytes = self.lib.btb_dist_operand_bytes(engine, self.n, self.len, self.sigma)
        self.operands = torch.empty(max(16, nbytes), dtype=torch.uint8, device=seqs.device)
        if out_dtype is None:
            out_dtype = torch.uint16 if self.len < 65536 else torch.uint32
        self.out_dtype = out_dtype
        self.elem = 2 if out_dtype == torch.uint16 else 4
        self.tiles = tile_count(self.n)

    def encode(self, row0=0, row1=None):
        """All rows, or only the rows [row0, row1) (what a rank whose block rows start later needs)."""
        row1 = self.n if row1 is None else row1
        if row0 == 0 and row1 == self.n:
            _capi.check(self.lib.btb_dist_encode(self.engine, _ptr(self.seqs), self.n, self.len, self.row_stride,
                                                 self.table, self.sigma, _ptr(self.operands), _stream()),
                        "btb_dist_encode")
        else:
            _capi.check(self.lib.btb_dist_encode_rows(self.engine, _ptr(self.seqs), self.n, self.len, self.row_stride,
                                                      self.table, self.sigma, _ptr(self.operands), row0, row1, _stream()),
                        "btb_dist_encode_rows")

    def alloc_matrix(self):
        ld = (self.n + 7) // 8 * 8
        return torch.zeros((self.n, ld), dtype=self.out_dtype, device=self.seqs.device)

    def alloc_tiles(self, lo, hi):
        return torch.zeros((max(hi - lo, 1), _capi.TILE, _capi.TILE), dtype=self.out_dtype, device=self.seqs.device)

    def compute(self, out, lo=0, hi=None, packed=False, mirror=True):
        hi = self.tiles if hi is None else hi
        ld = 0 if packed else out.stride(0)
        _capi.check(self.lib.btb_distance_tiles(self.engine, _ptr(self.operands), self.n, self.len, self.sigma,
                                                lo, hi, _ptr(out), self.elem, ld, 1 if packed else 0,
                                                1 if mirror else 0, _stream()),
                    "btb_distance_tiles")

    def compute_rows(self, out, row_lo, row_hi, mirror=True):
        """Every cell (r, c >= r) of the block rows [row_lo, row_hi) (+ mirrors) into the n x n matrix `out`."""
        _capi.check(self.lib.btb_distance_rows(self.engine, _ptr(self.operands), self.n, self.len, self.sigma,
                                               row_lo, row_hi, _ptr(out), self.elem, out.stride(0),
                                               1 if mirror else 0, _stream()),
                    "btb_distance_rows")


def distance_matrix_host(seqs, allchars=False, keepcase=False, engine=_capi.ENGINE_AUTO, device=0,
                         rank=0, world=1, out=None, dtype=np.uint32, upper_only=False):
    """Level-2 call on numpy buffers: n x L uint8 -> n x n distances (btb_dist_matrix_host; with
    upper_only only the cells (r, c >= r) are written: btb_dist_matrix_host_upper)."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    n, L = seqs.shape
    if out is None:
        out = np.zeros((n, n), dtype=dtype)
    elem = out.dtype.itemsize
    fn = _capi.lib().btb_dist_matrix_host_upper if upper_only else _capi.lib().btb_dist_matrix_host
    _capi.check(fn(seqs.ctypes.data, n, L, seqs.strides[0], int(allchars), int(keepcase), engine, device, rank, world,
                   out.ctypes.data, elem, out.strides[0] // elem), "btb_dist_matrix_host")
    return out
