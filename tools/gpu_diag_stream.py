#!/usr/bin/env python
"""Diagnostic: where does the streamed level-2 path differ from the plain one?  argv[1] = pre-call kind"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from btb_phylo_b200 import _capi, device, synth
lib = _capi.lib()
pre = sys.argv[1] if len(sys.argv) > 1 else "none"
small = synth.snp_alignment(777, 831, seed=5, heavy_n=0.2, lower=0.01)
small_dense = synth.snp_alignment(777, 831, seed=5)
if pre == "u16":
    device.distance_matrix_host(small, engine=1, dtype=np.uint16)
elif pre == "u32":
    device.distance_matrix_host(small, engine=1, dtype=np.uint32)
elif pre == "popc16":
    device.distance_matrix_host(small, engine=0, dtype=np.uint16)
elif pre == "dense16":
    device.distance_matrix_host(small_dense, engine=1, dtype=np.uint16)
elif pre == "dense32":
    device.distance_matrix_host(small_dense, engine=1, dtype=np.uint32)
n, L = 4500, 333
dense = synth.snp_alignment(n, L, seed=17)
dense[::5, ::3] |= 0x20
want = device.distance_matrix_host(dense)
_capi.check(lib.btb_set_option(b"host_stream", 1))
for rows in (8, 16):
    _capi.check(lib.btb_set_option(b"host_stream_rows", rows))
    for dtype in (np.uint16, np.uint32, np.uint16):
        out = np.full((n, n + 8), 0xFFFF, dtype=dtype)[:, :n]
        device.distance_matrix_host(dense, out=out)
        bad = np.argwhere(out != want)
        v = ctypes.c_int64(-2); lib.btb_query(b"host_stream_last", ctypes.byref(v))
        print(pre, "rows", rows, dtype.__name__, "streamed", v.value, "mismatches", len(bad),
              "" if not len(bad) else "rows %d..%d cols %d..%d first %s got %s want %s nonzero_got %d" % (
                  bad[:, 0].min(), bad[:, 0].max(), bad[:, 1].min(), bad[:, 1].max(), bad[0], out[tuple(bad[0])], want[tuple(bad[0])],
                  int((out[bad[:, 0], bad[:, 1]] != 0).sum())), flush=True)
