#!/usr/bin/env python
"""Level-2 call (btb_dist_matrix_host, pinned host buffers) at C4: streamed band height sweep."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from btb_phylo_b200 import _capi
import bench

lib = _capi.lib()
n, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (50000, 30000)
d = bench.make_alignment_cuda(n, L, torch.device("cuda", 0))
h_seqs = torch.empty((n, d.shape[1]), dtype=torch.uint8, pin_memory=True); h_seqs.copy_(d); del d
ld = (n + 7) // 8 * 8
h_out = torch.zeros((n, ld), dtype=torch.uint16, pin_memory=True)
ref = None
for key, val in [(b"host_stream", 0)] + [(b"host_stream_rows", r) for r in (8, 16, 24, 32, 48, 64, 96)] + [(b"host_stream", 0)]:
    _capi.check(lib.btb_set_option(b"host_stream", 1))
    _capi.check(lib.btb_set_option(key, val))
    ts = []
    for rep in range(4):
        h_out.zero_() if rep == 0 else None
        t0 = time.perf_counter()
        _capi.check(lib.btb_dist_matrix_host(ctypes.c_void_p(h_seqs.data_ptr()), n, L, h_seqs.stride(0), 0, 0, 1, 0, 0, 1,
                                             ctypes.c_void_p(h_out.data_ptr()), 2, ld), "l2")
        ts.append(time.perf_counter() - t0)
    chk = int(h_out[::499, ::503].to(torch.int64).sum().item()) + int(h_out[n - 1, :n].to(torch.int64).sum().item())
    ref = chk if ref is None else ref
    print("%s=%d s=%s same=%s" % (key.decode(), val, ["%.4f" % t for t in ts], chk == ref), flush=True)
