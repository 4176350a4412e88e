#!/usr/bin/env python
"""One pack_planes launch on synthetic 60-column FASTA text, for timing/profiling: python tools/gpu_pack.py [records] [G]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from btb_phylo_b200 import _capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
G = int(sys.argv[2]) if len(sys.argv) > 2 else 4349904
W = 60
lib = _capi.lib()
g = torch.Generator(device="cuda"); g.manual_seed(1)
lines = (G + W - 1) // W
rec_len = 32 + G + lines                       # 32-byte pseudo header keeps records mutually unaligned
acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
text = torch.full((n * rec_len + 64,), 10, dtype=torch.uint8, device="cuda")
col = torch.arange(G, device="cuda")
pos = col + col // W
offs = []
for i in range(n):
    base = i * rec_len + 32 - (i % 7)
    offs.append(base)
    text[base + pos] = acgt[torch.randint(0, 4, (G,), generator=g, device="cuda")]
d_off = torch.tensor(offs, dtype=torch.int64, device="cuda")
d_lw = torch.full((n,), W, dtype=torch.int32, device="cuda")
d_el = torch.ones(n, dtype=torch.int32, device="cuda")
words = lib.btb_plane_words(G)
planes = torch.zeros(3 * n * words, dtype=torch.int32, device="cuda")
flags = torch.zeros(n, dtype=torch.int32, device="cuda")
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
run = lambda: _capi.check(lib.btb_pack_planes(P(text), n * rec_len, P(d_off), P(d_lw), P(d_el), n, 0, n, G, P(planes), P(flags), st))
for _ in range(3):
    run()
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
byt = n * (G + lines) + 3 * n * words * 4
print("pack: options=%s records=%d G=%d ms=%.3f GB/s=%.0f flags=%d" % (os.environ.get("BTB_OPTIONS", ""), n, G, best, byt / best / 1e6, int(flags.sum().item())))
