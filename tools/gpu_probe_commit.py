#!/usr/bin/env python
"""cycles per stage (8 kind::mxf4 MMAs of 128x240x64 per CTA) with a commit every `every` stages."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from btb_phylo_b200 import _capi
lib = _capi.lib()
for cg in (1, 2):
    for every in (1, 2, 4, 8):
        v = ctypes.c_int64(0)
        rc = lib.btb_query(("probe_cg%d_e%d" % (cg, every)).encode(), ctypes.byref(v))
        print("cg=%d commit every %d stage(s): rc=%d cycles/stage=%.1f %s" % (cg, every, rc, v.value / 100.0, _capi.last_error() if rc else ""), flush=True)
