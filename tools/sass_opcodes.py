#!/usr/bin/env python
"""SASS opcode histogram of the shipped library: the tensor-core, TMA, tensor-memory, mbarrier and IDP4A mnemonics
per kernel (the evidence profiles/r02_sass_opcodes.txt holds).

    python tools/sass_opcodes.py [btb_phylo_b200/libbtbphylo.so] > profiles/r02_sass_opcodes.txt
"""
import collections
import datetime
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = re.compile(r"^(UTC|UTMA|UBLKCP|LDTM|STTM|SYNCS|IDP|ELECT)")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "btb_phylo_b200", "libbtbphylo.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    names = {}
    per = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m and cur and WANT.match(m.group(1)):
            per[cur][m.group(1)] += 1
    if per:
        dem = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
        names = dict(zip(per, (re.sub(r"\(.*", "", d).replace("void ", "") for d in dem)))
    print("# SASS opcode histogram of btb_phylo_b200/libbtbphylo.so (cuobjdump -sass, sm_100a): tensor-core (UTC*MMA), TMA (UTMALDG,")
    print("# UBLKCP), tensor-memory (LDTM, STTM), mbarrier (SYNCS), IDP4A and ELECT mnemonics per kernel.  %s  (tools/sass_opcodes.py)"
          % datetime.date.today().isoformat())
    for fn in sorted(per, key=lambda f: names.get(f, f)):
        for op, c in sorted(per[fn].items()):
            print("%6d  %-34s %s" % (c, op, names.get(fn, fn)))


if __name__ == "__main__":
    main()
