#!/bin/bash
# Pins the CPU oracle -- and with it every "byte-identical" claim of this repository -- against the REAL
# tools the reference installs (install/install-snp-sites.sh: Ubuntu's snp-sites; install/install-snp-dists.sh:
# tseemann/snp-dists).  Neither is available offline where this repository was built, so the golden files
# under tests/golden/ were produced by the oracle itself (oracle/*.c, a restatement of the published
# behaviour) and parity is "unpinned".  With the two binaries on PATH this script closes that gap in a minute:
#
#   1. every golden input is run through the real snp-sites -c / snp-dists and the result compared with the
#      committed golden output (produced by the oracle);
#   2. a batch of random alignments (lower case, N, -, IUPAC, literal '.', CRLF, wrapped lines, -a / -k / -m /
#      -b modes) is run through both the real tools and the oracle executables, and diffed byte for byte;
#   3. on success it writes tests/golden/PINNED with the tool versions, so the provenance is recorded.
#
#   usage: tools/pin_against_upstream.sh [--cases 200]
# Exit status 0 = the oracle (hence the GPU path, which tests/ hold equal to it) matches the real tools.
set -u
cd "$(dirname "$0")/.."
CASES=200
[ "${1:-}" = "--cases" ] && CASES=$2
for t in snp-sites snp-dists; do
    if ! real=$(command -v $t) || [ "$(readlink -f "$real")" = "$(readlink -f tools/$t)" ]; then
        echo "real '$t' not found on PATH (the shim tools/$t does not count): nothing pinned" >&2
        exit 2
    fi
done
echo "snp-sites: $(command -v snp-sites)  ($(snp-sites -V 2>&1 | head -1))"
echo "snp-dists: $(command -v snp-dists)  ($(snp-dists -v 2>&1 | head -1))"
make -C oracle -s || exit 1
tmp=$(mktemp -d)
trap 'rm -rf "$tmp"' EXIT
fail=0
check() {   # label, file a, file b
    if cmp -s "$2" "$3"; then echo "  ok      $1"; else echo "  DIFFERS $1"; fail=$((fail + 1)); fi
}

echo "== 1. committed golden vectors"
G=tests/golden
snp-dists -q "$G/kat1_in.fa" > "$tmp/kat1.tsv" 2>/dev/null;            check "KAT-1 snp-dists README example (TSV)" "$tmp/kat1.tsv" "$G/kat1_expected.tsv"
snp-sites -c -o "$tmp/kat2c.fas" "$G/kat2_in.fa" >/dev/null 2>&1;      check "KAT-2 snp-sites -c README example"   "$tmp/kat2c.fas" "$G/kat2_expected_pure.fas"
snp-sites -o "$tmp/kat2.fas" "$G/kat2_in.fa" >/dev/null 2>&1;          check "KAT-2 snp-sites default mode"        "$tmp/kat2.fas" "$G/kat2_expected_default.fas"
snp-sites -c -o "$tmp/kat3.fas" "$G/kat3_multi_fasta.fas" >/dev/null 2>&1; check "KAT-3 snps.fas"                      "$tmp/kat3.fas" "$G/kat3_snps.fas"
snp-dists -c -q "$G/kat3_snps.fas" > "$tmp/kat3.csv" 2>/dev/null;          check "KAT-3 snps.csv"                      "$tmp/kat3.csv" "$G/kat3_snps.csv"
snp-sites -c -o "$tmp/c1.fas" "$G/c1_multi_fasta.fas" >/dev/null 2>&1;     check "C1 (4 x 831) snps.fas"               "$tmp/c1.fas" "$G/c1_snps.fas"
snp-dists -c -q "$G/c1_snps.fas" > "$tmp/c1.csv" 2>/dev/null;              check "C1 (4 x 831) snps.csv"               "$tmp/c1.csv" "$G/c1_snps.csv"
snp-dists -q -m -c "$G/kat1_in.fa" > "$tmp/kat1m.csv" 2>/dev/null;         check "KAT-1 molten CSV"                    "$tmp/kat1m.csv" "$G/kat1_expected_molten.csv"
snp-sites -c -o "$tmp/edge_p.fas" "$G/edge_sites_in.fa" >/dev/null 2>&1;   check "edge vector snp-sites -c (CRLF, wrap, comments)" "$tmp/edge_p.fas" "$G/edge_sites_expected_pure.fas"
snp-sites -o "$tmp/edge_d.fas" "$G/edge_sites_in.fa" >/dev/null 2>&1;      check "edge vector snp-sites default mode"  "$tmp/edge_d.fas" "$G/edge_sites_expected_default.fas"

echo "== 2. $CASES random alignments, real tools against the oracle executables"
python - "$tmp" "$CASES" <<'PY' || fail=$((fail + 1))
import os, subprocess, sys
import numpy as np
tmp, cases = sys.argv[1], int(sys.argv[2])
sys.path.insert(0, os.getcwd())
from btb_phylo_b200 import synth
O = os.path.join("oracle", "_build")
rng = np.random.default_rng(2024)
bad = 0
def run(cmd, out=None):
    with open(out, "wb") if out else open(os.devnull, "wb") as f:
        return subprocess.run(cmd, stdout=f, stderr=subprocess.PIPE).returncode
for k in range(cases):
    n, G = int(rng.integers(2, 40)), int(rng.integers(20, 3000))
    aln = synth.genome_alignment(n, G, int(rng.integers(1, max(2, G // 10))), seed=k, n_frac=float(rng.choice([0, 0.02, 0.2])),
                                 lower=float(rng.choice([0, 0.05])))
    if k % 3 == 0:
        junk = np.frombuffer(b"RYKMSWN-?.", dtype=np.uint8)
        hit = rng.random(aln.shape) < 0.01
        aln[hit] = junk[rng.integers(0, len(junk), int(hit.sum()))]
    names = synth.sample_names(n)
    mf = os.path.join(tmp, "in.fas")
    open(mf, "wb").write(synth.fasta_bytes(names, aln, wrap=[60, 70, 0, 33][k % 4], eol=b"\r\n" if k % 5 == 0 else b"\n"))
    for mode in ([["-c"]] + ([[]] if k % 2 else [])):
        a, b = os.path.join(tmp, "a.fas"), os.path.join(tmp, "b.fas")
        for f in (a, b):
            if os.path.exists(f): os.remove(f)
        ra = run(["snp-sites", *mode, "-o", a, mf]); rb = run([os.path.join(O, "snp-sites"), *mode, "-o", b, mf])
        same = (ra != 0) == (rb != 0) and (ra != 0 or open(a, "rb").read() == open(b, "rb").read())
        if not same:
            bad += 1; print("  DIFFERS snp-sites", mode, "case", k, "exit", ra, rb)
        if ra == 0 and rb == 0 and mode == ["-c"]:
            for flags in (["-c"], ["-c", "-a"], ["-k"], ["-a", "-k", "-b"], ["-m", "-c"]):
                src = a if flags == ["-c"] else mf          # also the raw alignment (N, IUPAC, lower case) for the other modes
                ca, cb = os.path.join(tmp, "a.csv"), os.path.join(tmp, "b.csv")
                ra2 = run(["snp-dists", "-q", *flags, src], ca); rb2 = run([os.path.join(O, "snp-dists"), "-q", *flags, src], cb)
                if (ra2 != 0) != (rb2 != 0) or (ra2 == 0 and open(ca, "rb").read() != open(cb, "rb").read()):
                    bad += 1; print("  DIFFERS snp-dists", flags, "case", k, "exit", ra2, rb2)
print("  %d random cases, %d differences" % (cases, bad))
sys.exit(1 if bad else 0)
PY

if [ $fail -eq 0 ]; then
    { echo "pinned $(date -u +%Y-%m-%dT%H:%M:%SZ)"; echo "snp-sites $(snp-sites -V 2>&1 | head -1)"; echo "snp-dists $(snp-dists -v 2>&1 | head -1)"; } > tests/golden/PINNED
    echo "PINNED: the oracle matches the real tools on every vector (recorded in tests/golden/PINNED)"
else
    echo "NOT PINNED: $fail check(s) differ -- the oracle (oracle/*.c) needs fixing where it differs" >&2
fi
exit $fail
