"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

Python access to the CPU oracle (the C restatement in this directory) plus an independent
numpy restatement used to cross-check the C one.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product package
(btb_phylo_b200) never does.

What is restated (the arithmetic is in third-party programs absent from /root/reference):
  * snp-sites -c      -> btbphylo/phylogeny.py:133   (spec: SURVEY.md App. A.2)
  * snp-dists 0.8.2   -> btbphylo/phylogeny.py:149   (spec: SURVEY.md App. A.3)
PARITY UNPINNED against real binaries; pinned by upstream README vectors (tests/golden/).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
SNP_SITES = os.path.join(BUILD, "snp-sites")
SNP_DISTS = os.path.join(BUILD, "snp-dists")
LIB = os.path.join(BUILD, "liboracle.so")

_lib = None


def build(force=False):
    """Compile the oracle (gcc, seconds).  No-op when the artefacts already exist."""
    srcs = [os.path.join(HERE, f) for f in ("snp_sites_ref.c", "snp_dists_ref.c", "fasta_ref.h", "Makefile")]
    outs = (SNP_SITES, SNP_DISTS, LIB)
    if force or not all(os.path.exists(p) for p in outs) or not _up_to_date(srcs):
        subprocess.run(["make", "-B", "-C", HERE], check=True, capture_output=True)
        with open(os.path.join(BUILD, "sources.sha256"), "w") as f:
            f.write(_digest(srcs))
    return BUILD


def _digest(paths):
    import hashlib
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _up_to_date(srcs):
    """Content, not mtime: the copy of the tree on the GPU box has arbitrary mtimes."""
    try:
        with open(os.path.join(BUILD, "sources.sha256")) as f:
            return f.read().strip() == _digest(srcs)
    except OSError:
        return False


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_snp_dists_rows.restype = ctypes.c_int
        _lib.oracle_snp_dists_rows.argtypes = [
            ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
        _lib.oracle_snp_dists_prepare.restype = ctypes.c_void_p
        _lib.oracle_snp_dists_prepare.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_int]
        _lib.oracle_snp_dists_release.restype = None
        _lib.oracle_snp_dists_release.argtypes = [ctypes.c_void_p]
        _lib.oracle_snp_dists_rows_prepared.restype = ctypes.c_int
        _lib.oracle_snp_dists_rows_prepared.argtypes = [
            ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
        _lib.oracle_snp_sites_mask.restype = ctypes.c_long
        _lib.oracle_snp_sites_mask.argtypes = [
            ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_snp_sites_file.restype = ctypes.c_int
        _lib.oracle_snp_sites_file.argtypes = [
            ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int,
            ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long)]
    return _lib


# ---------------------------------------------------------------- C oracle, in-memory forms
def snp_dists_rows(seqs, allchars=False, keepcase=False, threads=1, row0=0, row1=None):
    """seqs: uint8 [n, L].  Returns uint32 [(row1-row0), n] exactly as snp-dists would print."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    n, L = seqs.shape
    row1 = n if row1 is None else row1
    out = np.empty((row1 - row0, n), dtype=np.uint32)
    rc = lib().oracle_snp_dists_rows(seqs.ctypes.data, n, L, int(allchars), int(keepcase),
                                     int(threads), row0, row1, out.ctypes.data)
    if rc:
        raise RuntimeError("oracle_snp_dists_rows failed: %d" % rc)
    return out


class PreparedAlignment:
    """snp-dists' load step done once (upper-casing, non-ACGT -> '.'); rows() is its loop nest.
    Splitting the two keeps the load out of a timed region, as it is in the real tool."""

    def __init__(self, seqs, allchars=False, keepcase=False):
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        self.n, self.L = seqs.shape
        self._p = lib().oracle_snp_dists_prepare(seqs.ctypes.data, self.n, self.L, int(allchars), int(keepcase))
        if not self._p:
            raise MemoryError("oracle_snp_dists_prepare")

    def rows(self, row0, row1, threads=1, out=None):
        if out is None:
            out = np.empty((row1 - row0, self.n), dtype=np.uint32)
        rc = lib().oracle_snp_dists_rows_prepared(self._p, self.n, self.L, int(threads), row0, row1, out.ctypes.data)
        if rc:
            raise RuntimeError("oracle_snp_dists_rows_prepared failed: %d" % rc)
        return out

    def close(self):
        if self._p:
            lib().oracle_snp_dists_release(self._p)
            self._p = None

    def __del__(self):
        self.close()


def snp_sites_mask(seqs, pure=True):
    """seqs: uint8 [n, G].  Returns bool [G], True where snp-sites keeps the column."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    n, G = seqs.shape
    keep = np.zeros(G, dtype=np.uint8)
    lib().oracle_snp_sites_mask(seqs.ctypes.data, n, G, int(pure), keep.ctypes.data)
    return keep.astype(bool)


# ---------------------------------------------------------------- C oracle, file forms (the CLIs)
def run_snp_sites(in_fasta, out_fasta, pure=True):
    build()
    cmd = [SNP_SITES] + (["-c"] if pure else []) + ["-o", out_fasta, in_fasta]
    return subprocess.run(cmd, capture_output=True)


def run_snp_dists(in_fasta, out_path, threads=1, flags=("-c",)):
    build()
    with open(out_path, "wb") as f:
        return subprocess.run([SNP_DISTS, *flags, "-j", str(threads), in_fasta], stdout=f,
                              stderr=subprocess.PIPE)


# ---------------------------------------------------------------- independent numpy restatement
def np_prepare(seqs, allchars=False, keepcase=False):
    s = np.array(seqs, dtype=np.uint8, copy=True)
    if not keepcase:
        lower = (s >= ord("a")) & (s <= ord("z"))
        s[lower] -= 32
    if not allchars:
        ok = (s == ord("A")) | (s == ord("C")) | (s == ord("G")) | (s == ord("T"))
        s[~ok] = ord(".")
    return s


def np_snp_dists(seqs, allchars=False, keepcase=False):
    """Full n x n matrix by direct evaluation of the definition (small n only)."""
    s = np_prepare(seqs, allchars, keepcase)
    n = s.shape[0]
    dot = s != ord(".")
    out = np.zeros((n, n), dtype=np.uint32)
    for j in range(n):
        out[j] = ((s[j][None, :] != s) & dot[j][None, :] & dot).sum(axis=1)
    return out


def np_snp_sites_mask(seqs, pure=True):
    s = np.asarray(seqs, dtype=np.uint8)
    up = s.copy()
    lower = (up >= ord("a")) & (up <= ord("z"))
    up[lower] -= 32
    if pure:
        ok = (up == ord("A")) | (up == ord("C")) | (up == ord("G")) | (up == ord("T"))
        allok = ok.all(axis=0)
        nd = np.zeros(s.shape[1], dtype=np.int64)
        for b in b"ACGT":
            nd += (up == b).any(axis=0)
        return allok & (nd >= 2)
    unknown = (s == ord("N")) | (s == ord("n")) | (s == ord("-")) | (s == ord("?"))
    keep = np.zeros(s.shape[1], dtype=bool)
    for k in range(s.shape[1]):
        vals = set(up[~unknown[:, k], k].tolist())
        keep[k] = len(vals) >= 2
    return keep


def format_matrix(names, mat, sep=",", corner=True):
    """The bytes snp-dists prints for a full matrix (App. A.3 output rules)."""
    head = ("snp-dists 0.8.2" if corner else "") + "".join(sep + n for n in names) + "\n"
    rows = [names[j] + "".join(sep + str(int(v)) for v in mat[j]) + "\n" for j in range(len(names))]
    return (head + "".join(rows)).encode()
