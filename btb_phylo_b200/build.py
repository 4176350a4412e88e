"""Builds libbtbphylo.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the library has a plain C ABI and is loaded with ctypes).  Every source is compiled to its own
object (in parallel, rebuilt only when it or a header changed) and the objects are linked."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbtbphylo.so")
STAMP = LIB + ".stamp"
SOURCES = ["capi.cu", "host_io.cpp", "dist_encode.cu", "dist_popc.cu", "dist_umma.cu", "extract.cu", "phylo_files.cu",
           "dist_comm.cu", "csv_writer.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC,-O3,-pthread", "-ccbin", "/usr/bin/g++", "-Xptxas", "-v"]
# BTB_EXPERIMENTS=1 compiles the measured-and-dropped kernel variants back in (tools/ only)
if os.environ.get("BTB_EXPERIMENTS"):
    FLAGS += ["-DBTB_EXPERIMENTS"]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return hs + [os.path.join(HERE, "..", "include", "btb_phylo.h"), os.path.abspath(__file__)]


def _obj(src):
    tag = "_exp" if os.environ.get("BTB_EXPERIMENTS") else ""
    return os.path.join(OBJ, os.path.basename(src) + tag + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _fingerprint():
    """Content hash of everything the library is built from: the GPU box receives a copy of the tree
    whose mtimes mean nothing, so staleness is decided by content."""
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for f in sorted(sources() + _headers()):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _fingerprint()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    log = []

    def compile_one(src):
        obj = _obj(src)
        if not force and not _stale(obj, [src] + hdrs):
            return 0, ""
        cmd = [NVCC, *FLAGS, "-c", "-o", obj, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return r.returncode, " ".join(cmd) + "\n" + r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    failed = False
    for rc, text in results:
        log.append(text)
        if rc:
            failed = True
    if not failed:
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "--shared",
               "-Xcompiler", "-fPIC,-pthread", "-o", LIB, *[_obj(s) for s in sources()], "-lz", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        failed = r.returncode != 0
    if verbose or failed:
        sys.stderr.write("".join(log))
    if failed:
        raise RuntimeError("nvcc failed building libbtbphylo.so")
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write("".join(log))
    with open(STAMP, "w") as f:
        f.write(_fingerprint() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv))
