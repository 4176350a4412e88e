"""One rank of a multi-process distance job (helper of tests/test_gpu_comm.py and tools/, not a test).

  python tests/comm_worker.py <name> <rank> <world> <device> <seqs.npy> <matrix file> <flags> <calls>

Every rank maps the SAME matrix file (a shared-memory file under /dev/shm) and writes its share of
the n x n uint16/uint32 matrix into it through btb_dist_matrix_host_comm.  flags: a = -a, k = -k,
4 = uint32 cells.  Prints "path <0|1>" (1 = sharded upload + GPU-to-GPU exchange)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    name, rank, world, device, seqs_path, out_path, flags, calls = sys.argv[1:9]
    rank, world, device, calls = int(rank), int(world), int(device), int(calls)
    from btb_phylo_b200 import _capi
    lib = _capi.lib()
    seqs = np.load(seqs_path)
    n, L = seqs.shape
    elem = 4 if "4" in flags else 2
    out = np.memmap(out_path, dtype=np.uint32 if elem == 4 else np.uint16, mode="r+", shape=(n, n))
    comm = ctypes.c_void_p()
    _capi.check(lib.btb_comm_create(name.encode(), rank, world, device, ctypes.byref(comm)), "btb_comm_create")
    try:
        for _ in range(calls):
            _capi.check(lib.btb_dist_matrix_host_comm(comm, seqs.ctypes.data, n, L, seqs.strides[0], int("a" in flags),
                                                      int("k" in flags), _capi.ENGINE_AUTO if "p" not in flags else _capi.ENGINE_POPC,
                                                      out.ctypes.data, elem, n), "btb_dist_matrix_host_comm")
        v = ctypes.c_int64(-1)
        lib.btb_query(b"comm_last_path", ctypes.byref(v))
        print("path %d" % v.value, flush=True)
    finally:
        lib.btb_comm_destroy(comm)
    out.flush()


if __name__ == "__main__":
    main()
