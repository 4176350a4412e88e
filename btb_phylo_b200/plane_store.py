"""Incremental snp-sites stage: a PLANE STORE for the consensus files.

btb-phylo is run weekly on a sample set that grows by a few samples (Readme.md:186); the reference
re-reads every consensus file every time -- 217 GB of text at 50,000 samples -- and already keeps a
cache of the downloaded files themselves (btbphylo/phylogeny.py:40-43).  The kept-column set of
snp-sites depends on ALL samples, so nothing downstream of the column reduction can be reused; what
can be reused is k1's output: the three bit-planes of every sample (3 bits per base: 1.63 MB for a
4.35 Mb genome instead of 4.42 MB of text, and no parsing).

    snp_sites(snps.fas, [consensus files ...], plane_store="/path/to/store")

keeps one blob per input file, keyed by the file's absolute path and validated by its size and
modification time (the rule `make` uses).  A later call packs only the files that are new or have
changed (k1, with the column state fused in), loads the planes of all others from the store (their
column state comes from the stand-alone k2 over those planes), and then selects and compacts exactly
as the uncached call does: the result is the same snps.fas, byte for byte -- tests/test_gpu_store.py.

Everything that computes runs in libbtbphylo.so's CUDA kernels through the C-ABI's device-pointer
level (include/btb_phylo.h); PyTorch holds the buffers.  One GPU (the planes of 50,000 samples are
81 GB and fit one B200).  Inputs the predicted framing does not take (gzip, FASTQ, mixed or irregular
layouts) and records that fail the kernel's layout checks are not an error: the call falls back to
btb_snp_sites_files, which parses anything, and leaves the store alone.
"""
import ctypes
import hashlib
import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _capi

MAGIC = b"BTBPLN02"             # 02: column 32 w + cl at bit cl of plane word w
HEADER = struct.Struct("<8sqqqqi")      # magic, G, words, source size, source mtime (ns), name length
CHUNK_BYTES = 64 << 20


def _blob_path(store, path):
    return os.path.join(store, hashlib.sha1(os.path.abspath(path).encode()).hexdigest() + ".planes")


def _read_header(blob, st):
    """(G, words, name) when the blob is valid for a source file with stat result `st`, else None."""
    try:
        with open(blob, "rb") as f:
            raw = f.read(HEADER.size)
            if len(raw) != HEADER.size:
                return None
            magic, G, words, size, mtime, nl = HEADER.unpack(raw)
            if magic != MAGIC or size != st.st_size or mtime != st.st_mtime_ns or not 0 < nl < 4096:
                return None
            name = f.read(nl)
            if len(name) != nl or os.fstat(f.fileno()).st_size != HEADER.size + nl + 3 * words * 4:
                return None
            return G, words, name.decode()
    except OSError:
        return None


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class NotPredictable(Exception):
    """The inputs (or one record) need the exact parser: the caller falls back to the uncached path."""


def snp_sites_cached(lib, out_path, paths, threads, store):
    """-> (n_samples, G, L, n_from_store).  Raises NotPredictable for inputs this path does not take."""
    import torch
    os.makedirs(store, exist_ok=True)
    n = len(paths)
    stats = [os.stat(p) for p in paths]
    heads = [_read_header(_blob_path(store, p), st) for p, st in zip(paths, stats)]
    new_idx = [i for i in range(n) if heads[i] is None]
    old_idx = [i for i in range(n) if heads[i] is not None]

    # ---- frame the new files (one record each, equal layout: what consensus files are) ------------------
    frames, names = None, [None] * n
    if new_idx:
        arr = (ctypes.c_char_p * len(new_idx))(*[os.fsencode(paths[i]) for i in new_idx])
        out = (ctypes.c_int64 * (6 * len(new_idx)))()
        nbuf = ctypes.create_string_buffer(4096 * len(new_idx) + 16)
        nrec = ctypes.c_int64(0)
        _capi.check(lib.btb_fasta_frame_files(arr, len(new_idx), out, len(new_idx), nbuf, len(nbuf), ctypes.byref(nrec)),
                    "btb_fasta_frame_files")
        if nrec.value != len(new_idx):
            raise NotPredictable("the new files are not one regular record each")
        frames = np.frombuffer(out, dtype=np.int64).reshape(-1, 6).copy()
        if not (frames[:, 0] == np.arange(len(new_idx))).all():
            raise NotPredictable("a record spans files")
        for k, nm in enumerate(nbuf.value.decode().split("\n")):
            names[new_idx[k]] = nm
    for i in old_idx:
        names[i] = heads[i][2]
    Gs = set(int(frames[k, 3]) for k in range(len(new_idx))) | set(heads[i][0] for i in old_idx)
    if len(Gs) != 1:
        raise NotPredictable("sequences of unequal length: the exact path words the error")      # (same message as the tool's)
    G = Gs.pop()
    words = int(lib.btb_plane_words(G))
    if any(heads[i][1] != words for i in old_idx):
        raise NotPredictable("store written for another plane layout")

    dev = torch.device("cuda", torch.cuda.current_device())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    nU = len(new_idx)
    planes = torch.empty((3, n, words), dtype=torch.int32, device=dev)        # plane rows: the new samples, then the stored ones
    state = torch.zeros(5 * words, dtype=torch.int32, device=dev)
    flags = torch.zeros(max(1, nU), dtype=torch.int32, device=dev)
    pool = ThreadPoolExecutor(max_workers=max(1, min(int(threads), 64)))
    # two page-locked staging buffers serve both the text of new files and the planes of stored ones
    blob_bytes = 3 * words * 4
    need = CHUNK_BYTES
    if nU:
        need = max(need, int((frames[:, 2].max() + 15) // 16 * 16) + 64)
    if old_idx:
        need = max(need, blob_bytes)
    staging = [torch.empty(need + 64, dtype=torch.uint8, pin_memory=True) for _ in range(2)]

    # ---- new files: text -> pinned chunk -> device -> k1 (planes + column state in one pass) ------------------
    if nU:
        raw = frames[:, 2]
        slot = (raw + 15) // 16 * 16
        per_chunk = max(1, int(need // max(1, int(slot.max()))))
        cap = need + 64
        pinned = staging
        d_text = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(2)]
        busy = [None, None]
        for c, k0 in enumerate(range(0, nU, per_chunk)):
            k1 = min(nU, k0 + per_chunk)
            b = c & 1
            if busy[b] is not None:
                busy[b].synchronize()
            offs = np.concatenate([[0], np.cumsum(slot[k0:k1])])[:-1]
            view = memoryview(pinned[b].numpy())

            def read_one(k, view=view, offs=offs, k0=k0):
                fd = os.open(paths[new_idx[k]], os.O_RDONLY)
                try:
                    o, left, at = int(offs[k - k0]), int(raw[k]), int(frames[k, 1])
                    while left:
                        got = os.preadv(fd, [view[o:o + left]], at)
                        if got <= 0:
                            raise OSError("short read of %s" % paths[new_idx[k]])
                        o += got; at += got; left -= got
                finally:
                    os.close(fd)
            list(pool.map(read_one, range(k0, k1)))
            total = int(offs[-1] + slot[k1 - 1])
            d_text[b][:total + 16].copy_(pinned[b][:total + 16], non_blocking=True)
            d_off = torch.from_numpy(offs.astype(np.int64)).to(dev)
            d_lw = torch.from_numpy(frames[k0:k1, 4].astype(np.int32)).to(dev)
            d_el = torch.from_numpy(frames[k0:k1, 5].astype(np.int32)).to(dev)
            _capi.check(lib.btb_pack_planes_state(_ptr(d_text[b]), total, _ptr(d_off), _ptr(d_lw), _ptr(d_el), k1 - k0, k0, n, G,
                                                  _ptr(planes), _ptr(flags[k0:]), _ptr(state), st), "btb_pack_planes_state")
            busy[b] = torch.cuda.Event()
            busy[b].record()
        if int(flags[:nU].max().item()) != 0:
            raise NotPredictable("a record is not laid out as its first line promised")

    # ---- stored samples: planes from the store, their column state from k2 --------------------------------------
    if old_idx:
        per_chunk = max(1, need // blob_bytes)
        pinned = [s_[:per_chunk * blob_bytes].view(torch.int32).view(3, per_chunk, words) for s_ in staging]
        torch.cuda.synchronize()                                     # (the staging buffers may still feed the last text chunk)
        busy = [None, None]
        for c, k0 in enumerate(range(0, len(old_idx), per_chunk)):
            k1 = min(len(old_idx), k0 + per_chunk)
            b = c & 1
            if busy[b] is not None:
                busy[b].synchronize()
            hostv = pinned[b].numpy()

            def load_one(k, hostv=hostv, k0=k0):
                i = old_idx[k]
                with open(_blob_path(store, paths[i]), "rb") as f:
                    f.seek(HEADER.size + len(heads[i][2].encode()))
                    for p in range(3):
                        if f.readinto(memoryview(hostv[p, k - k0]).cast("B")) != words * 4:
                            raise OSError("short plane blob for %s" % paths[i])
            list(pool.map(load_one, range(k0, k1)))
            for p in range(3):
                planes[p, nU + k0:nU + k1].copy_(pinned[b][p, :k1 - k0], non_blocking=True)
            busy[b] = torch.cuda.Event()
            busy[b].record()
        stored = planes[:, nU:, :]                                   # (k2 walks rows nU .. n-1 of each plane: stride n * words)
        _capi.check(lib.btb_column_mask(_ptr(stored), n - nU, n, G, _ptr(state), 1, st), "btb_column_mask")

    # ---- k3: select, compact ----------------------------------------------------------------------------------------
    keep = torch.empty(words, dtype=torch.int32, device=dev)
    idx = torch.empty(words * 32, dtype=torch.int32, device=dev)
    cnt = torch.zeros(4, dtype=torch.int32, device=dev)
    _capi.check(lib.btb_select_columns(_ptr(state), G, 1, _ptr(keep), _ptr(idx), _ptr(cnt), st), "btb_select_columns")
    L = int(cnt[0].item())
    if L > 0:
        ld = (L + 15) // 16 * 16
        snps = torch.empty((n, ld), dtype=torch.uint8, device=dev)
        _capi.check(lib.btb_compact_columns(_ptr(planes), 0, n, n, G, _ptr(idx), L, _ptr(snps), ld, st), "btb_compact_columns")
        rows = snps[:, :L].cpu().numpy()
        row_of = np.empty(n, dtype=np.int64)                         # plane row of input file i
        row_of[new_idx] = np.arange(nU)
        row_of[old_idx] = nU + np.arange(len(old_idx))
        tmp = "%s.%d.tmp" % (out_path, os.getpid())
        with open(tmp, "wb") as f:
            for i in range(n):
                f.write(b">" + names[i].encode() + b"\n" + rows[row_of[i]].tobytes() + b"\n")
        os.replace(tmp, out_path)

    # ---- the new samples' planes go to the store (also when no SNP was found: the planes are what they are) -------
    if nU:
        per_chunk = max(1, need // blob_bytes)
        pinned = [s_[:per_chunk * blob_bytes].view(torch.int32).view(3, per_chunk, words) for s_ in staging]
        torch.cuda.synchronize()
        chunks = [(k0, min(nU, k0 + per_chunk)) for k0 in range(0, nU, per_chunk)]
        done = [None, None]

        def fetch(c):                                                 # device -> page-locked chunk c & 1, asynchronously
            k0, k1 = chunks[c]
            for p in range(3):
                pinned[c & 1][p, :k1 - k0].copy_(planes[p, k0:k1], non_blocking=True)
            done[c & 1] = torch.cuda.Event()
            done[c & 1].record()
        fetch(0)
        for c, (k0, k1) in enumerate(chunks):
            done[c & 1].synchronize()
            if c + 1 < len(chunks):
                fetch(c + 1)                                          # (its buffer was written out one iteration ago)
            host = pinned[c & 1].numpy()

            def save_one(k, host=host, k0=k0):
                i = new_idx[k]
                blob = _blob_path(store, paths[i])
                nm = names[i].encode()
                part = "%s.%d.tmp" % (blob, os.getpid())
                with open(part, "wb") as f:
                    f.write(HEADER.pack(MAGIC, G, words, stats[i].st_size, stats[i].st_mtime_ns, len(nm)))
                    f.write(nm)
                    for p in range(3):
                        f.write(memoryview(host[p, k - k0]).cast("B"))
                os.replace(part, blob)
            list(pool.map(save_one, range(k0, k1)))
    pool.shutdown()
    return n, G, L, len(old_idx)
