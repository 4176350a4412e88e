"""Incremental snp-sites stage: a PLANE STORE for the consensus files.

btb-phylo is run weekly on a sample set that grows by a few samples (Readme.md:186); the reference
re-reads every consensus file every time -- 217 GB of text at 50,000 samples -- and already keeps a
cache of the downloaded files themselves (btbphylo/phylogeny.py:40-43).  The kept-column set of
snp-sites depends on ALL samples, so nothing downstream of the column reduction can be reused; what
can be reused is k1's output: the three bit-planes of every sample (3 bits per base: 1.63 MB for a
4.35 Mb genome instead of 4.42 MB of text, and no parsing).

    snp_sites(snps.fas, [consensus files ...], plane_store="/path/to/store")

keeps one blob per input file, keyed by the file's absolute path and validated by its size and
modification time (the rule `make` uses), plus a small index of what the blob headers say (a hint that
saves one open per blob; every header is still read and compared when its planes are).  A later call packs only the files that are new or have
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
import json
import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _capi

MAGIC = b"BTBPLN02"             # 02: column 32 w + cl at bit cl of plane word w
HEADER = struct.Struct("<8sqqqqi")      # magic, G, words, source size, source mtime (ns), name length
CHUNK_BYTES = 64 << 20


INDEX = "index.json"           # blob file name -> [G, words, source size, source mtime (ns), record name]: what the blob headers say


def _blob_path(store, path):
    return os.path.join(store, hashlib.sha1(os.fsencode(os.path.abspath(path))).hexdigest() + ".planes")


def _load_index(store):
    """The store's index: one small file instead of one open + read per blob (0.75 s for 16,000 blobs).  It is only a
    hint -- every blob's header is read together with its planes and must say the same -- and it heals itself: blobs
    it does not list are looked at directly, entries whose blob turns out stale are dropped."""
    try:
        with open(os.path.join(store, INDEX)) as f:
            d = json.load(f)
        if d.get("magic") == MAGIC.decode() and isinstance(d.get("blobs"), dict):
            return d["blobs"]
    except (OSError, ValueError):
        pass
    return {}


def _save_index(store, blobs):
    part = os.path.join(store, "%s.%d.tmp" % (INDEX, os.getpid()))
    with open(part, "w") as f:
        json.dump({"magic": MAGIC.decode(), "blobs": blobs}, f)
    os.replace(part, os.path.join(store, INDEX))


class _StaleBlob(Exception):
    """A blob the index vouched for is not what its header should say: drop it and start over."""


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
            return G, words, name.decode("utf-8", "surrogateescape")
    except OSError:
        return None


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class NotPredictable(Exception):
    """The inputs (or one record) need the exact parser: the caller falls back to the uncached path."""


def snp_sites_cached(lib, out_path, paths, threads, store):
    """-> (n_samples, G, L, n_from_store).  Raises NotPredictable for inputs this path does not take."""
    with ThreadPoolExecutor(max_workers=max(1, min(int(threads), 64))) as pool:
        for _ in range(4):
            try:
                return _snp_sites_cached(lib, out_path, paths, store, pool)
            except _StaleBlob as stale:
                blobs = _load_index(store)
                for b in stale.args[0]:
                    blobs.pop(os.path.basename(b), None)
                    try:
                        os.unlink(b)
                    except OSError:
                        pass
                _save_index(store, blobs)
    raise NotPredictable("the plane store keeps changing under this call")


def _snp_sites_cached(lib, out_path, paths, store, pool):
    import sys
    import time
    import torch
    marks = [("start", time.perf_counter())]                         # BTB_STORE_VERBOSE=1: phase timers on stderr
    os.makedirs(store, exist_ok=True)
    n = len(paths)
    stats = [os.stat(p) for p in paths]
    index = _load_index(store)
    blob_of = [_blob_path(store, p) for p in paths]
    index_dirty = False

    def head_of(k):
        e = index.get(os.path.basename(blob_of[k]))
        if e is not None and e[2] == stats[k].st_size and e[3] == stats[k].st_mtime_ns:
            return e[0], e[1], e[4]
        return _read_header(blob_of[k], stats[k])                     # not listed (or listed for an older source): look
    heads = [head_of(k) for k in range(n)]
    for k in range(n):
        key = os.path.basename(blob_of[k])
        if heads[k] is not None and key not in index:
            index[key] = [heads[k][0], heads[k][1], stats[k].st_size, stats[k].st_mtime_ns, heads[k][2]]
            index_dirty = True
    new_idx = [i for i in range(n) if heads[i] is None]
    old_idx = [i for i in range(n) if heads[i] is not None]

    marks.append(("store headers", time.perf_counter()))
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
        for k, nm in enumerate(nbuf.value.decode("utf-8", "surrogateescape").split("\n")):
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

    marks.append(("frame new files", time.perf_counter()))
    dev = torch.device("cuda", torch.cuda.current_device())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    nU = len(new_idx)
    planes = torch.empty((3, n, words), dtype=torch.int32, device=dev)        # plane rows: the new samples, then the stored ones
    state = torch.zeros(5 * words, dtype=torch.int32, device=dev)
    flags = torch.zeros(max(1, nU), dtype=torch.int32, device=dev)
    # two page-locked staging buffers serve both the text of new files and the planes of stored ones; larger ones
    # for large sample sets (fewer, better filled rounds of the reader threads)
    blob_bytes = 3 * words * 4
    need = min(max(CHUNK_BYTES, n * blob_bytes // 64), 4 * CHUNK_BYTES)
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

    marks.append(("pack new files", time.perf_counter()))
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
                nm = heads[i][2].encode("utf-8", "surrogateescape")
                want = HEADER.pack(MAGIC, G, words, stats[i].st_size, stats[i].st_mtime_ns, len(nm)) + nm
                head = bytearray(len(want))
                try:
                    fd = os.open(blob_of[i], os.O_RDONLY)
                except OSError:
                    return blob_of[i]
                try:                                                  # one scatter read: header, then the planes to their rows
                    bufs = [memoryview(head)] + [memoryview(hostv[p, k - k0]).cast("B") for p in range(3)]
                    total, got = len(want) + blob_bytes, 0
                    while got < total:
                        rest, skip = [], got
                        for b in bufs:
                            if skip >= len(b):
                                skip -= len(b)
                            else:
                                rest.append(b[skip:])
                                skip = 0
                        more = os.preadv(fd, rest, got)
                        if more <= 0:
                            return blob_of[i]
                        got += more
                    if bytes(head) != want or os.fstat(fd).st_size != total:
                        return blob_of[i]
                finally:
                    os.close(fd)
                return None
            stale = [b for b in pool.map(load_one, range(k0, k1)) if b is not None]
            if stale:
                torch.cuda.synchronize()
                raise _StaleBlob(stale)
            for p in range(3):
                planes[p, nU + k0:nU + k1].copy_(pinned[b][p, :k1 - k0], non_blocking=True)
            busy[b] = torch.cuda.Event()
            busy[b].record()
        stored = planes[:, nU:, :]                                   # (k2 walks rows nU .. n-1 of each plane: stride n * words)
        _capi.check(lib.btb_column_mask(_ptr(stored), n - nU, n, G, _ptr(state), 1, st), "btb_column_mask")

    marks.append(("load stored planes", time.perf_counter()))
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
            for i0 in range(0, n, 512):
                f.write(b"".join(b">" + names[i].encode("utf-8", "surrogateescape") + b"\n" + rows[row_of[i]].tobytes() + b"\n"
                                 for i in range(i0, min(n, i0 + 512))))
        os.replace(tmp, out_path)

    marks.append(("select, compact, snps.fas", time.perf_counter()))
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
                blob = blob_of[i]
                nm = names[i].encode("utf-8", "surrogateescape")
                part = "%s.%d.%d.tmp" % (blob, os.getpid(), k)        # (the same file may be listed twice)
                with open(part, "wb") as f:
                    f.write(HEADER.pack(MAGIC, G, words, stats[i].st_size, stats[i].st_mtime_ns, len(nm)))
                    f.write(nm)
                    for p in range(3):
                        f.write(memoryview(host[p, k - k0]).cast("B"))
                os.replace(part, blob)
            list(pool.map(save_one, range(k0, k1)))
        for i in new_idx:
            index[os.path.basename(blob_of[i])] = [G, words, stats[i].st_size, stats[i].st_mtime_ns, names[i]]
        index_dirty = True
    if index_dirty:
        _save_index(store, index)
    marks.append(("save new planes", time.perf_counter()))
    if os.environ.get("BTB_STORE_VERBOSE"):
        print("plane store: %d samples, %d from the store; " % (n, len(old_idx)) +
              ", ".join("%s %.3f s" % (marks[k][0], marks[k][1] - marks[k - 1][1]) for k in range(1, len(marks))), file=sys.stderr)
    return n, G, L, len(old_idx)
