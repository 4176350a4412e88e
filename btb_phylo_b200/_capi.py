"""ctypes binding of libbtbphylo.so (include/btb_phylo.h).  The library is the product; this file
only declares signatures and turns error codes into Python exceptions.  It deliberately has no
fallback: if the shared library is missing or no sm_100 GPU is visible, calls fail loudly."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbtbphylo.so")

ENGINE_AUTO, ENGINE_POPC, ENGINE_UMMA = -1, 0, 1
TILE = 256
SIGMA_DENSE4 = -4
OK, EINVAL, EIO, EFORMAT, ENOSNPS, ECUDA, ENODEVICE, ENOMEM, ECOMM = range(9)

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_vp = ctypes.c_void_p
p_i64 = ctypes.POINTER(ctypes.c_int64)

# every symbol include/btb_phylo.h declares: (restype, argtypes)
SIGNATURES = {
    "btb_last_error": (c_int, [ctypes.c_char_p, ctypes.c_size_t]),
    "btb_version": (ctypes.c_char_p, []),
    "btb_device_count": (c_int, [ctypes.POINTER(c_int)]),
    "btb_set_option": (c_int, [ctypes.c_char_p, c_int]),
    "btb_query": (c_int, [ctypes.c_char_p, p_i64]),
    "btb_snp_sites": (c_int, [ctypes.c_char_p, ctypes.c_char_p, c_int, c_int, c_int, p_i64, p_i64, p_i64]),
    "btb_snp_sites_files": (c_int, [ctypes.POINTER(ctypes.c_char_p), c_i64, ctypes.c_char_p, ctypes.c_char_p, c_int, c_int, c_int,
                                    p_i64, p_i64, p_i64]),
    "btb_fasta_frame_files": (c_int, [ctypes.POINTER(ctypes.c_char_p), c_i64, p_i64, c_i64, ctypes.c_char_p, ctypes.c_size_t, p_i64]),
    "btb_snp_dists": (c_int, [ctypes.c_char_p, ctypes.c_char_p, c_int, c_int, c_int, ctypes.c_char, c_int, c_int,
                              c_int, c_int, c_int, p_i64, p_i64]),
    "btb_dist_matrix_host": (c_int, [c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_vp, c_int, c_i64]),
    "btb_dist_matrix_host_upper": (c_int, [c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int,
                                           c_vp, c_int, c_i64]),
    "btb_release_workspace": (c_int, []),
    "btb_comm_create": (c_int, [ctypes.c_char_p, c_int, c_int, c_int, ctypes.POINTER(c_vp)]),
    "btb_comm_destroy": (c_int, [c_vp]),
    "btb_dist_matrix_host_comm": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_vp, c_int, c_i64]),
    "btb_comm_last_timings": (c_int, [c_vp, ctypes.POINTER(ctypes.c_double), c_int]),
    "btb_extract_host": (c_int, [c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_vp, p_i64, c_vp, c_i64]),
    "btb_plane_words": (c_i64, [c_i64]),
    "btb_pack_planes": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "btb_pack_planes_state": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "btb_column_mask": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_int, c_vp]),
    "btb_select_columns": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    "btb_compact_columns": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "btb_dist_alphabet": (c_int, [c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_vp, ctypes.POINTER(c_int), c_vp]),
    "btb_dist_operand_bytes": (c_i64, [c_int, c_i64, c_i64, c_int]),
    "btb_dist_encode": (c_int, [c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_int, c_vp, c_vp]),
    "btb_dist_encode_rows": (c_int, [c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_int, c_vp, c_i64, c_i64, c_vp]),
    "btb_dist_tile_count": (c_i64, [c_i64]),
    "btb_dist_tile_coords": (c_int, [c_i64, c_i64, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "btb_distance_tiles": (c_int, [c_int, c_vp, c_i64, c_i64, c_int, c_i64, c_i64, c_vp, c_int, c_i64, c_int, c_int, c_vp]),
    "btb_dist_block_rows": (c_i64, [c_i64]),
    "btb_dist_row_share": (c_int, [c_i64, c_int, c_int, p_i64, p_i64]),
    "btb_dist_row_share_weighted": (c_int, [c_i64, c_int, c_int, ctypes.c_double, p_i64, p_i64]),
    "btb_distance_rows": (c_int, [c_int, c_vp, c_i64, c_i64, c_int, c_i64, c_i64, c_vp, c_int, c_i64, c_int, c_vp]),
    "btb_format_csv_rows": (c_int, [c_vp, c_int, c_i64, c_i64, c_i64, c_i64, ctypes.POINTER(ctypes.c_char_p),
                                    ctypes.c_char, c_int, c_vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "btb_process_sample_name": (c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]),
    "btb_fasta_names": (c_int, [ctypes.c_char_p, c_int, ctypes.c_char_p, ctypes.c_size_t, p_i64, ctypes.POINTER(ctypes.c_size_t)]),
    "btb_csv_row_names": (c_int, [ctypes.c_char_p, ctypes.c_char, ctypes.c_char_p, ctypes.c_size_t, p_i64, ctypes.POINTER(ctypes.c_size_t)]),
    "btb_snp_dists_labelled": (c_int, [ctypes.c_char_p, ctypes.c_char_p, c_int, c_int, c_int, ctypes.c_char, c_int, c_int,
                                       c_int, c_int, c_int, ctypes.c_char_p, c_i64, p_i64, p_i64]),
    "btb_relabel_csv": (c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char, ctypes.c_char_p, c_i64]),
}


class BtbError(Exception):
    """A libbtbphylo call failed.  `.code` is the BTB_E* value, the text is btb_last_error()."""

    def __init__(self, code, message, call=""):
        super().__init__("%s failed with code %d: %s" % (call or "libbtbphylo", code, message))
        self.code = code
        self.message = message


_lib = None


def lib():
    """Loads libbtbphylo.so (built in-tree by btb_phylo_b200.build).  No fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m btb_phylo_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
        # BTB_OPTIONS="umma_lockstep=0,umma_dense_simplex=2": tuning knobs of btb_set_option
        for item in filter(None, os.environ.get("BTB_OPTIONS", "").split(",")):
            key, _, val = item.partition("=")
            if L.btb_set_option(key.strip().encode(), int(val)) != OK:
                raise ValueError("BTB_OPTIONS: unknown option %r" % key)
    return _lib


def last_error():
    buf = ctypes.create_string_buffer(2048)
    lib().btb_last_error(buf, len(buf))
    return buf.value.decode(errors="replace")


def check(rc, call=""):
    if rc != OK:
        raise BtbError(rc, last_error(), call)


def device_count():
    n = c_int(0)
    lib().btb_device_count(ctypes.byref(n))
    return n.value
