"""Drop-in for the two hot functions of btbphylo/phylogeny.py.

The reference shells out to two CPU programs:
    snp_sites()         btbphylo/phylogeny.py:128-141  ->  `snp-sites <in> -c -o <out>`      (:133)
    build_snp_matrix()  btbphylo/phylogeny.py:144-150  ->  `snp-dists -c -j T <in> > <out>`  (:149)
Here the same names, argument order and files are kept, but the work is done by libbtbphylo.so
(hand-written CUDA for sm_100a) through its C-ABI (include/btb_phylo.h).  A caller switches by
changing one import:   from btb_phylo_b200 import phylogeny

Error behaviour follows btbphylo/utils.py:155-176: any failure surfaces as a bare `Exception`
whose text carries the equivalent command line and exit code 1 (snp-sites and snp-dists both
exit 1 on bad input / no SNPs).  There is no CPU fallback: without the CUDA library or a B200 the
calls raise.
"""
import ctypes
import os
import re

from . import _capi

__all__ = ["snp_sites", "build_snp_matrix", "build_multi_fasta", "post_process_snps_csv", "post_process_snps_df",
           "process_sample_name", "extract_submission_no", "PhylogenyError"]


class PhylogenyError(Exception):
    """Raised like the reference's bare Exception (utils.py:170-174); kept a subclass so that
    `except Exception` in callers (btb_phylo.py) behaves identically."""


def _gpus(input_bytes=None):
    """GPU count: BTBPHYLO_GPUS, else the visible sm_100 devices -- but no more than one per 4 GB of
    input text when the size is known (a second CUDA context costs more than it saves on small files)."""
    env = os.environ.get("BTBPHYLO_GPUS")
    if env:
        return max(1, int(env))
    n = max(1, _capi.device_count())
    if input_bytes is not None:
        n = max(1, min(n, input_bytes // (4 << 30)))
    return n


def _engine():
    return {"popc": _capi.ENGINE_POPC, "umma": _capi.ENGINE_UMMA}.get(
        os.environ.get("BTBPHYLO_ENGINE", "auto").lower(), _capi.ENGINE_AUTO)


def _raise_like_run(cmd, err):
    # same shape as the message utils.run builds (btbphylo/utils.py:170-174)
    raise PhylogenyError("""*****
            %s
            cmd failed with exit code %i
            (%s)
          *****""" % (cmd, 1, err))


def snp_sites(snp_sites_outpath, multi_fasta_path, threads=None, plane_store=None):
    """Run the snp-sites stage on consensus files (reference: phylogeny.py:128-141).

    Writes `snp_sites_outpath` (one unwrapped record per sample) and returns
    {"number_of_snps": L}.  `threads` (host threads that read the file into the upload buffers) is
    optional -- the reference signature has two positional arguments and those keep working; the
    default is every core (BTBPHYLO_THREADS overrides), since the stage is bound by reading the file.

    `multi_fasta_path` may also be a list / tuple of per-sample consensus FASTA paths: they are
    taken as if concatenated in that order -- the file build_multi_fasta (phylogeny.py:48-88) would
    have written -- without writing or re-reading it.

    `plane_store` (a directory; only with a list of per-sample files) makes the call incremental
    (the weekly run): the bit-planes of every file are kept there, keyed by path,
    size and modification time, and a later call reads the text of new or changed files only --
    see plane_store.py.  Same snps.fas, byte for byte.
    """
    n = ctypes.c_int64(0)
    g = ctypes.c_int64(0)
    l = ctypes.c_int64(0)
    if threads is None:
        threads = int(os.environ.get("BTBPHYLO_THREADS", 0)) or min(32, os.cpu_count() or 1)
    def size_of(path):
        try:
            return os.path.getsize(path)
        except OSError:
            return 0                                       # the library reports the unreadable file
    if plane_store and isinstance(multi_fasta_path, (list, tuple)) and len(multi_fasta_path) > 0:
        from . import plane_store as _store
        try:
            _, _, n_snps, _ = _store.snp_sites_cached(_capi.lib(), snp_sites_outpath, [os.fspath(x) for x in multi_fasta_path],
                                                      threads, os.fspath(plane_store))
        except _store.NotPredictable:
            n_snps = None                                  # inputs that need the exact parser: the plain path below
        if n_snps is not None:
            if n_snps == 0:
                _raise_like_run("snp-sites <%d consensus files> -c -o %s" % (len(multi_fasta_path), snp_sites_outpath),
                                "Warning: No SNPs were detected so there is nothing to output.")
            return {"number_of_snps": n_snps}
    if isinstance(multi_fasta_path, (list, tuple)):
        paths = [os.fsencode(x) for x in multi_fasta_path]
        shown = "<%d consensus files>" % len(paths)
        arr = (ctypes.c_char_p * max(1, len(paths)))(*paths)
        rc = _capi.lib().btb_snp_sites_files(arr, len(paths), shown.encode(), os.fsencode(snp_sites_outpath), 1,
                                             _gpus(sum(size_of(x) for x in paths)), int(threads),
                                             ctypes.byref(n), ctypes.byref(g), ctypes.byref(l))
    else:
        shown = multi_fasta_path
        rc = _capi.lib().btb_snp_sites(os.fsencode(multi_fasta_path), os.fsencode(snp_sites_outpath), 1,
                                       _gpus(size_of(multi_fasta_path)), int(threads),
                                       ctypes.byref(n), ctypes.byref(g), ctypes.byref(l))
    if rc != _capi.OK:
        _raise_like_run("snp-sites %s -c -o %s" % (shown, snp_sites_outpath), _capi.last_error())
    # read the number of snps for metadata exactly as the reference does (phylogeny.py:136-140)
    with open(snp_sites_outpath, "r") as f:
        next(f)
        for line in f:
            metadata = {"number_of_snps": len(line) - 1}
            break
    return metadata


def build_multi_fasta(multi_fasta_path, consensus_paths):
    """The local half of the reference's build_multi_fasta (phylogeny.py:48-88): append the consensus
    files, in order, to `multi_fasta_path` (the S3 download half, phylogeny.py:26-45, is the control
    plane and out of scope).  Only needed when the multi-FASTA itself is wanted (light_mode=False,
    btb_phylo.py:402-403); snp_sites() takes the list of consensus files directly."""
    with open(multi_fasta_path, "wb") as out:
        for path in consensus_paths:
            with open(path, "rb") as f:
                while True:
                    block = f.read(1 << 24)
                    if not block:
                        break
                    out.write(block)


def build_snp_matrix(snp_dists_outpath, snp_sites_outpath, threads=1, post_process=False):
    """Run the snp-dists stage (reference: phylogeny.py:144-150).  `threads` may be int or str
    (btb_phylo.py:683 passes the CLI string through).

    post_process=True (not in the reference signature) makes the CSV writer print the ViewBovine
    labels directly, i.e. the file is what build_snp_matrix() followed by post_process_snps_csv()
    (phylogeny.py:172-187, called at btb_phylo.py:598) leaves behind -- without the pandas parse and
    re-serialisation of the O(N^2) matrix."""
    t = max(1, int(threads))
    # one GPU computes the 50,000-sample matrix in 30 ms: the stage is bound by reading snps.fas and writing
    # snps.csv, and every further GPU costs a CUDA context and an upload.  More GPUs only when asked for
    # (BTBPHYLO_GPUS) or when the alignment is large enough (> 8 GB) for its operands to crowd one device.
    try:
        fas_bytes = os.path.getsize(snp_sites_outpath)
    except OSError:
        fas_bytes = 0
    gpus = _gpus(fas_bytes // 2)
    if post_process:
        labels = _labels_for(_names_from(_capi.lib().btb_fasta_names, os.fsencode(snp_sites_outpath), t))
        blob = "\n".join(labels).encode("utf-8")
        rc = _capi.lib().btb_snp_dists_labelled(os.fsencode(snp_sites_outpath), os.fsencode(snp_dists_outpath),
                                                0, 0, 0, b",", 1, 0, gpus, t, _engine(), blob, len(labels), None, None)
    else:
        rc = _capi.lib().btb_snp_dists(os.fsencode(snp_sites_outpath), os.fsencode(snp_dists_outpath),
                                       0, 0, 0, b",", 1, 0, gpus, t, _engine(), None, None)
    if rc != _capi.OK:
        _raise_like_run("snp-dists -c -j %s %s > %s" % (threads, snp_sites_outpath, snp_dists_outpath),
                        _capi.last_error())


# ---------------------------------------------------------------- labels for ViewBovine
_SUBMISSION = re.compile(r"\d{2,2}-\d{4,5}-\d{2,2}")   # the pattern btbphylo/utils.py:116 searches for


def extract_submission_no(sample_name):
    """Submission number of a sample name (reference: btbphylo/utils.py:110-119): the first
    NN-NNNN[N]-NN run prefixed with 'AF-', or the name itself; upper-cased."""
    found = _SUBMISSION.search(sample_name)
    return ("AF-" + found.group(0) if found else sample_name).upper()


def process_sample_name(sample_name):
    """Label used by the cattle / movement datasets (reference: phylogeny.py:206-216): the
    '_consensus' suffix is dropped, then extract_submission_no() applies."""
    if sample_name.endswith("_consensus"):
        sample_name = sample_name[: -len("_consensus")]
    return extract_submission_no(sample_name)


def post_process_snps_df(snp_matrix):
    """Reference: phylogeny.py:190-203 -- a copy of the dataframe whose index and columns are the
    processed index labels."""
    out = snp_matrix.copy(deep=True)
    labels = snp_matrix.index.map(process_sample_name)
    out.index = labels
    out.columns = labels
    return out


# what pandas.read_csv turns into NaN by default; such a row name makes the reference fail
_PANDAS_NA = {"", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN", "<NA>",
              "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"}
_NUMERIC = re.compile(r"[+-]?(\d+\.?\d*(e[+-]?\d+)?|\.\d+(e[+-]?\d+)?|inf|infinity)", re.IGNORECASE)
_BOOLS = {"True", "TRUE", "true", "False", "FALSE", "false"}


def _names_from(call, *args):
    n, need = ctypes.c_int64(0), ctypes.c_size_t(0)
    _capi.check(call(*args, None, 0, ctypes.byref(n), ctypes.byref(need)), "names")
    buf = ctypes.create_string_buffer(max(1, need.value))
    _capi.check(call(*args, buf, need.value, ctypes.byref(n), ctypes.byref(need)), "names")
    return buf.value.decode("utf-8").split("\n") if n.value else []


def _labels_for(names):
    """process_sample_name over the row names, refusing the inputs on which the reference's pandas
    round trip (read_csv(index_col=0) -> rename -> to_csv) does not give a plain relabelled file:
    names pandas would read as missing values, an all-numeric / all-boolean index (not strings any
    more: the reference dies in .endswith), and names that need CSV quoting."""
    if any(x in _PANDAS_NA for x in names):
        raise PhylogenyError("sample name read as a missing value by pandas: cannot be relabelled")
    if names and (all(_NUMERIC.fullmatch(x) for x in names) or all(x in _BOOLS for x in names)):
        raise PhylogenyError("sample names are all numeric / boolean: pandas would not keep them as strings")
    if any(x.startswith('"') or "," in x or "\r" in x for x in names):
        raise PhylogenyError("sample name needs CSV quoting: snps.csv would not parse")
    # a quote inside a name survives read_csv as a literal and makes to_csv quote the field
    return ['"%s"' % l.replace('"', '""') if '"' in l else l for l in map(process_sample_name, names)]


def post_process_snps_csv(snp_dists_outpath):
    """Reference: phylogeny.py:172-187 -- rewrites snps.csv in place with the row and column labels
    processed.  The reference parses all N^2 numbers with pandas and prints them again; here one
    streaming pass swaps the names and copies every distance byte through (btb_relabel_csv)."""
    path = os.fsencode(snp_dists_outpath)
    labels = _labels_for(_names_from(_capi.lib().btb_csv_row_names, path, b","))
    rc = _capi.lib().btb_relabel_csv(path, path, b",", "\n".join(labels).encode("utf-8"), len(labels))
    if rc != _capi.OK:
        raise PhylogenyError("post_process_snps_csv(%s): %s" % (snp_dists_outpath, _capi.last_error()))
