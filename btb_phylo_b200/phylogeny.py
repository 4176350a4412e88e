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

from . import _capi

__all__ = ["snp_sites", "build_snp_matrix", "PhylogenyError"]


class PhylogenyError(Exception):
    """Raised like the reference's bare Exception (utils.py:170-174); kept a subclass so that
    `except Exception` in callers (btb_phylo.py) behaves identically."""


def _gpus():
    """GPU count: BTBPHYLO_GPUS, else every visible sm_100 device."""
    env = os.environ.get("BTBPHYLO_GPUS")
    if env:
        return max(1, int(env))
    return max(1, _capi.device_count())


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


def snp_sites(snp_sites_outpath, multi_fasta_path, threads=1):
    """Run the snp-sites stage on consensus files (reference: phylogeny.py:128-141).

    Writes `snp_sites_outpath` (one unwrapped record per sample) and returns
    {"number_of_snps": L}.  `threads` (host staging threads) is optional; the reference
    signature has two positional arguments and those keep working.
    """
    n = ctypes.c_int64(0)
    g = ctypes.c_int64(0)
    l = ctypes.c_int64(0)
    rc = _capi.lib().btb_snp_sites(os.fsencode(multi_fasta_path), os.fsencode(snp_sites_outpath), 1,
                                   _gpus(), int(threads), ctypes.byref(n), ctypes.byref(g), ctypes.byref(l))
    if rc != _capi.OK:
        _raise_like_run("snp-sites %s -c -o %s" % (multi_fasta_path, snp_sites_outpath), _capi.last_error())
    # read the number of snps for metadata exactly as the reference does (phylogeny.py:136-140)
    with open(snp_sites_outpath, "r") as f:
        next(f)
        for line in f:
            metadata = {"number_of_snps": len(line) - 1}
            break
    return metadata


def build_snp_matrix(snp_dists_outpath, snp_sites_outpath, threads=1):
    """Run the snp-dists stage (reference: phylogeny.py:144-150).  `threads` may be int or str
    (btb_phylo.py:683 passes the CLI string through)."""
    t = max(1, int(threads))
    rc = _capi.lib().btb_snp_dists(os.fsencode(snp_sites_outpath), os.fsencode(snp_dists_outpath),
                                   0, 0, 0, b",", 1, 0, _gpus(), t, _engine(), None, None)
    if rc != _capi.OK:
        _raise_like_run("snp-dists -c -j %s %s > %s" % (threads, snp_sites_outpath, snp_dists_outpath),
                        _capi.last_error())
