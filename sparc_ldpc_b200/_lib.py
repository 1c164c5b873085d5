"""ctypes binding of libsparc_b200.so (C ABI declared in include/sparc_b200.h).

The library is the product: there is NO CPU fallback.  If the shared object is missing
this module raises at import of any symbol with build instructions; if no CUDA device is
usable the entry points return SB_ECUDA and `check()` raises with the library's message.
"""
import ctypes as ct
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPARC_B200_LIB: another build of the same library (kernel A/B experiments, tools/ab_build.sh)
LIB_PATH = os.environ.get("SPARC_B200_LIB") or os.path.join(_HERE, "libsparc_b200.so")

SB_BP_SUMPROD2, SB_BP_SUMPROD, SB_BP_MINSUM, SB_BP_SUMPROD2_FAST = 0, 1, 2, 3
SB_AMP_STOPPED, SB_AMP_REF_NAN = 1, 2
SB_AMP_STRICT, SB_AMP_FAST, SB_AMP_F64 = 0, 1, 2
SB_MAX_ITCOUNT = 200

_lib = None

_vp, _i, _l, _d = ct.c_void_p, ct.c_int, ct.c_long, ct.c_double

# name -> (restype, argtypes); mirrors include/sparc_b200.h one to one
SIGNATURES = {
    "sb_last_error": (ct.c_char_p, []),
    "sb_version": (_i, []),
    "sb_launch_count": (_l, []),
    "sb_launch_count_reset": (None, []),
    "sumprod": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "sumprod2": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "minsum": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _d]),
    "Lxor": (_d, [_d, _d, _i]),
    "Lxfb": (_d, [_vp, _l, _i]),
    "sb_graph_create": (_i, [_vp, _vp, _vp, _i, _i, _i, ct.POINTER(_vp)]),
    "sb_graph_destroy": (None, [_vp]),
    "sb_bp_batch": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _d, _vp]),
    "sb_bp_lxor_peak": (_i, [_i, ct.POINTER(_d)]),
    "sb_operator_create": (_i, [_vp, _i, _i, _i, ct.POINTER(_vp)]),
    "sb_operator_destroy": (None, [_vp]),
    "sb_fast_tables_check": (_i, [_vp, _i, _i, _i, _vp]),
    "sb_pair_tables_check": (_i, [_vp, _i, _i, _i, _vp]),
    "sb_fht_inplace_host": (_i, [_vp, _l]),
    "sb_Ab_batch": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "sb_Az_batch": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "sb_onehot_apply_batch": (_i, [_vp, _vp, _vp, _vp, _d, _i, _vp, _vp]),
    "sb_amp_pair_enable": (_i, [_i]),
    "sb_pair_tables_check_f64": (_i, [_vp, _i, _i, _i, _vp]),
    "sb_amp_batch": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sb_sp2bp_llr_batch": (_i, [_vp, _l, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _l, _vp]),
    "sb_bp2sp_prior_batch": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "sb_section_softmax_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sb_dense_create": (_i, [_vp, _i, _i, ct.POINTER(_vp)]),
    "sb_dense_destroy": (None, [_vp]),
    "sb_dense_apply_batch": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "sb_dense_amp_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sb_dense_amp_batch_sharded": (_i, [_vp, _vp, _vp, _d, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sb_enable_peer_access": (_i, [_i]),
    "sb_p2p_alloc": (_i, [_l, ct.POINTER(_vp), _vp]),
    "sb_p2p_open": (_i, [_vp, ct.POINTER(_vp)]),
    "sb_p2p_close": (_i, [_vp]),
    "sb_p2p_free": (_i, [_vp]),
    "sb_dense_amp_batch_p2p": (_i, [_vp, _vp, _vp, _d, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sb_argmax_batch": (_i, [_vp, _l, _i, _i, _i, _vp, _l, _vp]),
    "sb_llr2idx_batch": (_i, [_vp, _l, _i, _i, _i, _vp, _l, _vp]),
    "sb_count_errors_batch": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "sb_onehot_beta_batch": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "sb_ldpc_encode_batch": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "sb_bits2idx_batch": (_i, [_vp, _l, _i, _i, _i, _vp, _l, _vp]),
    "sb_threshold_peel_batch": (_i, [_vp, _i, _i, _i, _d, _i, _vp, _vp, _vp, _vp]),
    "sb_exit_hist_batch": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp]),
}


class SparcB200Error(RuntimeError):
    pass


def lib():
    """Load libsparc_b200.so (once) and declare every prototype."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SparcB200Error(
                "%s not found: build it with `make -C sparc_ldpc_b200/csrc` (or __graft_entry__.build()); "
                "there is no CPU fallback" % LIB_PATH)
        L = ct.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc < 0:
        msg = lib().sb_last_error().decode(errors="replace")
        raise SparcB200Error("%s failed (%d): %s" % (what or "libsparc_b200 call", rc, msg))
    return rc


def launch_count():
    return lib().sb_launch_count()
