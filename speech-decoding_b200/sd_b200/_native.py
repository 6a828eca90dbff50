"""ctypes binding of include/sd_b200.h (libsd_b200.so, built in-tree by
csrc/Makefile).  There is no fallback: if the library is missing or a call
fails, a RuntimeError is raised."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libsd_b200.so")

SD_F32, SD_BF16, SD_TF32 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_GLU = 0, 1, 2
OUT_BTC, OUT_NCT_F32 = 0, 1
SA_MPARTS = 32          # SD_SA_MPARTS
IMPL_AUTO, IMPL_SIMT, IMPL_TC, IMPL_TC_1CTA, IMPL_TC_WS = 0, 1, 2, 3, 4

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float


class ConvArgs(C.Structure):
    _fields_ = [("inp", vp), ("w", vp), ("bias", vp), ("res", vp), ("widx", vp), ("out", vp),
                ("preact", vp), ("stats", vp), ("rownorm2", vp),
                ("B", i32), ("T", i32), ("K", i32), ("Kp", i32), ("N", i32), ("Np", i32),
                ("taps", i32), ("dil", i32), ("G", i32),
                ("act", i32), ("out_mode", i32), ("dtype", i32), ("affine", vp), ("in_lo", vp), ("w_lo", vp),
                ("bnr_y", vp), ("bnr_ss", vp)]


class AdamEntry(C.Structure):
    _fields_ = [("param", vp), ("grad", vp), ("exp_avg", vp), ("exp_avg_sq", vp), ("n", i64),
                ("step_size", f32), ("bias_correction2_sqrt", f32)]


class WgradArgs(C.Structure):
    _fields_ = [("dout", vp), ("inp", vp), ("dw", vp), ("dbias", vp), ("sample_order", vp),
                ("group_offsets", vp),
                ("B", i32), ("T", i32), ("K", i32), ("Kp", i32), ("N", i32), ("Np", i32),
                ("taps", i32), ("dil", i32), ("G", i32),
                ("gs", i64), ("sn", i64), ("sk", i64), ("sj", i64), ("dtype", i32),
                ("workspace", vp), ("workspace_bytes", i64), ("dout_lo", vp), ("in_lo", vp)]


class PackEntry(C.Structure):
    _fields_ = [("w", vp), ("wf", vp), ("wd", vp), ("N", i32), ("K", i32), ("taps", i32),
                ("Np", i32), ("Kp", i32), ("dtype", i32)]


# name -> argtypes (every symbol include/sd_b200.h declares; tests check the export list)
SIGNATURES = {
    "sd_abi_version": [],
    "sd_device_info": [C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
    "sd_set_impl": [i32],
    "sd_set_sm_limit": [i32],
    "sd_nct_to_btc": [vp, vp, i32, i32, i32, i32, i32, vp],
    "sd_nct_to_btc_bf16in": [vp, vp, i32, i32, i32, i32, i32, vp],
    "sd_btc_to_nct": [vp, vp, i32, i32, i32, i32, i32, vp],
    "sd_pack_weight": [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "sd_pack_weights": [vp, i32, i32, vp],
    "sd_sa_weights_fwd": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "sd_sa_weights_bwd": [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "sd_conv_fwd": [C.POINTER(ConvArgs), vp],
    "sd_conv_wgrad": [C.POINTER(WgradArgs), vp],
    "sd_colstats": [vp, vp, i64, i32, i32, vp],
    "sd_colsum_add": [vp, vp, vp, i64, i32, i32, i32, vp],
    "sd_bn_finalize": [vp, i32, i32, i64, vp, vp, vp, vp, vp, f32, f32, i32, vp, vp],
    "sd_bn_gelu_fwd": [vp, vp, vp, i64, i32, i32, vp],
    "sd_bn_gelu_bwd_reduce": [vp, vp, vp, vp, i64, i32, i32, vp],
    "sd_bn_gelu_bwd_reduce_g": [vp, vp, vp, vp, i64, i32, i32, vp],
    "sd_bn_bwd_apply": [vp, vp, vp, vp, vp, vp, i64, i64, f32, i32, i32, i32, i32, vp],
    "sd_bn_bwd_apply_g": [vp, vp, vp, vp, vp, vp, i64, i64, f32, i32, i32, i32, i32, vp],
    "sd_glu_fwd": [vp, vp, i64, i32, i32, i32, i32, vp],
    "sd_glu_bwd": [vp, vp, vp, i64, i32, i32, i32, i32, vp],
    "sd_gelu_bwd": [vp, vp, i64, i32, i32, vp],
    "sd_gelu_bwd_nct": [vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "sd_rownorm2": [vp, vp, i32, i64, vp],
    "sd_clip_dots": [vp, vp, vp, i32, i32, i64, vp],
    "sd_clip_phase1": [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp],
    "sd_clip_phase2": [vp, vp, vp, vp, vp, vp, f32, i32, vp, vp, vp, vp, i32, i32, vp],
    "sd_clip_dots_workspace_bytes": [i32, i32, i64],
    "sd_clip_dots_tc": [vp, vp, vp, vp, i32, i32, i64, vp],
    "sd_clip_dz_tc": [vp, vp, vp, vp, vp, vp, i32, i32, i64, vp],
    "sd_collate_preproc": [vp, vp, i64, i32, i32, C.c_float, i32, vp],
    "sd_adam_step": [vp, i32, i32, f32, f32, f32, f32, vp],
    "sd_cast_rows_bf16": [vp, vp, vp, i32, i64, vp],
    "sd_rownorm2_bf16": [vp, vp, i32, i64, vp],
    "sd_clip_merge_row_stats": [vp, i32, i32, vp, vp],
    "sd_clip_coef_t_bf16": [vp, vp, i32, i32, i32, vp],
    "sd_clip_dots_tc_bf16": [vp, vp, vp, vp, i32, i32, i64, vp],
    "sd_clip_dz_tc_bf16": [vp, vp, vp, vp, vp, vp, i32, i32, i64, vp],
    "sd_clip_dz": [vp, vp, vp, vp, vp, vp, i32, i32, i64, vp],
    "sd_tf32_split": [vp, vp, vp, i64, vp],
    "sd_peer_alloc": [C.POINTER(vp), i64],
    "sd_peer_free": [vp],
    "sd_ipc_handle_bytes": [],
    "sd_ipc_get_handle": [vp, vp],
    "sd_ipc_open_handle": [vp, C.POINTER(vp)],
    "sd_ipc_close_handle": [vp],
    "sd_memcpy_async": [vp, vp, i64, vp],
    "sd_copy_small": [vp, vp, i64, vp],
    "sd_peer_wait_flags": [vp, i32, i32, vp],
    "sd_peer_exchange": [vp, i64, vp, i32, i32, i64, i32, vp, vp],
    "sd_sum_rows": [vp, vp, i32, i32, i32, vp],
}

_lib = None


def lib():
    """Load libsd_b200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "sd_b200: %s not found -- build it with `make -C speech-decoding_b200/csrc` "
                "(or __graft_entry__.build()); there is no CPU or PyTorch fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.sd_last_error.restype = C.c_char_p
        l.sd_last_error.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = i64 if name == "sd_clip_dots_workspace_bytes" else i32       # (sd_ipc_handle_bytes returns the size)
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("sd_b200 %s failed: %s" % (what, lib().sd_last_error().decode()))


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError("sd_b200 %s failed: %s" % (name, lib().sd_last_error().decode()))
