"""ctypes binding of libffr_sm100.so (include/ffr_sm100.h). There is no fallback: if the library is missing or a
call fails, a RuntimeError is raised."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libffr_sm100.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ffr_sm100.h")
PROBE_LIB_PATH = os.path.join(_HERE, "lib", "libffr_sm100_probe.so")
PROBE_HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ffr_sm100_probe.h")

_lib = None

_p = ctypes.c_void_p
_i = ctypes.c_int
_u32 = ctypes.c_uint32
_i64 = ctypes.c_int64

_f = ctypes.c_float


class EPI:
    """Epilogue flags of ffr_conv_gemm / ffr_conv_gemm_ex (FFR_EPI_* in include/ffr_sm100.h)."""
    BIAS, BORDER_BIAS, PRELU, GEOM = 1 << 0, 1 << 1, 1 << 2, 1 << 3
    POOL, OUT_S2D, OUT_F32_ATOMIC, SIGMOID = 1 << 4, 1 << 5, 1 << 6, 1 << 7
    SCATTER, RESIDUAL, STATS, OUT_F32 = 1 << 8, 1 << 9, 1 << 10, 1 << 11
    COSFACE, PIXMAJOR, PIX_DGRAD = 1 << 12, 1 << 13, 1 << 14
    MUL_DSIG, RES_F16, OUT_F16 = 1 << 15, 1 << 16, 1 << 17


class ConvGemmDesc(ctypes.Structure):
    """ffr_conv_gemm_desc (include/ffr_sm100.h)."""
    _fields_ = [("a", _p), ("a_rows", _i64), ("a_cols", _i), ("a_ld", _i),
                ("wp", _p), ("Cin", _i), ("Cout", _i), ("ntaps", _i),
                ("tap_row_shift", _p), ("tap_ch_off", _p),
                ("M", _i), ("rows_per_img", _i), ("Wp", _i), ("S", _i), ("h0", _i), ("n_img", _i),
                ("flags", _u32),
                ("bias", _p), ("slope", _p),
                ("out", _p), ("ldo", _i), ("s2d_So", _i),
                ("pool", _p), ("out_f32", _p),
                ("res", _p), ("ldres", _i),
                ("stats", _p), ("stats_part", _p),
                ("num_splits", _i),
                ("scatter", _p), ("scatter_n", _i), ("out_rows_per_img", _i),
                ("b_rows_per_mtile", _i), ("b_mtile_div", _i),
                ("a_hilo", _i), ("a_lo_off", _i),
                ("f16", _i)]


class PrepTrainDesc(ctypes.Structure):
    """ffr_prep_train_desc (include/ffr_sm100.h)."""
    _fields_ = ([(k, _p) for k in ("x", "w0", "b0", "slope1", "slope4", "slope7", "A1", "c1", "A2", "c2", "w8", "b8")] +
                [("s0_h", _p), ("s0_ld", _i), ("s0_lo", _i), ("s0_b", _p), ("s0_ldb", _i),
                 ("cm_h", _p), ("cm_ld", _i), ("cm_lo", _i), ("cm_b", _p), ("cm_ldb", _i),
                 ] +
                [(k, _p) for k in ("g0", "g1", "g2", "h7b", "xk", "mch2", "x3", "inv_c", "tmat", "ss_space")])


_SIGNATURES = {
    "ffr_conv_gemm_ex": (_i, [ctypes.POINTER(ConvGemmDesc), _p]),
    "ffr_bn_finalize": (_i, [_p, _i, _i, _i, _i, _i, _i, _f, _f, _p, _p, _p, _p, _p]),
    "ffr_bn_act_fwd": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _i, _p, _i, _i, _p, _i, _p, _i, _i, _p, _i, _i, _i, _i, _i, _p]),
    "ffr_nchw_to_h9_f32": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "ffr_bn_act_bwd_partial_rows": (_i, [_i, _i]),
    "ffr_bn_act_bwd": (_i, [_p, _i, _i, _p, _i, _p, _i, _i, _p, _i, _f, _p, _i, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p,
                            _i, _i, _p, _i, _i, _i, _i, _p]),
    "ffr_h9_avgpool": (_i, [_p, _i, _p, _i, _i, _i, _p]),
    "ffr_h9_avgpool_bf16": (_i, [_p, _i, _p, _i, _i, _i, _p]),
    "ffr_wgrad_workspace_floats": (_i64, [_i, _i, _i, _i, _i]),
    "ffr_wgrad": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "ffr_pack_conv3x3_f16": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "ffr_recnet_prep_train": (_i, [ctypes.POINTER(PrepTrainDesc), _i, _p]),
    "ffr_fc_scatter": (_i, [_p, _p, _i, _i, _p, _i, _i, _p]),
    "ffr_chan_compose": (_i, [_p] * 14 + [_p]),
    "ffr_chan_compose_bwd": (_i, [_p] * 18 + [_i, _p]),
    "ffr_feat_space_train": (_i, [_p, _p, _p, _i, _i, _p, _i, _p, _i, _i, _p]),
    "ffr_feat_space_bwd": (_i, [_p, _p, _p, _i, _p, _i, _p, _i, _p]),
    "ffr_fc_bwd_gather": (_i, [_p, _i, _p, _i, _p]),
    "ffr_chan_bwd_part_floats": (_i, []),
    "ffr_chan_bwd": (_i, [_p] * 20 + [_i, _i, _p]),
    "ffr_selfsim_channel_pack": (_i, [_p, _i, _p, _i, _i, _p, _p, _p, _p, _p]),
    "ffr_selfsim_channel_bwd": (_i, [_p, _p, _i, _p, _f, _p, _i, _i, _p]),
    "ffr_sumsq_reduce": (_i, [_p, _i, _i, _i, _p, _p]),
    "ffr_selfsim_space_loss": (_i, [_p, _i, _p, _i, _i, _f, _p, _p, _i, _p]),
    "ffr_triplet_identity": (_i, [_p, _p, _p, _p, _i, _f, _f, _f, _p, _p, _p, _p]),
    "ffr_loss_finalize": (_i, [_p, _p, _p, _p, _i, _i, _f, _f, _f, _f, _p, _p]),
    "ffr_add3_f32": (_i, [_p, _p, _p, _p, _i64, _p]),
    "ffr_cosface_ce_bwd_grouped": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i, _f, _f, _p, _p, _p]),
    "ffr_version": (_i, []),
    "ffr_last_error": (ctypes.c_char_p, []),
    "ffr_launch_count": (ctypes.c_longlong, []),
    "ffr_conv_gemm": (_i, [_p, _i64, _i, _i, _p, _i, _i, _i, _p, _p, _i, _i, _i, _i, _i, _i, _u32, _p, _p, _p, _i, _i,
                           _p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _p]),
    "ffr_recnet_prep": (_i, [_p, _i] + [_p] * 15 + [_p]),
    "ffr_recnet_convlayer_fwd": (_i, [_p, _i, _i, _p, _i, _p, _p, _p, _i, _i, _p, _i, _p, _i, _i, _p, _p, _p]),
    "ffr_self_similarity": (_i, [_p, _i, _p, _p, _p]),
    "ffr_feat_space": (_i, [_p, _p, _p, _p, _i, _p]),
    "ffr_feat_space_xt": (_i, [_p, _p, _p, _p, _i, _p]),
    "ffr_rows_to_nchw": (_i, [_p, _i, _i, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ffr_pack_conv3x3": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "ffr_gallery_cosine": (_i, [_p, _i, _p, _i, _i, _p, _p, _p]),
    "ffr_roc_hist": (_i, [_p, _i, _i, _i, _p, _p, _p, _i, _p, _p]),
    "ffr_cosface_pack": (_i, [_p, _i, _i, _i, _p, _p, _i, _p]),
    "ffr_cosface_ce_fwd": (_i, [_p, _i, _p, _i, _i, _p, ctypes.c_float, ctypes.c_float, _p, _p, _p, _p, _p, _p]),
    "ffr_cosface_ce_finish": (_i, [_p, _p, _p, _i, ctypes.c_float, _p, _p, _p]),
    "ffr_cosface_ce_bwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, ctypes.c_float, ctypes.c_float, _p, _p, _p]),
    "ffr_normalize_bwd": (_i, [_p, _p, _i, _p, _p]),
    "ffr_clip_adam": (_i, [_p, _p, _i, _p, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                           ctypes.c_float, _p]),
    "ffr_pair_cosine": (_i, [_p, _p, _p, _i, _i, _p]),
    "ffr_threshold_sweep": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ffr_scale_f32": (_i, [_p, _p, _i64, ctypes.c_float, _p]),
    "ffr_conv3x3_bnpre_prelu_fwd": (_i, [_p, _i, _i, _i, _p, _i, _p, _p, _p, _i, _p]),
    "ffr_conv3x3_bn_pool_fwd": (_i, [_p, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p]),
    "ffr_conv1x1_bn_fwd": (_i, [_p, _i, _i, _i, _p, _i, _p, _p, _p]),
    "ffr_subsample2": (_i, [_p, _p, _i, _i, _i, _p]),
    "ffr_stem_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _p]),
    "ffr_stem_u8_fwd": (_i, [_p, _p, _i, _p, _p, _p, _p, _i, _i, _p]),
    "ffr_se_pool_part_floats": (_i64, [_i, _i, _i]),
    "ffr_se_gate_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "ffr_se_residual_fwd": (_i, [_p, _p, _p, _i, _p, _i, _i, _i, _p]),
    "ffr_se_gate_residual_fwd": (_i, [_p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _p]),
    "ffr_head_workspace_floats": (_i64, [_i, _i, _i]),
    "ffr_export_nchw_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "ffr_head_fwd": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ffr_debug_set_window": (_i, [_i]),
    "ffr_debug_set_pair": (None, [_i]),
    "ffr_debug_set_pdl": (None, [_i]),
    "ffr_debug_set_lean_epilogue": (None, [_i]),
    "ffr_debug_set_stem_strip": (None, [_i]),
    "ffr_debug_set_streamk": (None, [_i]),
    "ffr_debug_set_prep_mma": (None, [_i]),
    "ffr_debug_last_streamk": (_i, []),
    "ffr_conv_scratch_bytes": (ctypes.c_longlong, []),
    "ffr_set_conv_scratch": (_i, [_p, ctypes.c_longlong]),
    "ffr_debug_mn_probe": (_i, [_p, _p, _p, _i, _i, _p]),
    "ffr_debug_set_counters": (_i, [_p]),
    "ffr_debug_set_wgrad_splits": (None, [_i]),
    "ffr_pixmajor_profitable": (_i, [_i]),
    "ffr_debug_set_pixmajor": (None, [_i]),
    "ffr_debug_set_pixmajor_backbone": (None, [_i]),
    "ffr_debug_mma_bench": (_i, [_p, _i, _i, _i, _i, _i, _p]),
    "ffr_debug_rowshift_probe": (_i, [_p, _p, _p, _i, _i, _p]),
}


def declared_symbols(header=None):
    """Names of every function a header under include/ declares (used by the CPU-side export test)."""
    with open(header or HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"FFR_API\s+[\w\s\*]+?\b(ffr_\w+)\s*\(", text)))


_probe = None


def load_probe():
    """The separate debug library with the hardware probes / micro-benchmarks (csrc/probe.cu)."""
    global _probe
    if _probe is None:
        if not os.path.exists(PROBE_LIB_PATH):
            raise RuntimeError("libffr_sm100_probe.so not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(PROBE_LIB_PATH)
        for name in declared_symbols(PROBE_HEADER_PATH):
            fn = getattr(lib, name)
            if name in _SIGNATURES:
                fn.restype, fn.argtypes = _SIGNATURES[name]
        _probe = lib
    return _probe


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libffr_sm100.so not built (%s missing): run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU / PyTorch fallback for this path" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name in declared_symbols():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        if name in _SIGNATURES:
            fn.restype, fn.argtypes = _SIGNATURES[name]
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().ffr_last_error()
        raise RuntimeError("libffr_sm100 %s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


# Parameter tensors are updated in place by the fused optimizer kernel through raw pointers, which torch's
# per-tensor version counters do not see; packed-weight caches key on this generation counter as well.
_weights_generation = 0


def weights_generation():
    return _weights_generation


def bump_weights_generation():
    global _weights_generation
    _weights_generation += 1


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
