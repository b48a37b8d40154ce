"""ctypes binding of the C-ABI library (include/ftc_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails the error is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import arch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FTC_B200_LIB: A/B a differently built library from the tools/ scripts (same C-ABI); the default is the in-tree build
LIB_PATH = os.environ.get("FTC_B200_LIB") or os.path.join(_HERE, "lib", "libftc_b200.so")

FTC_MAX_STAGES = 8
FTC_MAX_HEADS = 9
PREC_F32, PREC_BF16 = 0, 1
GEMM_SIMT, GEMM_TCGEN05, GEMM_TCGEN05_IM2COL = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_GELU, ACT_SWIGLU = 0, 1, 2, 3
DT_F32, DT_BF16 = 0, 1
INPUT_NCHW_UNIT, INPUT_NHWC_255 = 0, 1
PART_EARLY, PART_REST = 1, 2


class StageCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("fused", "expand", "kernel", "stride", "cin", "cout", "layers")]


class DetectorConfig(C.Structure):
    _fields_ = [
        ("stem_out", C.c_int),
        ("n_stages", C.c_int),
        ("stages", StageCfg * FTC_MAX_STAGES),
        ("last_channel", C.c_int),
        ("n_heads", C.c_int),
        ("head_out", C.c_int * FTC_MAX_HEADS),
        ("head_names", (C.c_char * 32) * FTC_MAX_HEADS),
        ("height", C.c_int),
        ("width", C.c_int),
        ("precision", C.c_int),
        ("gemm_backend", C.c_int),
    ]


class TransformerConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("enc_input_dim", "embed_dim", "head_num", "enc_blocks", "dec_blocks", "max_enc_len",
                                       "max_dec_len", "precision", "gemm_backend")]


# every symbol include/ftc_b200.h declares: name -> (restype, argtypes)
_vp, _i, _f, _d, _sz, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_int64
SYMBOLS = {
    "ftc_version": (_i, []),
    "ftc_last_error": (C.c_char_p, []),
    "ftc_launch_count": (_i64, []),
    "ftc_detector_create": (_i, [C.POINTER(DetectorConfig), C.POINTER(_vp)]),
    "ftc_detector_destroy": (None, [_vp]),
    "ftc_detector_weight_bytes": (_sz, [_vp]),
    "ftc_detector_workspace_bytes": (_sz, [_vp, _i]),
    "ftc_detector_pack_weights": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(_i64), _vp, _sz, _vp]),
    "ftc_detector_forward": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ftc_detector_forward_part": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ftc_detector_set_input_format": (_i, [_vp, _i]),
    "ftc_detector_num_ops": (_i, [_vp]),
    "ftc_detector_forward_timed": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _sz, _vp, _i, C.POINTER(C.c_float),
                                        C.POINTER(C.c_double), C.POINTER(_i)]),
    "ftc_detector_tap": (_i, [_vp, _i, _i, _vp, C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "ftc_transformer_create": (_i, [C.POINTER(TransformerConfig), C.POINTER(_vp)]),
    "ftc_transformer_destroy": (None, [_vp]),
    "ftc_transformer_weight_bytes": (_sz, [_vp]),
    "ftc_transformer_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "ftc_transformer_pack_weights": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(_i64), _vp, _sz, _vp]),
    "ftc_transformer_logit_stride": (_i, []),
    "ftc_transformer_head_stride": (_i, []),
    "ftc_transformer_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "ftc_transformer_predict": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, C.POINTER(_i), C.POINTER(_i), _vp, _sz, _vp]),
    "ftc_transformer_predict_each": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "ftc_mask_predict_step": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "ftc_adamw_sf_chunk_elems": (_i, []),
    "ftc_adamw_sf_step": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _d, _d, _d, _d, _d, _vp]),
    "ftc_adamw_sf_step_dev": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ftc_radam_sf_step_dev": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ftc_radam_sf_step": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _d, _d, _d, _d, _d, _i, _vp]),
    "ftc_heatmap_loss_scratch_bytes": (_sz, []),
    "ftc_heatmap_loss": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "ftc_heatmap_loss_grad": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ftc_ce_rows": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "ftc_ce_rows_grad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "ftc_peak_decode_scratch_bytes": (_sz, [_i, _i, _i]),
    "ftc_peak_decode": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _f, _f, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ftc_peak_pick": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "ftc_op_conv2d": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "ftc_op_conv2d_wpack_bytes": (_sz, [_i, _i, _i]),
    "ftc_op_conv2d_dgrad": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "ftc_op_dwconv3x3": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "ftc_op_dwconv3x3_tiles": (_i, [_i, _i, _i, _i]),
    "ftc_op_se_fc": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "ftc_debug_set_trace": (_i, [_vp]),
    "ftc_debug_set_gemm_tuning": (_i, [_i, _i, _i, _i, _i]),
    "ftc_debug_bench_gemm": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_float)]),
    "ftc_debug_bench_conv3x3": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_float)]),
    "ftc_op_upsample2x": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ftc_op_dwconv3x3_se": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ftc_op_head_top_conv": (_i, [_vp, _i, _i, _i, C.POINTER(_i), _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ftc_train_reduce_scratch_bytes": (_sz, [_i64, _i]),
    "ftc_train_bn_stats": (_i, [_vp, _i, _i64, _i, _vp, _vp, _vp, _vp]),
    "ftc_train_bn_stats_running": (_i, [_vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp]),
    "ftc_train_bn_act": (_i, [_vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp]),
    "ftc_train_bn_act_bwd": (_i, [_vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp]),
    "ftc_train_bn_act_bwd_ld": (_i, [_vp, _vp, _i64, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp]),
    "ftc_train_upsample2x_bwd_ld": (_i, [_vp, _i64, _vp, _i, _i, _i, _i, _i, _vp]),
    "ftc_train_conv2d_wgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ftc_train_conv2d_wgrad_scratch_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "ftc_train_conv2d_wgrad_ws": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "ftc_train_conv2d_dgrad": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ftc_train_dwconv3x3": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ftc_train_dwconv3x3_dgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ftc_train_dwconv3x3_wgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ftc_train_spatial_sum": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "ftc_train_scale_bc": (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _vp]),
    "ftc_train_se_fc": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ftc_train_se_fc_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ftc_train_upsample2x_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ftc_train_layernorm": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _f, _vp]),
    "ftc_train_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "ftc_train_swiglu": (_i, [_vp, _vp, _vp, _i, _i64, _vp]),
    "ftc_train_swiglu_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i64, _vp]),
    "ftc_train_embed3": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _i64, _i, _vp]),
    "ftc_train_embed3_bwd": (_i, [_vp, _vp, _i, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ftc_train_attention_bwd_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "ftc_train_attention_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ftc_box_hists": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp]),
    "ftc_select_boxes_scratch_bytes": (_sz, [_i]),
    "ftc_select_boxes": (_i, [_vp, _vp, _i, _vp, _i, _vp, _d, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ftc_crop_sample_bytes": (_i, []),
    "ftc_crop_scratch_bytes": (_sz, [_i, _i]),
    "ftc_crop_batch": (_i, [_vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ftc_distort_scratch_bytes": (_sz, [_i]),
    "ftc_distort_batch": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ftc_debug_set_wgrad_mma": (_i, [_i]),
    "ftc_debug_set_bn_unroll": (_i, [_i]),
    "ftc_debug_set_wgrad_tc": (_i, [_i]),
    "ftc_page_maps": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp]),
    "ftc_op_attention": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libftc_b200.so (built by ``python -m findtextcenternet_b200.build``); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m findtextcenternet_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ftc_last_error().decode(errors="replace")
        raise RuntimeError(f"libftc_b200 {what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(load().ftc_launch_count())


def make_detector_config(model_size: str = "xl", precision: int = PREC_BF16, gemm_backend: int = GEMM_TCGEN05,
                         height: int = arch.HEIGHT, width: int = arch.WIDTH) -> DetectorConfig:
    stem, stages, last = arch.backbone_cfg(model_size)
    cfg = DetectorConfig()
    cfg.stem_out = stem
    cfg.n_stages = len(stages)
    for i, st in enumerate(stages):
        cfg.stages[i] = StageCfg(int(st.fused), st.expand, st.kernel, st.stride, st.cin, st.cout, st.layers)
    cfg.last_channel = last
    cfg.n_heads = len(arch.HEADS)
    for i, (name, od) in enumerate(arch.HEADS):
        cfg.head_out[i] = od
        cfg.head_names[i].value = name.encode()
    cfg.height, cfg.width = height, width
    cfg.precision, cfg.gemm_backend = precision, gemm_backend
    return cfg
