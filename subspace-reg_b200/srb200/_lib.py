"""ctypes binding of libsrb200.so (the C ABI declared in include/srb200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``csrc/Makefile``.  There is no
fallback: if the library is missing, or the device is not sm_100, every op raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRB200_LIB", os.path.join(_HERE, "libsrb200.so"))   # override: A/B timing of library builds

SR_EPI_ACT, SR_EPI_ACT_POOL2, SR_EPI_ACT_AVG, SR_EPI_RAW_STATS = 0, 1, 2, 3
SR_PULL_NONE, SR_PULL_FIXED, SR_PULL_PROJECT = 0, 1, 2
SR_OPT_SGD, SR_OPT_ADAM = 0, 1
SR_TRACE_COLS = 8

EXPORTS = [
    "sr_last_error", "sr_version", "sr_check_device", "sr_pack_input", "sr_bn_fold", "sr_pack_weight", "sr_conv",
    "sr_bn_finalize", "sr_bn_apply", "sr_subspace_factor_workspace_bytes", "sr_subspace_factor",
    "sr_head_workspace_bytes", "sr_head_run", "sr_eval_logits", "sr_semantic_pullers", "sr_linear_fwd",
    "sr_linear_bwd", "sr_sqdist", "sr_diff_scale", "sr_project_rows", "sr_score_logits", "sr_host_bernoulli", "sr_host_dropblock", "sr_conv_plan", "sr_pack_input_u8", "sr_global_avg", "sr_mse_grad", "sr_sgd_update", "sr_eval_workspace_bytes", "sr_train_block", "sr_backbone_eval_workspace_bytes", "sr_backbone_eval",
    "sr_mt_jump_table_bytes", "sr_mt_jump_table", "sr_host_mt_advance", "sr_device_bernoulli_workspace_bytes",
    "sr_device_bernoulli", "sr_dropblock_keep", "sr_fit_linear_map_workspace_bytes", "sr_fit_linear_map", "sr_head_plan",
]


class ConvPanel(C.Structure):
    _fields_ = [("act", C.c_void_p), ("wgt", C.c_void_p), ("cin_pad", C.c_int32), ("taps", C.c_int32),
                ("act_lo", C.c_void_p), ("wgt_lo", C.c_void_p)]


class ConvArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("cout", C.c_int32),
        ("n_panels", C.c_int32), ("panel", ConvPanel * 2),
        ("shift", C.c_void_p), ("residual", C.c_void_p), ("residual_lo", C.c_void_p), ("slope", C.c_float),
        ("epilogue", C.c_int32), ("out", C.c_void_p), ("out_lo", C.c_void_p), ("stats", C.c_void_p),
        ("weights_per_image", C.c_int32), ("skip_if_nonzero", C.c_void_p), ("max_cout_per_cta", C.c_int32),
    ]


class BnApplyArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("channels", C.c_int32),
        ("raw", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("res_raw", C.c_void_p), ("res_mean", C.c_void_p), ("res_invstd", C.c_void_p), ("res_gamma", C.c_void_p),
        ("res_beta", C.c_void_p), ("res_act", C.c_void_p), ("res_act_lo", C.c_void_p), ("lrelu", C.c_int32),
        ("slope", C.c_float), ("pool", C.c_int32), ("keep", C.c_void_p), ("keep_scale", C.c_float), ("out", C.c_void_p),
        ("out_lo", C.c_void_p), ("keep_scale_dev", C.c_void_p),
    ]


class MaskRegion(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("p", C.c_double), ("n", C.c_int64), ("out", C.c_void_p)]


class TrainBlockArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("cin_pad", C.c_int32), ("cout", C.c_int32),
        ("downsample", C.c_int32), ("x", C.c_void_p), ("x_lo", C.c_void_p), ("w", C.c_void_p * 4), ("w_lo", C.c_void_p * 4),
        ("gamma", C.c_void_p * 2), ("beta", C.c_void_p * 2), ("running_mean", C.c_void_p * 4), ("running_var", C.c_void_p * 4),
        ("eps", C.c_float), ("momentum", C.c_float), ("slope", C.c_float), ("stats", C.c_void_p), ("mean_invstd", C.c_void_p),
        ("raw", C.c_void_p * 4), ("h1", C.c_void_p), ("h1_lo", C.c_void_p), ("h2", C.c_void_p), ("h2_lo", C.c_void_p),
    ]


class EvalBlock(C.Structure):
    _fields_ = [("cout", C.c_int32), ("pool", C.c_int32), ("downsample", C.c_int32), ("reserved", C.c_int32),
                ("w1", C.c_void_p), ("w2", C.c_void_p), ("w3", C.c_void_p), ("wd", C.c_void_p),
                ("w1_lo", C.c_void_p), ("w2_lo", C.c_void_p), ("w3_lo", C.c_void_p), ("wd_lo", C.c_void_p),
                ("s1", C.c_void_p), ("s2", C.c_void_p), ("s3", C.c_void_p)]


class BackboneEvalArgs(C.Structure):
    _fields_ = [("n_blocks", C.c_int32), ("blocks", C.POINTER(EvalBlock)), ("batch", C.c_int32), ("height", C.c_int32),
                ("width", C.c_int32), ("cin_pad", C.c_int32), ("x", C.c_void_p), ("x_lo", C.c_void_p), ("slope", C.c_float),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("features", C.c_void_p)]


class HeadArgs(C.Structure):
    _fields_ = [
        ("feat", C.c_void_p), ("dim", C.c_int32), ("n_support", C.c_int32), ("support_row0", C.c_int32),
        ("n_memory", C.c_int32), ("memory_row0", C.c_int32), ("labels_support", C.c_void_p),
        ("labels_memory", C.c_void_p), ("weight", C.c_void_p), ("n_classes", C.c_int32), ("opt_state", C.c_void_p),
        ("base_weight", C.c_void_p), ("n_base", C.c_int32), ("reserve_weight", C.c_void_p),
        ("n_prev_novel", C.c_int32), ("n_new", C.c_int32), ("pull_mode", C.c_int32), ("pull", C.c_void_p),
        ("q_rows", C.c_int32), ("lmbd_base", C.c_float), ("lmbd_novel", C.c_float), ("gamma", C.c_float),
        ("optimizer", C.c_int32), ("lr", C.c_float), ("momentum", C.c_float), ("weight_decay", C.c_float),
        ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float), ("step0", C.c_int32),
        ("max_epochs", C.c_int32), ("epoch0", C.c_int32), ("stable", C.c_int32), ("stable_epochs", C.c_int32),
        ("stable_count0", C.c_int32), ("min_novel_epochs", C.c_int32), ("max_novel_epochs", C.c_int32),
        ("convergence_epsilon", C.c_double), ("target_train_loss", C.c_double), ("prev_loss", C.c_float),
        ("loss_trace", C.c_void_p), ("status", C.c_void_p), ("resume_status", C.c_void_p), ("logits_support", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("cta_budget", C.c_int32), ("reserved0", C.c_int32),
    ]


class EvalArgs(C.Structure):
    _fields_ = [
        ("feat", C.c_void_p), ("weight", C.c_void_p), ("labels", C.c_void_p),
        ("n", C.c_int32), ("dim", C.c_int32), ("n_classes", C.c_int32),
        ("logits", C.c_void_p), ("pred", C.c_void_p), ("counts", C.c_void_p), ("loss_sum", C.c_void_p),
        ("confusion", C.c_void_p), ("conf_dim", C.c_int32), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


_lib = None


def load():
    """dlopen libsrb200.so (once) and declare signatures.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "srb200: %s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.sr_last_error.restype = C.c_char_p
    lib.sr_last_error.argtypes = []
    lib.sr_version.restype = i32
    lib.sr_check_device.restype = i32
    lib.sr_check_device.argtypes = [i32]
    lib.sr_pack_input.restype = i32
    lib.sr_pack_input.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.sr_bn_fold.restype = i32
    lib.sr_bn_fold.argtypes = [vp, vp, vp, vp, f32, vp, vp, i32, vp]
    lib.sr_pack_weight.restype = i32
    lib.sr_pack_weight.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.sr_conv.restype = i32
    lib.sr_conv.argtypes = [C.POINTER(ConvArgs), vp]
    lib.sr_bn_finalize.restype = i32
    lib.sr_bn_finalize.argtypes = [vp, i64, f32, f32, vp, vp, vp, vp, i32, vp]
    lib.sr_bn_apply.restype = i32
    lib.sr_bn_apply.argtypes = [C.POINTER(BnApplyArgs), vp]
    lib.sr_backbone_eval_workspace_bytes.restype = i64
    lib.sr_backbone_eval_workspace_bytes.argtypes = [C.POINTER(BackboneEvalArgs)]
    lib.sr_backbone_eval.restype = i32
    lib.sr_backbone_eval.argtypes = [C.POINTER(BackboneEvalArgs), vp]
    lib.sr_train_block.restype = i32
    lib.sr_train_block.argtypes = [C.POINTER(TrainBlockArgs), vp]
    lib.sr_subspace_factor_workspace_bytes.restype = i64
    lib.sr_subspace_factor_workspace_bytes.argtypes = [i32, i32]
    lib.sr_subspace_factor.restype = i32
    lib.sr_subspace_factor.argtypes = [vp, i32, i32, vp, vp, vp, i64, vp]
    lib.sr_head_workspace_bytes.restype = i64
    lib.sr_head_workspace_bytes.argtypes = [C.POINTER(HeadArgs)]
    lib.sr_head_plan.restype = i32
    lib.sr_head_plan.argtypes = [C.POINTER(HeadArgs), C.POINTER(C.c_int32)]
    lib.sr_head_run.restype = i32
    lib.sr_head_run.argtypes = [C.POINTER(HeadArgs), vp]
    lib.sr_eval_workspace_bytes.restype = i64
    lib.sr_eval_workspace_bytes.argtypes = [i32, i32, i32]
    lib.sr_eval_logits.restype = i32
    lib.sr_eval_logits.argtypes = [C.POINTER(EvalArgs), vp]
    lib.sr_mse_grad.restype = i32
    lib.sr_mse_grad.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.sr_sgd_update.restype = i32
    lib.sr_sgd_update.argtypes = [vp, vp, i64, f32, f32, vp]
    lib.sr_global_avg.restype = i32
    lib.sr_global_avg.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.sr_host_bernoulli.restype = i64
    lib.sr_host_bernoulli.argtypes = [vp, i64, i32, C.c_double, i64, vp]
    lib.sr_pack_input_u8.restype = i32
    lib.sr_pack_input_u8.argtypes = [vp, vp, vp, i32, i32, i32, i32, C.POINTER(f32), C.POINTER(f32), i32, vp, vp, i32, vp]
    lib.sr_conv_plan.restype = i32
    lib.sr_conv_plan.argtypes = [C.POINTER(ConvArgs), C.POINTER(C.c_int32)]
    lib.sr_host_dropblock.restype = i64
    lib.sr_host_dropblock.argtypes = [vp, i64, i32, i32, i32, vp]
    lib.sr_mt_jump_table_bytes.restype = i64
    lib.sr_mt_jump_table_bytes.argtypes = [i32]
    lib.sr_mt_jump_table.restype = i32
    lib.sr_mt_jump_table.argtypes = [vp, i32]
    lib.sr_host_mt_advance.restype = i32
    lib.sr_host_mt_advance.argtypes = [vp, i64, i64, vp, i32]
    lib.sr_device_bernoulli_workspace_bytes.restype = i64
    lib.sr_device_bernoulli_workspace_bytes.argtypes = [i64]
    lib.sr_device_bernoulli.restype = i32
    lib.sr_device_bernoulli.argtypes = [vp, i64, C.POINTER(MaskRegion), i32, vp, vp, i32, vp, i64, vp]
    lib.sr_fit_linear_map_workspace_bytes.restype = i64
    lib.sr_fit_linear_map_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.sr_fit_linear_map.restype = i32
    lib.sr_fit_linear_map.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, f32, f32, vp, vp, i64, vp]
    lib.sr_dropblock_keep.restype = i32
    lib.sr_dropblock_keep.argtypes = [vp, i64, i32, i32, i32, vp, vp, vp]
    lib.sr_score_logits.restype = i32
    lib.sr_score_logits.argtypes = [C.POINTER(EvalArgs), vp]
    lib.sr_semantic_pullers.restype = i32
    lib.sr_semantic_pullers.argtypes = [vp, vp, vp, i32, i32, i32, i32, f32, i32, vp, vp]
    lib.sr_linear_fwd.restype = i32
    lib.sr_linear_fwd.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]
    lib.sr_linear_bwd.restype = i32
    lib.sr_linear_bwd.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    lib.sr_sqdist.restype = i32
    lib.sr_sqdist.argtypes = [vp, vp, i64, vp, vp]
    lib.sr_diff_scale.restype = i32
    lib.sr_diff_scale.argtypes = [vp, vp, i64, f32, vp, vp, vp, vp]
    lib.sr_project_rows.restype = i32
    lib.sr_project_rows.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().sr_last_error()
        raise RuntimeError("srb200 %s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
