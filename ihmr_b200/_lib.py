"""ctypes binding of libihmr_b200.so (the C ABI declared in include/ihmr_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing the import of any
compute entry point raises, and every call checks the integer status the ABI returns.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IHMR_B200_LIB", os.path.join(_HERE, "_lib", "libihmr_b200.so"))   # override: experiments only

EXPORTS = (
    "ihmr_last_error", "ihmr_abi_version", "ihmr_model_create", "ihmr_model_destroy",
    "ihmr_model_update_shapedirs", "ihmr_model_set_sdf_conventions", "ihmr_mano_workspace_bytes", "ihmr_mano_forward",
    "ihmr_mano_backward", "ihmr_sdf_workspace_bytes", "ihmr_sdf_loss", "ihmr_sdf_loss_exact", "ihmr_opt_workspace_bytes", "ihmr_opt_stage",
    "ihmr_opt_final", "ihmr_opt_value_and_grad", "ihmr_launch_count", "ihmr_opt_profile_iteration", "ihmr_sdf_stats", "ihmr_gemm_tf32x3", "ihmr_gemm_reference_fp32", "ihmr_eval_metrics", "ihmr_measure_fp32_peak", "ihmr_select_snapshots", "ihmr_opt_criteria", "ihmr_mlp_input", "ihmr_linear", "ihmr_mlp_apply", "ihmr_select_better",
)
KERNEL_CLASSES = ("pose_prep", "blend_fwd", "skin_fwd", "sdf", "frame_loss", "skin_bwd", "blend_bwd", "pose_bwd", "step")

P_CAM, P_TRANS, P_R_ORIENT, P_R_POSE, P_L_ORIENT, P_L_POSE, P_R_SHAPE, P_L_SHAPE = (1, 2, 4, 8, 16, 32, 64, 128)
LOSS_IDS = {"joints_3d_loss_p": 0, "collision_loss": 1, "joints_2d_loss_p": 2}
OPTIMIZERS = {"adam": 0, "sgd": 1}
STAGE_GENERIC_KERNELS = 1
PARAM_MASKS = {
    "pred_cam_params": P_CAM, "pred_hand_trans": P_TRANS, "pred_right_orient": P_R_ORIENT,
    "pred_right_pose_params": P_R_POSE, "pred_left_orient": P_L_ORIENT,
    "pred_left_pose_params": P_L_POSE, "pred_right_shape_params": P_R_SHAPE,
    "pred_left_shape_params": P_L_SHAPE,
}


class Stage(C.Structure):
    _fields_ = [
        ("update_mask", C.c_uint32), ("lr", C.c_float), ("epoch", C.c_int32),
        ("w_joints_2d", C.c_float), ("w_joints_3d", C.c_float), ("w_trans", C.c_float),
        ("w_shape_reg", C.c_float), ("w_collision", C.c_float), ("w_finger_reg", C.c_float),
        ("n_filters", C.c_int32), ("filter_loss", C.c_int32 * 4), ("filter_percent", C.c_float * 4),
        ("select_loss", C.c_int32), ("flags", C.c_uint32),
    ]


class Targets(C.Structure):
    _fields_ = [
        ("init_joints_2d", C.c_void_p), ("init_joints_3d", C.c_void_p), ("init_hand_trans_j", C.c_void_p),
        ("gt_joints_3d", C.c_void_p), ("hand_type_array", C.c_void_p),
    ]


class IhmrError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Loads the library (once) and declares every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IhmrError(
            f"{LIB_PATH} is missing. ihmr_b200 has no CPU or PyTorch fallback: build the CUDA "
            "library first with `python -m ihmr_b200.build` (needs nvcc, targets sm_100a).")
    lib = C.CDLL(LIB_PATH)
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    lib.ihmr_last_error.restype = C.c_char_p
    lib.ihmr_last_error.argtypes = []
    lib.ihmr_abi_version.restype = i32
    lib.ihmr_abi_version.argtypes = []
    lib.ihmr_model_create.restype = i32
    lib.ihmr_model_create.argtypes = [vp] * 9 + [i32, C.POINTER(vp)]
    lib.ihmr_model_destroy.restype = None
    lib.ihmr_model_destroy.argtypes = [vp]
    lib.ihmr_model_update_shapedirs.restype = i32
    lib.ihmr_model_update_shapedirs.argtypes = [vp, vp, vp]
    lib.ihmr_model_set_sdf_conventions.restype = i32
    lib.ihmr_model_set_sdf_conventions.argtypes = [vp, C.c_float, i32]
    lib.ihmr_mano_workspace_bytes.restype = sz
    lib.ihmr_mano_workspace_bytes.argtypes = [i32]
    lib.ihmr_mano_forward.restype = i32
    lib.ihmr_mano_forward.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.ihmr_mano_backward.restype = i32
    lib.ihmr_mano_backward.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.ihmr_sdf_workspace_bytes.restype = sz
    lib.ihmr_sdf_workspace_bytes.argtypes = [i32]
    lib.ihmr_sdf_loss.restype = i32
    lib.ihmr_sdf_loss.argtypes = [vp, i32, vp, vp, vp, vp, vp, f32, vp, sz, vp]
    lib.ihmr_sdf_loss_exact.restype = i32
    lib.ihmr_sdf_loss_exact.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.ihmr_opt_workspace_bytes.restype = sz
    lib.ihmr_opt_workspace_bytes.argtypes = [i32]
    lib.ihmr_opt_stage.restype = i32
    lib.ihmr_opt_stage.argtypes = [vp, i32, i32, vp, C.POINTER(Targets), C.POINTER(Stage), i32, i32, vp, sz, vp]
    lib.ihmr_opt_final.restype = i32
    lib.ihmr_opt_final.argtypes = [vp, i32, vp, C.POINTER(Targets), vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.ihmr_opt_value_and_grad.restype = i32
    lib.ihmr_opt_value_and_grad.argtypes = [vp, i32, i32, vp, C.POINTER(Targets), C.POINTER(Stage), vp, vp, vp, sz, vp]
    for fn in (lib.ihmr_gemm_tf32x3, lib.ihmr_gemm_reference_fp32):
        fn.restype = i32
        fn.argtypes = [i32, i32, i32, vp, i32, vp, i32, vp, i32, vp]
    lib.ihmr_eval_metrics.restype = i32
    lib.ihmr_eval_metrics.argtypes = [i32, vp, vp, vp, vp, vp, vp]
    lib.ihmr_opt_criteria.restype = i32
    lib.ihmr_opt_criteria.argtypes = [vp, i32, vp, C.POINTER(Targets), f32, f32, vp, vp, sz, vp]
    lib.ihmr_mlp_input.restype = i32
    lib.ihmr_mlp_input.argtypes = [i32, vp, vp, vp, vp]
    lib.ihmr_linear.restype = i32
    lib.ihmr_linear.argtypes = [i32, i32, i32, vp, i32, vp, vp, i32, vp, i32, vp]
    lib.ihmr_mlp_apply.restype = i32
    lib.ihmr_mlp_apply.argtypes = [i32, vp, i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp, vp, vp]
    lib.ihmr_select_better.restype = i32
    lib.ihmr_select_better.argtypes = [i32, vp, vp, C.POINTER(Stage), vp, vp, vp, vp]
    lib.ihmr_select_snapshots.restype = i32
    lib.ihmr_select_snapshots.argtypes = [i32, i32, vp, C.POINTER(Stage), vp, vp]
    lib.ihmr_measure_fp32_peak.restype = i32
    lib.ihmr_measure_fp32_peak.argtypes = [vp, C.POINTER(C.c_float), vp, vp]
    lib.ihmr_sdf_stats.restype = i32
    lib.ihmr_sdf_stats.argtypes = [vp, i32, vp, vp, vp, vp, sz, vp]
    lib.ihmr_launch_count.restype = C.c_ulonglong
    lib.ihmr_launch_count.argtypes = []
    lib.ihmr_opt_profile_iteration.restype = i32
    lib.ihmr_opt_profile_iteration.argtypes = [vp, i32, i32, vp, C.POINTER(Targets), C.POINTER(Stage), C.POINTER(C.c_float), vp, sz, vp]
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().ihmr_last_error().decode("utf-8", "replace")
        raise IhmrError(f"{what} failed with status {status}: {msg}")


def make_stage(stage: dict) -> Stage:
    """Converts one strategy stage dict (src/strategies/opt_default.py layout) into the ABI struct."""
    s = Stage()
    mask = 0
    for name in stage["update_params"]:
        if name not in PARAM_MASKS:
            raise IhmrError(f"unknown update_params entry {name!r}")
        mask |= PARAM_MASKS[name]
    lw = stage["loss_weights"]
    s.update_mask, s.lr, s.epoch = mask, float(stage["lr"]), int(stage["epoch"])
    s.w_joints_2d, s.w_joints_3d = float(lw["joints_2d_loss"]), float(lw["joints_3d_loss"])
    s.w_trans, s.w_shape_reg = float(lw["trans_loss_weight"]), float(lw["shape_reg_loss_weight"])
    s.w_collision, s.w_finger_reg = float(lw["collision_loss_weight"]), float(lw["finger_reg_loss_weight"])
    filters = stage.get("filter_loss", [])
    if len(filters) == 0:
        raise AssertionError("filter_loss must not be empty")       # opt_utils.py:118
    if len(filters) > 4:
        raise IhmrError("at most 4 filter criteria are supported")
    s.n_filters = len(filters)
    for i, (name, crit) in enumerate(filters):
        assert crit[0] in "+-"                                          # opt_utils.py:105
        if name not in LOSS_IDS:
            raise IhmrError(f"criterion {name!r} is not usable for filtering (opt_utils.py:57-67)")
        s.filter_loss[i] = LOSS_IDS[name]
        s.filter_percent[i] = float(crit)
    if stage["select_loss"] not in LOSS_IDS:
        raise IhmrError(f"criterion {stage['select_loss']!r} is not usable for selection")
    s.select_loss = LOSS_IDS[stage["select_loss"]]
    s.flags = STAGE_GENERIC_KERNELS if stage.get("generic_kernels") else 0
    return s
