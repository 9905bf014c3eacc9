"""ctypes binding of librsrx.so (include/rsrx.h).  There is no CPU fallback: if
the CUDA library is missing or does not load, importing the product API raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from .model import EnvCfg, ModelBlob

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librsrx.so")
CSRC = os.path.join(_HERE, "csrc")
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17"]

INFO_STRIDE = 20
INFO = dict(TARGET=0, TARGET2=3, NEWPOS=6, SITE=8, OBJ=11, LAST_ACTION=14, XITA=15, TARGET_W=16, STEPS=17,
            TRUNCATION=18)
STATUS_NONFINITE, STATUS_CONTACT_OVERFLOW, STATUS_SOLVER_CAP, STATUS_CONTACT_REDO = 1, 2, 4, 8
OBS_STRIDE = 24
METRICS_STRIDE = 8


class Layout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "data_stride", "qpos", "qvel", "ctrl", "qacc_warmstart", "time", "xpos", "xquat", "site_xpos", "geom_xpos",
        "obs_size", "obs_stride", "info_stride", "metrics_stride", "nq", "nv", "nu", "nbody", "nsite", "ngeom")]


class StateC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("data", "first_data", "obs", "first_obs", "reward", "done", "info",
                                          "metrics", "status")]


class PerEnvC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("geom_friction", "body_mass", "dof_damping", "dof_frictionloss")]


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc-compile csrc/rsrx_api.cu -> librsrx.so (sm_100a, in-tree)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs += [os.path.join(_HERE, "..", "include", f) for f in ("rsrx.h", "rsrx_model.h")]
    if (not force) and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = ["nvcc", *NVCC_FLAGS, "--threads", "3", "-o", LIB_PATH, os.path.join(CSRC, "rsrx_api.cu"),
           os.path.join(CSRC, "rsrx_redo.cu"), os.path.join(CSRC, "rsrx_mid.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return LIB_PATH


_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension is the only implementation of the Airbot stepper "
            "(no CPU fallback). Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    L = C.CDLL(LIB_PATH)
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    L.rsrx_last_error.restype = C.c_char_p
    L.rsrx_version.restype = C.c_char_p
    L.rsrx_model_blob_size.restype = C.c_size_t
    L.rsrx_env_cfg_size.restype = C.c_size_t
    L.rsrx_model_create.argtypes = [vp, C.c_size_t, C.POINTER(EnvCfg), C.POINTER(vp)]
    L.rsrx_model_destroy.argtypes = [vp]
    L.rsrx_model_destroy.restype = None
    L.rsrx_model_layout.argtypes = [vp, C.POINTER(Layout)]
    L.rsrx_env_reset.argtypes = [vp, i32, vp, vp, vp, C.POINTER(PerEnvC), StateC, vp]
    L.rsrx_env_step.argtypes = [vp, i32, StateC, vp, C.POINTER(PerEnvC), vp]
    L.rsrx_env_step_host.argtypes = [vp, i32, StateC, vp, vp, vp, vp, vp, C.POINTER(PerEnvC), vp]
    L.rsrx_physics_step.argtypes = [vp, i32, vp, i32, C.POINTER(PerEnvC), vp, vp]
    L.rsrx_physics_step_debug.argtypes = [vp, i32, vp, C.POINTER(PerEnvC), vp, vp]
    L.rsrx_debug_stride.restype = i32
    L.rsrx_max_contacts.restype = i32
    L.rsrx_kde.argtypes = [vp, i32, i32, vp, i32, f32, vp, vp]
    L.rsrx_rsr_loss.argtypes = [vp, i32, i32, vp, i32, vp, i32, vp, f32, f32, f32, vp, vp, vp, vp]
    L.rsrx_debug_narrowphase.argtypes = [vp, i32, i32, vp, vp]
    L.rsrx_act_bias_backward_workspace.argtypes = [i32, i32]
    L.rsrx_act_bias_backward_workspace.restype = C.c_size_t
    L.rsrx_act_bias_backward.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
    L.rsrx_tanh_normal_act.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]
    L.rsrx_gather_rows.argtypes = [vp, vp, vp, i32, vp, i32, vp]
    L.rsrx_ppo_head.argtypes = [vp] * 9 + [i32] * 3 + [f32] * 5 + [i32] + [vp] * 5
    L.rsrx_linear_forward.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp, i32, vp]
    L.rsrx_linear_dgrad.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, i32, vp, vp, i32, vp]
    L.rsrx_linear_wgrad.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, i32, vp]
    L.rsrx_reduce_partials.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    L.rsrx_value_head_backward.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp]
    L.rsrx_rsr_policy_term.argtypes = [vp, i32, vp, i32, vp, f32, f32, f32, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    L.rsrx_rsr_logit_grad.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]
    L.rsrx_ppo_prep.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, i32, vp, i32, vp]
    L.rsrx_small_mlp_forward.argtypes = [vp, vp, vp, i32, i32, vp, i32, i32, vp, vp, i32, vp, vp, vp]
    L.rsrx_small_mlp_backward.argtypes = [vp, vp, vp, i32, i32, vp, i32, i32, vp, vp, i32, vp, vp]
    L.rsrx_small_mlp_backward_ctas.argtypes = [i32]
    L.rsrx_small_mlp_backward_ctas.restype = i32
    L.rsrx_adam_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, f32, f32, f32, f32, f32, vp, vp]
    if L.rsrx_model_blob_size() != C.sizeof(ModelBlob) or L.rsrx_env_cfg_size() != C.sizeof(EnvCfg):
        raise RuntimeError("librsrx.so and rsr_mjx_b200/model.py disagree on the blob layout; rebuild the library")
    _LIB = L
    return L


def check(rc: int, what: str = "rsrx"):
    if rc != 0:
        raise RuntimeError(f"{what}: {lib().rsrx_last_error().decode()}")


EXPORTS = ("rsrx_model_create", "rsrx_model_destroy", "rsrx_model_layout", "rsrx_model_blob_size", "rsrx_env_cfg_size",
           "rsrx_env_reset", "rsrx_env_step", "rsrx_env_step_host", "rsrx_physics_step", "rsrx_debug_stride", "rsrx_max_contacts", "rsrx_physics_step_debug",
           "rsrx_rsr_loss", "rsrx_kde", "rsrx_ppo_head", "rsrx_debug_narrowphase", "rsrx_act_bias_backward", "rsrx_act_bias_backward_workspace", "rsrx_gather_rows", "rsrx_tanh_normal_act", "rsrx_last_error", "rsrx_version",
           "rsrx_linear_forward", "rsrx_linear_dgrad", "rsrx_linear_wgrad", "rsrx_reduce_partials", "rsrx_value_head_backward", "rsrx_adam_step",
           "rsrx_small_mlp_forward", "rsrx_small_mlp_backward", "rsrx_small_mlp_backward_ctas", "rsrx_ppo_prep", "rsrx_rsr_policy_term",
           "rsrx_rsr_logit_grad")
