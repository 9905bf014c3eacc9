"""ctypes mirror of include/rsrx_model.h and the Model -> blob packer.

The blob is the only form in which a compiled model crosses the C-ABI
(`rsrx_model_create`).  Field order/capacities must match the header exactly;
`rsrx_model_blob_size()` / `rsrx_env_cfg_size()` are checked at load time.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .mjcf import Model

RSRX_MAGIC = 0x52535258
RSRX_VERSION = 3
MAXBODY, MAXJNT, MAXQ, MAXV, MAXU, MAXGEOM, MAXSITE, MAXPAIR, MAXEQ = 16, 12, 24, 20, 8, 32, 4, 64, 2
ENV_SF, ENV_CUBE, ENV_T = 0, 1, 2

_i = C.c_int32
_d = C.c_double


def _I(n):
    return _i * n


def _D(*shape):
    t = _d
    for s in reversed(shape):
        t = t * s
    return t


class ModelBlob(C.Structure):
    _fields_ = [
        ("magic", _i), ("version", _i),
        ("nbody", _i), ("njnt", _i), ("nq", _i), ("nv", _i), ("nu", _i), ("ngeom", _i), ("nsite", _i),
        ("npair", _i), ("neq", _i),
        ("iterations", _i), ("ls_iterations", _i), ("pad0", _i),
        ("timestep", _d), ("gravity", _D(3)), ("tolerance", _d), ("ls_tolerance", _d), ("impratio", _d),
        ("meaninertia", _d),
        ("body_parentid", _I(MAXBODY)), ("body_rootid", _I(MAXBODY)), ("body_weldid", _I(MAXBODY)),
        ("body_jntadr", _I(MAXBODY)), ("body_jntnum", _I(MAXBODY)), ("body_dofadr", _I(MAXBODY)),
        ("body_dofnum", _I(MAXBODY)), ("body_depth", _I(MAXBODY)),
        ("body_pos", _D(MAXBODY, 3)), ("body_quat", _D(MAXBODY, 4)), ("body_ipos", _D(MAXBODY, 3)),
        ("body_iquat", _D(MAXBODY, 4)), ("body_mass", _D(MAXBODY)), ("body_inertia", _D(MAXBODY, 3)),
        ("body_invweight0", _D(MAXBODY, 2)),
        ("jnt_type", _I(MAXJNT)), ("jnt_qposadr", _I(MAXJNT)), ("jnt_dofadr", _I(MAXJNT)),
        ("jnt_bodyid", _I(MAXJNT)), ("jnt_limited", _I(MAXJNT)), ("jnt_actfrclimited", _I(MAXJNT)),
        ("jnt_pos", _D(MAXJNT, 3)), ("jnt_axis", _D(MAXJNT, 3)), ("jnt_range", _D(MAXJNT, 2)),
        ("jnt_actfrcrange", _D(MAXJNT, 2)), ("jnt_solref", _D(MAXJNT, 2)), ("jnt_solimp", _D(MAXJNT, 5)),
        ("jnt_margin", _D(MAXJNT)),
        ("qpos0", _D(MAXQ)),
        ("dof_bodyid", _I(MAXV)), ("dof_jntid", _I(MAXV)), ("dof_parentid", _I(MAXV)),
        ("dof_damping", _D(MAXV)), ("dof_frictionloss", _D(MAXV)), ("dof_armature", _D(MAXV)),
        ("dof_invweight0", _D(MAXV)), ("dof_solref", _D(MAXV, 2)), ("dof_solimp", _D(MAXV, 5)),
        ("geom_type", _I(MAXGEOM)), ("geom_bodyid", _I(MAXGEOM)), ("geom_contype", _I(MAXGEOM)),
        ("geom_conaffinity", _I(MAXGEOM)), ("geom_condim", _I(MAXGEOM)), ("geom_priority", _I(MAXGEOM)),
        ("geom_pos", _D(MAXGEOM, 3)), ("geom_quat", _D(MAXGEOM, 4)), ("geom_size", _D(MAXGEOM, 3)),
        ("geom_friction", _D(MAXGEOM, 3)), ("geom_solref", _D(MAXGEOM, 2)), ("geom_solimp", _D(MAXGEOM, 5)),
        ("geom_solmix", _D(MAXGEOM)), ("geom_margin", _D(MAXGEOM)), ("geom_gap", _D(MAXGEOM)),
        ("site_bodyid", _I(MAXSITE)), ("site_pos", _D(MAXSITE, 3)), ("site_quat", _D(MAXSITE, 4)),
        ("pair_geom1", _I(MAXPAIR)), ("pair_geom2", _I(MAXPAIR)),
        ("act_trnid", _I(MAXU)), ("act_ctrllimited", _I(MAXU)), ("act_forcelimited", _I(MAXU)),
        ("act_gear", _D(MAXU)), ("act_gainprm", _D(MAXU, 3)), ("act_biasprm", _D(MAXU, 3)),
        ("act_ctrlrange", _D(MAXU, 2)), ("act_forcerange", _D(MAXU, 2)),
        ("eq_obj1id", _I(MAXEQ)), ("eq_obj2id", _I(MAXEQ)),
        ("eq_data", _D(MAXEQ, 5)), ("eq_solref", _D(MAXEQ, 2)), ("eq_solimp", _D(MAXEQ, 5)),
    ]


class EnvCfg(C.Structure):
    _fields_ = [
        ("env_kind", _i), ("episode_length", _i), ("action_repeat", _i), ("n_frames", _i),
        ("cube_body", _i), ("target_body", _i), ("site_endpoint", _i), ("site_tail", _i),
        ("site_target_tail", _i), ("geom_base", _i), ("geom_vertical", _i),
        ("geom_target_base", _i), ("geom_target_vertical", _i),
        ("joint_qadr", _I(6)), ("pad0", _i),
        ("action_scale", _D(MAXU)),
        ("push_reward_weight", _d), ("siet_to_box_reward_weight", _d), ("healthy_reward", _d),
        ("endpoint_min_z_pos", _d),
    ]


_SCALARS = ("nbody", "njnt", "nq", "nv", "nu", "ngeom", "nsite", "npair", "neq", "iterations",
            "ls_iterations", "timestep", "tolerance", "ls_tolerance", "impratio", "meaninertia")


def pack_model(m: Model) -> ModelBlob:
    """Model -> fixed-capacity blob (raises if the model exceeds a capacity)."""
    caps = dict(nbody=MAXBODY, njnt=MAXJNT, nq=MAXQ, nv=MAXV, nu=MAXU, ngeom=MAXGEOM, nsite=MAXSITE,
                npair=MAXPAIR, neq=MAXEQ)
    for k, cap in caps.items():
        if getattr(m, k) > cap:
            raise ValueError(f"model {k}={getattr(m, k)} exceeds blob capacity {cap}")
    b = ModelBlob()
    b.magic, b.version = RSRX_MAGIC, RSRX_VERSION
    for k in _SCALARS:
        setattr(b, k, getattr(m, k))
    b.gravity[:] = list(m.gravity)
    skip = set(_SCALARS) | {"magic", "version", "pad0", "gravity"}
    for name, ctype in ModelBlob._fields_:
        if name in skip:
            continue
        src = np.asarray(m.arrays[name])
        dst = np.ctypeslib.as_array(getattr(b, name))
        dst[...] = 0
        if src.size:
            dst[tuple(slice(0, s) for s in src.shape)] = src
    return b


def blob_bytes(b: ModelBlob) -> bytes:
    return bytes(b)
