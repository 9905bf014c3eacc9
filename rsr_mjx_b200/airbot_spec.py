"""Host-side constants and reset sampling of the three reference Airbot envs.

Mirrors the constructor defaults, cached ids and `reset` sampling of
  * test/airbot.py                              (kind 'sf',   23-d obs)
  * ppo_train/airbot_training/cube_env.py       (kind 'cube', 23-d obs)
  * ppo_train/airbot_training/T_shape_env.py    (kind 'T',    16-d obs)
Everything here is pure NumPy host logic shared by the product env class and by
the tests that drive the CPU oracle.
"""
from __future__ import annotations

import os

import numpy as np

from . import prng
from .mjcf import Model, compile_mjcf
from .model import ENV_CUBE, ENV_SF, ENV_T, EnvCfg

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

KINDS = {"sf": ENV_SF, "cube": ENV_CUBE, "T": ENV_T}
MODEL_FILES = {"sf": "airbot_sf.xml", "cube": "airbot_cube.xml", "T": "airbot_T.xml"}
OBS_SIZE = {"sf": 23, "cube": 23, "T": 16}
N_INFO = 16  # floats of per-env info kept on device (see csrc/rsrx_layout.h)
METRIC_KEYS = {
    "sf": ("push_reward", "ctrl_cost", "siet_to_box_reward"),
    "cube": ("push_reward", "ctrl_cost", "siet_to_box_reward"),
    "T": ("push_reward", "siet2cube_reward", "health_reward", "task_complete_reward", "site_z_reward"),
}

# constructor defaults: test/airbot.py:10-41, cube_env.py:9-35, T_shape_env.py:11-37
DEFAULTS = {
    "sf": dict(push_reward_weight=6, siet_to_box_reward_weight=3.0, healthy_reward=1.0, endpoint_min_z_pos=0.8,
               noise_scale=1e-2, decimation=4,
               cube_min_x=0.28, cube_max_x=0.29, cube_min_y=-0.005, cube_max_y=0.005,
               target_min_x=0.5, target_max_x=0.51, target_min_y=-0.005, target_max_y=0.005),
    "cube": dict(push_reward_weight=6, siet_to_box_reward_weight=3.0, healthy_reward=1.0, endpoint_min_z_pos=0.778,
                 noise_scale=1e-2, decimation=4,
                 cube_min_x=0.29, cube_max_x=0.34, cube_min_y=-0.04, cube_max_y=0.01,
                 target_min_x=0.4364427, target_max_x=0.4864427, target_min_y=0.07352592, target_max_y=0.12352592),
    "T": dict(push_reward_weight=10.0, siet_to_box_reward_weight=3.0, healthy_reward=1.0, endpoint_min_z_pos=0.78,
              noise_scale=1e-2, decimation=4,
              cube_min_x=0.29, cube_max_x=0.34, cube_min_y=-0.04, cube_max_y=0.01,
              target_min_x=0.4364427, target_max_x=0.4864427, target_min_y=0.07352592, target_max_y=0.12352592),
}
ACTION_SCALE = (0.02, 0.02, 0.02, 0.0, 0.0)
_ARM_JOINTS = ("joint1", "joint2", "joint3", "joint4", "joint5", "joint6")
# reset offsets: test/airbot.py:110,115 ; T_shape_env.py:105,109
_ARM_OFFSET = {
    "sf": (0, -0.5422302, 0.45173569, 1.5718, -1.4794435, 1.1731174),
    "cube": (0, -0.5422302, 0.45173569, 1.5718, -1.4794435, 1.1731174),
    "T": (0, -0.57303354, 0.381795, 1.5718, -1.3787, 1.1731174),
}
_CTRL0 = {
    "sf": (0, -0.73151061, 0.455936904, -1.4794435, 1.1731174),
    "cube": (0, -0.73151061, 0.455936904, -1.4794435, 1.1731174),
    "T": (0, -0.57303354, 0.381795, -1.3787, 1.1731174),
}


def default_model_path(kind: str) -> str:
    return os.path.join(_ASSETS, MODEL_FILES[kind])


def load_model(kind: str, path: str | None = None) -> Model:
    return compile_mjcf(path or default_model_path(kind))


def env_ids(m: Model, kind: str) -> dict:
    """ids cached by AirbotPlayBase.__init__ (test/airbot.py:72-91, T_shape_env.py:65-92)."""
    ids = {}
    ids["joint_id"] = np.array([m.jnt_qposadr[m.joint(j)] for j in _ARM_JOINTS], np.int32)
    ids["finger_id"] = np.array([m.jnt_qposadr[m.joint("endleft")]], np.int32)
    ids["site_id"] = m.site("endpoint")
    if kind == "T":
        ids["T_id"] = ids["cube_id"] = m.body("T_block")
        ids["target_body_id"] = ids["target_pos_id"] = m.body("T_target")
        ids["_box_qposadr"] = int(m.jnt_qposadr[m.body_jntadr[ids["T_id"]]])
        ids["T_tail_id"] = m.site("T_tail")
        ids["T_target_tail_id"] = m.site("T_target_tail")
        ids["T_base_geom_id"] = m.geom("base_block")
        ids["T_vertical_geom_id"] = m.geom("vertical_block")
        ids["T_target_base_geom_id"] = m.geom("base_target")
        ids["T_target_vertical_geom_id"] = m.geom("vertical_target")
    else:
        ids["cube_id"] = m.body("cube_for_push")
        ids["target_pos_id"] = m.body("target_pos")
        ids["_box_qposadr"] = int(m.jnt_qposadr[m.body_jntadr[ids["cube_id"]]])
        ids["_site_qposadr"] = int(m.jnt_qposadr[m.body_jntadr[ids["target_pos_id"]]])
        ids["cube_geom_id"] = m.geom("geom_for_push")
    ids["fixed_gripper_geom_id"] = m.geom("fixed_gripper")
    ids["table_id"] = m.geom("table-b")
    return ids


def make_env_cfg(m: Model, kind: str, params: dict | None = None, episode_length: int = 1000,
                 action_repeat: int = 1) -> EnvCfg:
    p = dict(DEFAULTS[kind])
    p.update(params or {})
    ids = env_ids(m, kind)
    c = EnvCfg()
    c.env_kind = KINDS[kind]
    c.episode_length = int(episode_length)
    c.action_repeat = int(action_repeat)
    c.n_frames = int(p.get("n_frames", p["decimation"]))
    c.cube_body = ids["cube_id"]
    c.target_body = ids["target_pos_id"]
    c.site_endpoint = ids["site_id"]
    c.site_tail = ids.get("T_tail_id", 0)
    c.site_target_tail = ids.get("T_target_tail_id", 0)
    c.geom_base = ids.get("T_base_geom_id", 0)
    c.geom_vertical = ids.get("T_vertical_geom_id", 0)
    c.geom_target_base = ids.get("T_target_base_geom_id", 0)
    c.geom_target_vertical = ids.get("T_target_vertical_geom_id", 0)
    c.joint_qadr[:] = [int(x) for x in ids["joint_id"]]
    for i, a in enumerate(ACTION_SCALE):
        c.action_scale[i] = a
    c.push_reward_weight = float(p["push_reward_weight"])
    c.siet_to_box_reward_weight = float(p["siet_to_box_reward_weight"])
    c.healthy_reward = float(p["healthy_reward"])
    c.endpoint_min_z_pos = float(p["endpoint_min_z_pos"])
    return c


def sample_reset(m: Model, kind: str, keys: np.ndarray, params: dict | None = None):
    """The sampling half of `reset` for a batch of jax-style keys [N,2] (uint32).

    Returns float32 (qpos[N,nq], qvel[N,nv], ctrl[N,nu]) exactly as
    test/airbot.py:104-133 / T_shape_env.py:100-110 build them (float32 math).
    """
    p = dict(DEFAULTS[kind])
    p.update(params or {})
    ids = env_ids(m, kind)
    keys = np.asarray(keys, np.uint32).reshape(-1, 2)
    ks = prng.split(keys, 5)  # rng, rng1, rng2, rng3, rng4
    rng, rng1, rng2, rng3, rng4 = (ks[:, i] for i in range(5))
    lo, hi = np.float32(-p["noise_scale"]), np.float32(p["noise_scale"])
    f32 = np.float32
    qpos = m.qpos0.astype(f32)[None] + prng.uniform(rng1, (m.nq,), lo, hi)
    qpos[:, ids["joint_id"]] += np.array(_ARM_OFFSET[kind], f32)
    if kind != "T":
        qpos[:, ids["finger_id"]] = f32(-0.033)
    qvel = prng.uniform(rng2, (m.nv,), lo, hi)
    ctrl = np.array(_CTRL0[kind], f32)[None] + prng.uniform(rng3, (m.nu,), lo, hi)
    if kind != "T":
        tmin = np.array([p["target_min_x"], p["target_min_y"], 0.82], f32)
        tmax = np.array([p["target_max_x"], p["target_max_y"], 0.82], f32)
        cmin = np.array([p["cube_min_x"], p["cube_min_y"], 0.82], f32)
        cmax = np.array([p["cube_max_x"], p["cube_max_y"], 0.82], f32)
        target = prng.uniform(rng4, (3,), tmin, tmax)
        cube = prng.uniform(rng, (3,), cmin, cmax)
        qpos[:, ids["_box_qposadr"]:ids["_box_qposadr"] + 3] = cube
        qpos[:, ids["_site_qposadr"]:ids["_site_qposadr"] + 3] = target
    return qpos.astype(f32), qvel.astype(f32), ctrl.astype(f32)
