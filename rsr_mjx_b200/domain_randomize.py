"""Domain randomisation of the Airbot cube env, following the reference's
ppo_train/airbot_training/domain_randomize.py:26-91: per env, six scalars drawn
by successive `rng, key = split(rng)` / `uniform(key, minval, maxval)`:
table friction scale, cube friction scale, cube mass scale, finger friction
scale, damping scale, frictionloss scale (dofs 0:8).  Nominal `invweight0` /
`meaninertia` / `body_inertia` are NOT recomputed (as in the reference, which
only swaps the four arrays via `tree_replace`)."""
from __future__ import annotations

import numpy as np

from . import prng
from .mjcf import Model

_FRICTION_TABLE_CUBE = (0.68, 1.32)
_MASS_CUBE = (0.84, 1.16)
_FRICTION_FINGER = (0.76, 1.24)
_JOINT_SCALE = (0.92, 1.08)
_ARM_DOF_SLICE = slice(0, 8)


def domain_randomize_arrays(m: Model, rng: np.ndarray) -> dict:
    """rng: keys [N,2] -> dict of float32 arrays with a leading N axis."""
    rng = np.asarray(rng, np.uint32).reshape(-1, 2)
    N = rng.shape[0]
    scales = []
    for lo, hi in (_FRICTION_TABLE_CUBE, _FRICTION_TABLE_CUBE, _MASS_CUBE, _FRICTION_FINGER, _JOINT_SCALE, _JOINT_SCALE):
        ks = prng.split(rng, 2)
        rng, key = ks[:, 0], ks[:, 1]
        scales.append(prng.uniform(key, (), np.float32(lo), np.float32(hi)).reshape(N))
    table_s, cube_s, mass_s, finger_s, damp_s, floss_s = scales
    f32 = np.float32
    table, cube_g, cube_b = m.geom("table-b"), m.geom("geom_for_push"), m.body("cube_for_push")
    fingers = [g for g in range(m.ngeom) if m.geom_bodyid[g] in (m.body("left"), m.body("right"))]
    gf = np.tile(m.geom_friction.astype(f32)[None], (N, 1, 1))
    gf[:, table] *= table_s[:, None]
    gf[:, cube_g] *= cube_s[:, None]
    gf[:, fingers] *= finger_s[:, None, None]
    bm = np.tile(m.body_mass.astype(f32)[None], (N, 1))
    bm[:, cube_b] *= mass_s
    dd = np.tile(m.dof_damping.astype(f32)[None], (N, 1))
    dd[:, _ARM_DOF_SLICE] *= damp_s[:, None]
    fl = np.tile(m.dof_frictionloss.astype(f32)[None], (N, 1))
    fl[:, _ARM_DOF_SLICE] *= floss_s[:, None]
    return dict(geom_friction=gf, body_mass=bm, dof_damping=dd, dof_frictionloss=fl)


def domain_randomize(sys, rng):
    """Reference-shaped entry point: (sys, rng[N,2]) -> (sys_batched, in_axes)."""
    arrays = domain_randomize_arrays(sys.mj_model, rng)
    in_axes = {k: 0 for k in arrays}
    return sys.tree_replace(arrays), in_axes
