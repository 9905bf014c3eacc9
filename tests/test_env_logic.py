"""Env pre/post-physics logic of the oracle against an independent NumPy
transcription of the reference's Python (test/airbot.py:165-268), run on the
oracle's own physics outputs."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from rsr_mjx_b200 import airbot_spec as A, prng
from rsr_mjx_b200.model import pack_model


def _np_sf_step(m, ids, cfgp, pre, post, info, action):
    """returns (ctrl_sent, reward, done, obs, new_cube_pos, last_action) as the reference computes them"""
    jp = np
    action_scale = jp.array([0.02, 0.02, 0.02, 0.0, 0.0])
    act = pre["ctrl"] + action_scale * action
    act[3] = -(1.57 + pre["qpos"][ids["joint_id"][1]] + pre["qpos"][ids["joint_id"]][2])
    cube = pre["xpos"][ids["cube_id"]]
    target = info["target_pos"]
    ang = jp.arctan2(target[1] - cube[1], target[0] - cube[0] + 0.00001)
    d0 = jp.linalg.norm(target - cube)
    act[4] = info["last_action"] if d0 < 0.03 else -ang + act[0] + 1.5708
    last_action = act[4]
    act = jp.clip(act, m.act_ctrlrange[:, 0], m.act_ctrlrange[:, 1])
    cube1, site = post["xpos"][ids["cube_id"]], post["site_xpos"][ids["site_id"]]
    dis = jp.linalg.norm(target - cube1)
    dis = 0.0 if dis < 0.003 else dis
    push = 1 / (1 + 3 * dis) * cfgp["push_reward_weight"]
    task_complete = 5.0 if dis < 0.003 else 0.0
    old_new = info["new_cube_pos"].copy()
    site_z_reward = 1.0 if site[2] < 0.82 else 0.0
    dx, dy = target[0] - cube1[0], target[1] - cube1[1]
    ang = jp.arctan2(dy, dx + 0.00001)
    distance = jp.sqrt(dx ** 2 + dy ** 2) + 0.04
    new = jp.array([dx - distance * jp.cos(ang) + cube1[0], dy - distance * jp.sin(ang) + cube1[1]])
    sd = jp.linalg.norm(site[:2] - old_new)
    sd = 0 if sd < 0.042 else sd - 0.042
    s2c = (1 - jp.tanh(5 * sd)) * cfgp["siet_to_box_reward_weight"]
    if dis < 0.005:
        s2c = cfgp["siet_to_box_reward_weight"]
    done = 0.0
    if site[2] < cfgp["endpoint_min_z_pos"] or site[0] > 1.0 or site[0] < -0.6 or site[1] > 0.3 or site[1] < -0.3 or cube1[2] < 0.6:
        done = 1.0
    health = cfgp["healthy_reward"] * abs(done - 1.0)
    reward = float(jp.clip(push + s2c + health + task_complete + site_z_reward, -1e2, 1e2))
    done = 1.0 if dis < 0.003 else 0.0
    obs = jp.concatenate([post["qpos"][ids["joint_id"]], site, target, cube1, new, target - cube1, cube1 - site])
    return act, reward, done, obs, new, last_action


def test_sf_env_logic_matches_numpy_transcription(oracle_built):
    kind = "sf"
    m = A.load_model(kind)
    ids = A.env_ids(m, kind)
    blob, cfg = pack_model(m), A.make_env_cfg(m, kind, episode_length=1000)
    qpos, qvel, ctrl = A.sample_reset(m, kind, prng.split(prng.PRNGKey(21), 2))
    s = O.env_reset(blob, cfg, qpos[0], qvel[0], ctrl[0])
    rng = np.random.default_rng(3)

    def view(d):
        return dict(qpos=np.array(d.qpos)[:m.nq], ctrl=np.array(d.ctrl)[:m.nu],
                    xpos=np.ctypeslib.as_array(d.xpos)[:m.nbody].copy(),
                    site_xpos=np.ctypeslib.as_array(d.site_xpos)[:m.nsite].copy())

    for t in range(25):
        a = rng.uniform(-1, 1, 5)
        pre = view(s.d)
        info = dict(target_pos=np.array(s.target_pos), new_cube_pos=np.array(s.new_pos), last_action=s.last_action)
        O.env_step(blob, cfg, s, a)
        post = view(s.d)
        act, reward, done, obs, new, last = _np_sf_step(m, ids, A.DEFAULTS[kind], pre, post, info, a)
        np.testing.assert_allclose(post["ctrl"], act, atol=1e-12)
        assert s.reward == pytest.approx(reward, abs=1e-10)
        assert s.done == done
        np.testing.assert_allclose(np.array(s.obs)[:23], obs, atol=1e-12)
        np.testing.assert_allclose(np.array(s.new_pos), new, atol=1e-12)
        assert s.last_action == pytest.approx(last, abs=1e-12)
        assert s.steps == t + 1


def test_reset_sampling_follows_reference_ranges():
    for kind in ("sf", "cube", "T"):
        m = A.load_model(kind)
        ids = A.env_ids(m, kind)
        p = A.DEFAULTS[kind]
        qpos, qvel, ctrl = A.sample_reset(m, kind, prng.split(prng.PRNGKey(9), 256))
        assert qpos.dtype == np.float32 and np.abs(qvel).max() <= 0.01
        if kind != "T":
            adr = ids["_box_qposadr"]
            assert (qpos[:, adr] >= p["cube_min_x"] - 1e-7).all() and (qpos[:, adr] <= p["cube_max_x"]).all()
            assert (qpos[:, adr + 2] == np.float32(0.82)).all()
            assert (qpos[:, ids["finger_id"][0]] == np.float32(-0.033)).all()
            tadr = ids["_site_qposadr"]
            assert (qpos[:, tadr + 1] >= p["target_min_y"] - 1e-7).all() and (qpos[:, tadr + 1] <= p["target_max_y"]).all()
        # joint4 starts near 1.5718 (held by its limit afterwards)
        assert np.abs(qpos[:, 3] - 1.5718).max() <= 0.01 + 1e-6


def test_domain_randomize_ranges_and_untouched_leaves():
    from rsr_mjx_b200 import domain_randomize as DR
    m = A.load_model("cube")
    arr = DR.domain_randomize_arrays(m, prng.split(prng.PRNGKey(0), 128))
    gf, bm, dd, fl = arr["geom_friction"], arr["body_mass"], arr["dof_damping"], arr["dof_frictionloss"]
    table, cube = m.geom("table-b"), m.geom("geom_for_push")
    r = gf[:, table, 0] / m.geom_friction[table, 0]
    assert r.min() >= 0.68 - 1e-6 and r.max() <= 1.32 + 1e-6 and r.std() > 0.1
    r = bm[:, 13] / m.body_mass[13]
    assert r.min() >= 0.84 - 1e-6 and r.max() <= 1.16 + 1e-6
    np.testing.assert_allclose(dd[:, 8:], 0) and np.testing.assert_allclose(fl[:, 8:], 0)
    r = dd[:, :8] / m.dof_damping[:8]
    assert r.min() >= 0.92 - 1e-6 and r.max() <= 1.08 + 1e-6
    # geoms other than table / cube / fingers keep the nominal friction
    fingers = [g for g in range(m.ngeom) if m.geom_bodyid[g] in (m.body("left"), m.body("right"))]
    others = [g for g in range(m.ngeom) if g not in fingers + [table, cube]]
    np.testing.assert_array_equal(gf[:, others], np.tile(m.geom_friction[others].astype(np.float32), (128, 1, 1)))
