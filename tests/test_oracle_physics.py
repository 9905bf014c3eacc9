"""Self-consistency of the CPU oracle (no MJX is available to compare with —
SURVEY.md §8c): independent formulas, conservation laws, statics, KKT."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import oracle as O
from rsr_mjx_b200 import airbot_spec as A, mjcf, prng
from rsr_mjx_b200.model import pack_model

BOX_ON_PLANE = """
<mujoco>
  <compiler angle="radian"/>
  <option timestep="0.002" integrator="implicitfast" gravity="{gx} 0 {gz}"/>
  <default><geom condim="4" friction="{mu} 0.005 0.0001"/></default>
  <worldbody>
    <geom type="plane" size="5 5 0.1"/>
    <body name="b" pos="0 0 0.1">
      <freejoint/>
      <geom name="g" type="box" size="0.1 0.1 0.1"/>
    </body>
  </worldbody>
</mujoco>
"""
PENDULUM = """
<mujoco>
  <compiler angle="radian"/>
  <option integrator="implicitfast" timestep="0.0005" gravity="0 0 {g}"/>
  <default><geom contype="0" conaffinity="0"/></default>
  <worldbody>
    <body name="l1">
      <joint name="j1" type="hinge" axis="0 1 0"/>
      <inertial pos="0.1 0 -0.5" mass="2" diaginertia="0.1 0.2 0.3"/>
      <body name="l2" pos="0 0 -1" quat="0.9689124 0.2474040 0 0">
        <joint name="j2" type="hinge" axis="0 0 1"/>
        <inertial pos="0 0.2 -0.25" mass="1" diaginertia="0.01 0.02 0.03"/>
        <body name="l3" pos="0 0.3 0">
          <joint name="j3" type="slide" axis="1 0 0"/>
          <inertial pos="0 0 0.1" mass="0.5" diaginertia="0.01 0.01 0.01"/>
        </body>
      </body>
    </body>
  </worldbody>
</mujoco>
"""


@pytest.fixture(scope="module", autouse=True)
def _build(oracle_built):
    return oracle_built


def _random_state(m, seed):
    rng = np.random.default_rng(seed)
    q = m.qpos0 + rng.uniform(-0.3, 0.3, m.nq)
    for j in range(m.njnt):
        if m.jnt_type[j] == 0:
            a = m.jnt_qposadr[j]
            q[a + 3:a + 7] /= np.linalg.norm(q[a + 3:a + 7])
    return q, rng.uniform(-1, 1, m.nv)


@pytest.mark.parametrize("kind", ["sf", "T"])
def test_mass_matrix_equals_sum_JtIJ(kind):
    m = A.load_model(kind)
    blob = pack_model(m)
    for seed in range(3):
        q, v = _random_state(m, seed)
        ins = O.inspect(blob, O.make_data(blob, q, v))
        Mref, _ = mjcf.mass_matrix(m, q)
        np.testing.assert_allclose(ins["M"], Mref, rtol=1e-9, atol=1e-12)
        assert np.all(np.linalg.eigvalsh(ins["M"]) > 0)


def test_gravity_bias_equals_minus_Jt_mg():
    m = A.load_model("sf")
    blob = pack_model(m)
    q, _ = _random_state(m, 5)
    d = O.forward(blob, O.make_data(blob, q, np.zeros(m.nv)))
    kin = mjcf.kinematics(m, q)
    G = np.zeros(m.nv)
    for b in range(1, m.nbody):
        jp, _ = mjcf.jacobian(m, kin, kin["xipos"][b], b)
        G += jp.T @ (m.body_mass[b] * np.array([0, 0, 9.81]))
    np.testing.assert_allclose(np.array(d.qfrc_bias)[:m.nv], G, rtol=1e-9, atol=1e-10)


def test_energy_conserved_without_gravity_and_damping():
    m = mjcf.compile_mjcf(PENDULUM.format(g=0), from_string=True)
    blob = pack_model(m)
    q = np.array([0.4, -0.9, 0.05])
    v = np.array([1.5, -2.0, 0.3])
    d = O.make_data(blob, q, v)
    M0, _ = mjcf.mass_matrix(m, q)
    e0 = 0.5 * v @ M0 @ v
    O.step(blob, d, 2000)
    q1, v1 = np.array(d.qpos)[:3], np.array(d.qvel)[:3]
    M1, _ = mjcf.mass_matrix(m, q1)
    e1 = 0.5 * v1 @ M1 @ v1
    assert abs(e1 - e0) / e0 < 2e-3  # semi-implicit Euler, dt = 0.5 ms, 1 s of motion
    assert np.abs(q1 - q).max() > 0.5  # it really moved


def test_pendulum_period():
    m = mjcf.compile_mjcf("""
    <mujoco><compiler angle="radian"/><option integrator="implicitfast" timestep="0.0005"/>
    <worldbody><body><joint type="hinge" axis="0 1 0"/>
    <inertial pos="0 0 -1" mass="1" diaginertia="1e-6 1e-6 1e-6"/></body></worldbody></mujoco>""", from_string=True)
    blob = pack_model(m)
    d = O.make_data(blob, [0.05], [0.0])
    # small-angle period of a 1 m point pendulum: 2*pi*sqrt(1/9.81) = 2.006 s; a quarter period brings q to ~0
    nquarter = int(round(0.25 * 2 * math.pi * math.sqrt(1 / 9.81) / 0.0005))
    O.step(blob, d, nquarter)
    assert abs(d.qpos[0]) < 1e-3 and d.qvel[0] < 0


def test_box_rests_on_plane_with_weight_supported():
    m = mjcf.compile_mjcf(BOX_ON_PLANE.format(gx=0, gz=-9.81, mu=1.0), from_string=True)
    blob = pack_model(m)
    d = O.make_data(blob, m.qpos0, np.zeros(6))
    O.step(blob, d, 1500)
    assert abs(d.qpos[2] - 0.1) < 1e-3 and np.abs(np.array(d.qvel)[:6]).max() < 1e-4
    ins = O.inspect(blob, d)
    assert len(ins["contacts"]) == 4
    # pyramid rows all contain the normal direction with weight 1 -> normal force = sum of row forces
    fn = ins["force"].sum()
    mass = m.body_mass[1]
    assert fn == pytest.approx(mass * 9.81, rel=1e-3)
    # contact points: the four bottom corners, normal +z
    P = np.array(sorted([tuple(np.round(c["pos"][:2], 3)) for c in ins["contacts"]]))
    np.testing.assert_allclose(P, [[-0.1, -0.1], [-0.1, 0.1], [0.1, -0.1], [0.1, 0.1]], atol=2e-3)
    for c in ins["contacts"]:
        np.testing.assert_allclose(c["frame"][0], [0, 0, 1], atol=1e-9)


@pytest.mark.parametrize("theta,slides", [(25.0, False), (55.0, True)])
def test_friction_hold_and_slide(theta, slides):
    """gravity tilted by theta about y; mu = 1 along the pyramid axis x -> holds iff tan(theta) < 1"""
    th = math.radians(theta)
    m = mjcf.compile_mjcf(BOX_ON_PLANE.format(gx=9.81 * math.sin(th), gz=-9.81 * math.cos(th), mu=1.0), from_string=True)
    blob = pack_model(m)
    d = O.make_data(blob, m.qpos0, np.zeros(6))
    n = 500
    O.step(blob, d, n)
    t = n * 0.002
    if slides:
        # MuJoCo's soft pyramidal cone gives an effective friction <= mu (the +-t2 / torsion edges carry
        # normal load without opposing the slide), so the box accelerates at least as fast as Coulomb predicts
        a_coulomb, a_free = 9.81 * (math.sin(th) - 1.0 * math.cos(th)), 9.81 * math.sin(th)
        assert 0.95 * a_coulomb * t < d.qvel[0] < 0.75 * a_free * t
    else:
        assert abs(d.qvel[0]) < 5e-3 and abs(d.qpos[0]) < 5e-3  # soft-constraint creep only


def test_limits_and_equality_hold_on_airbot():
    m = A.load_model("sf")
    blob, cfg = pack_model(m), A.make_env_cfg(m, "sf", episode_length=10_000)
    qpos, qvel, ctrl = A.sample_reset(m, "sf", prng.split(prng.PRNGKey(3), 1))
    s = O.env_reset(blob, cfg, qpos[0], qvel[0], ctrl[0])
    for _ in range(60):
        O.env_step(blob, cfg, s, np.zeros(5))
    q = np.array(s.d.qpos)
    j4 = m.jnt_qposadr[m.joint("joint4")]
    assert 1.569 - 2e-3 < q[j4] < 1.571 + 2e-3
    r, l = m.jnt_qposadr[m.joint("endright")], m.jnt_qposadr[m.joint("endleft")]
    assert abs(q[l] + q[r]) < 1e-3
    assert -0.0331 - 1e-3 < q[l] < -0.0329 + 1e-3
    cube_z = s.d.xpos[cfg.cube_body][2]
    assert abs(cube_z - 0.82) < 1e-3  # resting on the table top (0.78) with half-size 0.04


@pytest.mark.parametrize("kind", ["sf", "T"])
def test_solver_kkt_residual_and_cone(kind):
    m = A.load_model(kind)
    blob, cfg = pack_model(m), A.make_env_cfg(m, kind)
    qpos, qvel, ctrl = A.sample_reset(m, kind, prng.split(prng.PRNGKey(11), 3))
    for e in range(3):
        s = O.env_reset(blob, cfg, qpos[e], qvel[e], ctrl[e])
        rng = np.random.default_rng(e)
        for _ in range(10):
            O.env_step(blob, cfg, s, rng.uniform(-1, 1, 5))
        d = s.d
        ins = O.inspect(blob, d)
        nv = m.nv
        qacc = np.array(d.qacc)[:nv]
        # stationarity: M qacc = qfrc_smooth + J^T f, with qfrc_smooth = M qacc_smooth
        lhs = ins["M"] @ (qacc - np.array(d.qacc_smooth)[:nv])
        rhs = ins["J"].T @ ins["force"]
        assert np.abs(lhs - rhs).max() <= 1e-6 * max(1.0, np.abs(rhs).max())
        assert d.solver_niter < m.iterations  # the iteration cap does not bind
        # pyramid edges only push: contact-row forces are >= 0
        ncontact_rows = 6 * len(ins["contacts"])
        if ncontact_rows:
            assert ins["force"][-ncontact_rows:].min() >= 0


def test_dense_mjx_work_pattern_gives_identical_result():
    m = A.load_model("sf")
    blob, cfg = pack_model(m), A.make_env_cfg(m, "sf")
    qpos, qvel, ctrl = A.sample_reset(m, "sf", prng.split(prng.PRNGKey(5), 1))
    a = O.env_reset(blob, cfg, qpos[0], qvel[0], ctrl[0])
    b = O.env_reset(blob, cfg, qpos[0], qvel[0], ctrl[0], dense=True)
    assert b.d.ncon == 4 * m.npair and a.d.ncon == a.d.ncon_active == b.d.ncon_active
    rng = np.random.default_rng(0)
    for _ in range(8):
        act = rng.uniform(-1, 1, 5)
        O.env_step(blob, cfg, a, act)
        O.env_step(blob, cfg, b, act, dense=True)
        np.testing.assert_allclose(np.array(a.d.qpos), np.array(b.d.qpos), rtol=0, atol=1e-10)
        np.testing.assert_allclose(np.array(a.d.qvel), np.array(b.d.qvel), rtol=0, atol=1e-8)
        assert a.reward == pytest.approx(b.reward, abs=1e-9)


def test_f32_oracle_tracks_f64_teacher_forced():
    """the float32 noise floor the CUDA kernel is judged against"""
    m = A.load_model("sf")
    blob, cfg = pack_model(m), A.make_env_cfg(m, "sf")
    qpos, qvel, ctrl = A.sample_reset(m, "sf", prng.split(prng.PRNGKey(0), 2))
    s = O.env_reset(blob, cfg, qpos[1], qvel[1], ctrl[1])
    rng = np.random.default_rng(1)
    eq, eo = [], []
    for _ in range(60):
        a = rng.uniform(-1, 1, 5)
        s32 = O.OrcEnvState()
        C.memmove(C.byref(s32), C.byref(s), C.sizeof(s))
        O.env_step(blob, cfg, s, a)
        O.env_step(blob, cfg, s32, a, precision="f32")
        eq.append(np.abs(np.array(s.d.qpos) - np.array(s32.d.qpos)).max())
        eo.append(np.abs(np.array(s.obs) - np.array(s32.obs)).max())
        assert s.done == s32.done
    assert np.median(eq) < 1e-6 and max(eq) < 1e-3
    assert np.median(eo) < 1e-6 and max(eo) < 1e-3
