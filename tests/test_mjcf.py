"""Model compiler: hand-computable cases + the shipped Airbot assets."""
import math
import os

import numpy as np
import pytest

from rsr_mjx_b200 import airbot_spec as A, mjcf

ONE_BOX = """
<mujoco>
  <compiler angle="radian"/>
  <option timestep="0.002" integrator="implicitfast"/>
  <worldbody>
    <geom type="plane" size="1 1 0.1"/>
    <body name="b" pos="0 0 1">
      <freejoint/>
      <geom name="g" type="box" size="0.1 0.2 0.3"/>
    </body>
  </worldbody>
</mujoco>
"""

PENDULUM = """
<mujoco>
  <compiler angle="radian"/>
  <option integrator="implicitfast"/>
  <default><geom contype="0" conaffinity="0"/></default>
  <worldbody>
    <body name="l1" pos="0 0 0">
      <joint name="j1" type="hinge" axis="0 1 0"/>
      <inertial pos="0 0 -0.5" mass="2" diaginertia="0.1 0.2 0.3"/>
      <body name="l2" pos="0 0 -1">
        <joint name="j2" type="hinge" axis="0 1 0"/>
        <inertial pos="0 0 -0.25" mass="1" diaginertia="0.01 0.02 0.03"/>
      </body>
    </body>
  </worldbody>
</mujoco>
"""


def test_single_box_inertia_from_geom():
    m = mjcf.compile_mjcf(ONE_BOX, from_string=True)
    assert (m.nbody, m.nq, m.nv, m.ngeom, m.npair) == (2, 7, 6, 2, 1)
    mass = 8 * 0.1 * 0.2 * 0.3 * 1000
    assert m.body_mass[1] == pytest.approx(mass)
    I = mass / 3 * np.array([0.2**2 + 0.3**2, 0.1**2 + 0.3**2, 0.1**2 + 0.2**2])
    np.testing.assert_allclose(m.body_inertia[1], I)
    np.testing.assert_allclose(m.qpos0, [0, 0, 1, 1, 0, 0, 0])
    # free body: M = diag(m,m,m,I) -> invweights are 1/m and mean(1/I)
    np.testing.assert_allclose(m.dof_invweight0[:3], 1 / mass)
    np.testing.assert_allclose(m.dof_invweight0[3:], np.mean(1 / I))
    np.testing.assert_allclose(m.body_invweight0[1], [1 / mass, np.mean(1 / I)])
    assert m.meaninertia == pytest.approx((3 * mass + I.sum()) / 6)
    # pair list: plane (type 0) first
    assert (m.pair_geom1[0], m.pair_geom2[0]) == (0, 1)


def test_double_pendulum_mass_matrix():
    m = mjcf.compile_mjcf(PENDULUM, from_string=True)
    q = np.array([0.3, -0.7])
    M, _ = mjcf.mass_matrix(m, q)
    # planar double pendulum about y: analytic M
    m1, m2, l1, c1, c2 = 2.0, 1.0, 1.0, 0.5, 0.25
    I1, I2 = 0.2, 0.02
    M22 = I2 + m2 * c2**2
    M12 = M22 + m2 * l1 * c2 * math.cos(q[1])
    M11 = I1 + m1 * c1**2 + I2 + m2 * (l1**2 + c2**2 + 2 * l1 * c2 * math.cos(q[1]))
    np.testing.assert_allclose(M, [[M11, M12], [M12, M22]], rtol=1e-12)
    assert m.dof_parentid.tolist() == [-1, 0]


def test_euler_is_intrinsic_xyz():
    q = mjcf.euler_to_quat([0.3, -0.2, 0.5], "xyz")
    Rx = lambda a: np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
    Ry = lambda a: np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
    Rz = lambda a: np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]])
    np.testing.assert_allclose(mjcf.quat_to_mat(q), Rx(0.3) @ Ry(-0.2) @ Rz(0.5), atol=1e-12)


@pytest.mark.parametrize("kind,dims", [("sf", (14, 10, 22, 20, 5, 23, 1, 45)), ("cube", (14, 10, 22, 20, 5, 23, 1, 45)),
                                       ("T", (14, 9, 15, 14, 5, 25, 3, 60))])
def test_airbot_model_dims(kind, dims):
    """sizes derived in SURVEY.md §8a / §A.1"""
    m = A.load_model(kind)
    assert (m.nbody, m.njnt, m.nq, m.nv, m.nu, m.ngeom, m.nsite, m.npair) == dims
    nplane = int((m.geom_type[m.pair_geom1] == mjcf.GEOM_PLANE).sum())
    assert nplane == 15 and m.npair - nplane == (45 if kind == "T" else 30)
    ids = A.env_ids(m, kind)
    assert ids["cube_id"] == 13 and ids["target_pos_id"] == 12
    assert ids["joint_id"].tolist() == [0, 1, 2, 3, 4, 5]
    # block structure of M: arm dofs 0..7, then free bodies
    M, _ = mjcf.mass_matrix(m, m.qpos0)
    assert np.allclose(M[:8, 8:], 0) and np.all(np.linalg.eigvalsh(M) > 0)
    # position actuators: gain kp, bias -kp*q
    np.testing.assert_allclose(m.act_gainprm[:, 0], [1000, 1000, 1000, 350, 100])
    np.testing.assert_allclose(m.act_biasprm[:, 1], [-1000, -1000, -1000, -350, -100])
    assert m.act_trnid.tolist() == [0, 1, 2, 4, 5]
    # equality endleft = -endright
    assert m.neq == 1 and m.eq_data[0].tolist() == [0, -1, 0, 0, 0]


def test_sf_contact_parameters():
    m = A.load_model("sf")
    cube, table = m.geom("geom_for_push"), m.geom("table-b")
    np.testing.assert_allclose(m.geom_friction[cube], [1.22, 0.1, 0.1])
    np.testing.assert_allclose(m.geom_friction[table], [0.4, 0.005, 0.0001])
    np.testing.assert_allclose(m.geom_solimp[cube], [0.8, 1.0, 0.01, 0.5, 2.0])
    assert (m.geom_condim == 4).all()
    # free 0.5 kg cube with its COM at the body origin
    np.testing.assert_allclose(m.body_invweight0[13, 0], 2.0, rtol=1e-12)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout not mounted")
def test_assets_equal_reference_models():
    ref = {"sf": "test/sf.xml", "cube": "ppo_train/airbot_training/cube.xml", "T": "ppo_train/airbot_training/T_shape.xml"}
    for kind, rel in ref.items():
        a = A.load_model(kind)
        b = mjcf.compile_mjcf(os.path.join("/root/reference", rel))
        assert a.names == b.names
        for k in a.arrays:
            np.testing.assert_array_equal(a.arrays[k], b.arrays[k], err_msg=f"{kind}:{k}")
        assert (a.timestep, a.iterations, a.meaninertia) == (b.timestep, b.iterations, b.meaninertia)
