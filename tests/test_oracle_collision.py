"""Narrow-phase known-answer cases for the oracle's plane-box / box-box."""
import math

import numpy as np
import pytest

from oracle import oracle as O

I3 = np.eye(3)


def rotz(a):
    return np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]])


def roty(a):
    return np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])


@pytest.fixture(scope="module", autouse=True)
def _build(oracle_built):
    return oracle_built


def test_box_on_box_face_contact_four_corners():
    # small cube (box 2) sitting 1 mm into a big slab (box 1)
    dist, pos, n = O.box_box([0, 0, 0], I3, [1, 1, 0.1], [0.2, 0.1, 0.1 + 0.05 - 0.001], I3, [0.05, 0.05, 0.05])
    np.testing.assert_allclose(n, [0, 0, 1], atol=1e-12)  # from box 1 to box 2
    np.testing.assert_allclose(dist, -0.001, atol=1e-12)
    got = sorted(map(tuple, np.round(pos[:, :2], 6)))
    assert got == [(0.15, 0.05), (0.15, 0.15), (0.25, 0.05), (0.25, 0.15)]
    np.testing.assert_allclose(pos[:, 2], 0.1 - 0.0005, atol=1e-12)  # midway between the surfaces


def test_box_box_separated_has_no_active_contact():
    dist, _, _ = O.box_box([0, 0, 0], I3, [1, 1, 0.1], [0, 0, 0.151], I3, [0.05, 0.05, 0.05])
    assert (dist >= 0).all()
    dist, _, _ = O.box_box([0, 0, 0], I3, [0.1, 0.1, 0.1], [0.5, 0.5, 0], rotz(0.3), [0.1, 0.1, 0.1])
    assert (dist >= 0).all()


def test_box_box_rotated_face_contact_clipped():
    # box 2 rotated 45 deg about z, overhanging the edge of box 1: clipped polygon, all depths equal
    dist, pos, n = O.box_box([0, 0, 0], I3, [0.1, 0.1, 0.1], [0.1, 0, 0.1 + 0.05 - 0.002], rotz(math.pi / 4), [0.05, 0.05, 0.05])
    act = dist < 0
    assert act.sum() == 4
    np.testing.assert_allclose(dist[act], -0.002, atol=1e-9)
    assert (pos[act, 0] <= 0.1 + 1e-9).all()  # clipped to the reference face
    np.testing.assert_allclose(np.abs(n), [0, 0, 1], atol=1e-9)


def test_box_box_edge_edge():
    # two crossed boxes, each tilted 45 deg about its long axis so that edges meet
    m1 = roty(0) @ np.array([[1, 0, 0], [0, math.cos(math.pi / 4), -math.sin(math.pi / 4)], [0, math.sin(math.pi / 4), math.cos(math.pi / 4)]])
    m2 = np.array([[math.cos(math.pi / 4), 0, math.sin(math.pi / 4)], [0, 1, 0], [-math.sin(math.pi / 4), 0, math.cos(math.pi / 4)]])
    h = 0.05 * math.sqrt(2)
    gap = -0.001
    dist, pos, n = O.box_box([0, 0, 0], m1, [0.5, 0.05, 0.05], [0, 0, 2 * h + gap], m2, [0.05, 0.5, 0.05])
    act = dist < 0
    assert act.sum() == 1
    np.testing.assert_allclose(dist[act], gap, atol=1e-9)
    np.testing.assert_allclose(pos[act][0], [0, 0, h + gap / 2], atol=1e-9)
    np.testing.assert_allclose(n, [0, 0, 1], atol=1e-9)


def test_plane_box_flat_and_tilted():
    dist, pos, n = O.plane_box([0, 0, 0], I3, [0, 0, 0.099], I3, [0.1, 0.2, 0.1])
    np.testing.assert_allclose(n, [0, 0, 1])
    np.testing.assert_allclose(dist, -0.001, atol=1e-12)
    assert sorted(map(tuple, np.round(pos[:, :2], 6))) == [(-0.1, -0.2), (-0.1, 0.2), (0.1, -0.2), (0.1, 0.2)]
    # tilted: only the lowest edge (2 vertices) is within the 1 mm skin
    dist, pos, _ = O.plane_box([0, 0, 0], I3, [0, 0, 0.1245], roty(0.3), [0.1, 0.2, 0.1])
    assert (dist < 0).sum() == 2
    # far above the plane: inactive
    dist, _, _ = O.plane_box([0, 0, 0], I3, [0, 0, 1.0], I3, [0.1, 0.2, 0.1])
    assert (dist > 0).all()


def test_manifold_is_precision_stable_on_rectangles():
    """the tie-break rule: f32 and f64 pick the same four corners"""
    rng = np.random.default_rng(0)
    for _ in range(50):
        p2 = [rng.uniform(-0.3, 0.3), rng.uniform(-0.1, 0.1), 0.01 + 0.04 - rng.uniform(1e-5, 1e-3)]
        R = rotz(rng.uniform(-0.02, 0.02)) @ roty(rng.uniform(-1e-3, 1e-3))
        a = O.box_box([0, 0, 0], I3, [0.8, 0.3, 0.01], p2, R, [0.04, 0.04, 0.04])
        b = O.box_box([0, 0, 0], I3, [0.8, 0.3, 0.01], p2, R, [0.04, 0.04, 0.04], precision="f32")
        assert ((a[0] < 0) == (b[0] < 0)).all()
