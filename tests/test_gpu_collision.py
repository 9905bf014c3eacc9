"""Narrow phase (csrc/rsrx_physics.cuh: half-warp cooperative box_box / plane_box) against the oracle's sequential
restatement of mjx collision_convex on explicit geom pairs: face contacts with clipping, edge-edge, separated boxes,
axis-aligned stacks (the manifold tie-break case) and boxes on a plane.  Called through the C-ABI
(`rsrx_debug_narrowphase`)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from rsr_mjx_b200 import _lib

pytestmark = pytest.mark.gpu


def _rot(g, n, generic=True, yaw_only=False):
    if yaw_only:
        a = g.uniform(-np.pi, np.pi, n)
        c, s, z, o = np.cos(a), np.sin(a), np.zeros(n), np.ones(n)
        return np.stack([c, -s, z, s, c, z, z, z, o], 1).reshape(n, 3, 3)
    q = g.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).reshape(n, 3, 3)


def _gpu(pairs, plane):
    t = torch.as_tensor(pairs, dtype=torch.float32, device="cuda").contiguous()
    out = torch.empty(t.shape[0], 19, device="cuda")
    _lib.check(_lib.lib().rsrx_debug_narrowphase(t.data_ptr(), t.shape[0], int(plane), out.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream), "narrowphase")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _compare(pairs, plane, max_mismatch_frac):
    out = _gpu(pairs, plane)
    bad, nact, err_d, err_p, err_n = 0, 0, 0.0, 0.0, 0.0
    for i, q in enumerate(pairs.astype(np.float32)):
        p1, m1, s1, p2, m2, s2 = q[0:3], q[3:12], q[12:15], q[15:18], q[18:27], q[27:30]
        if plane:
            d, pos, nrm = O.plane_box(p1, m1, p2, m2, s2, precision="f32")
        else:
            d, pos, nrm = O.box_box(p1, m1, s1, p2, m2, s2, precision="f32")
        gd, gp, gn = out[i, 0:4], out[i, 4:16].reshape(4, 3), out[i, 16:19]
        ao, ag = d < 0, gd < 0
        if not np.array_equal(ao, ag):
            bad += 1
            continue
        if not ao.any():
            continue
        # the order of the manifold points inside a pair is decided by exact ties on symmetric faces (rounding breaks
        # them differently with and without FMA contraction): compare the contacts as a set
        ro = np.concatenate([pos[ao], d[ao][:, None]], 1); rg = np.concatenate([gp[ag], gd[ag][:, None]], 1)
        ro = ro[np.lexsort(np.round(ro[:, :3], 4).T[::-1])]; rg = rg[np.lexsort(np.round(rg[:, :3], 4).T[::-1])]
        if np.abs(ro - rg).max() > 1e-4 or np.abs(gn - nrm).max() > 1e-4:
            bad += 1  # a different (equally deep) axis or manifold vertex: counted, must stay rare
            continue
        nact += int(ao.sum())
        err_d = max(err_d, np.abs(ro[:, 3] - rg[:, 3]).max()); err_p = max(err_p, np.abs(ro[:, :3] - rg[:, :3]).max())
        err_n = max(err_n, np.abs(gn - nrm).max())
    assert bad <= max_mismatch_frac * len(pairs), (bad, len(pairs))
    assert nact > 0
    assert err_d < 2e-6 and err_p < 2e-6 and err_n < 2e-6, (err_d, err_p, err_n)
    return bad, nact


def _pack(p1, m1, s1, p2, m2, s2):
    n = len(p1)
    return np.concatenate([p1, m1.reshape(n, 9), s1, p2, m2.reshape(n, 9), s2], 1)


def test_box_box_generic_orientations():
    """random orientations; the deepest vertex of box 1 sits 0-1 cm under a face of box 2: face + clipping, some edges"""
    g = np.random.default_rng(0)
    n = 1500
    s1, s2 = g.uniform(0.02, 0.1, (n, 3)), g.uniform(0.1, 0.3, (n, 3))
    m1, m2 = _rot(g, n), _rot(g, n)
    p2 = g.uniform(-0.5, 0.5, (n, 3))
    face = g.integers(0, 3, n); sign = g.choice([-1.0, 1.0], n)
    nrm = m2[np.arange(n), :, face] * sign[:, None]                      # outward normal of the chosen face
    local = g.uniform(-0.6, 0.6, (n, 3)) * s2
    local[np.arange(n), face] = sign * s2[np.arange(n), face]
    on_face = p2 + np.einsum("nij,nj->ni", m2, local)
    r1 = (np.abs(np.einsum("nij,ni->nj", m1, nrm)) * s1).sum(1)          # support of box 1 along the normal
    p1 = on_face + nrm * (r1 - g.uniform(1e-3, 1e-2, n))[:, None]
    bad, nact = _compare(_pack(p1, m1, s1, p2, m2, s2), False, 0.01)
    assert nact > n  # every pair touches


def test_box_box_axis_aligned_stack_and_yaw():
    """the cube-on-table case: exact ties between manifold vertices; plus small boxes yawed on a big one"""
    g = np.random.default_rng(1)
    n = 600
    s1 = g.uniform(0.02, 0.05, (n, 3)); s2 = np.tile([0.4, 0.6, 0.02], (n, 1))
    m2 = np.tile(np.eye(3), (n, 1, 1))
    m1 = _rot(g, n, yaw_only=True)
    m1[:200] = np.eye(3)
    p2 = np.zeros((n, 3))
    p1 = np.stack([g.uniform(-0.3, 0.3, n), g.uniform(-0.5, 0.5, n), s2[:, 2] + s1[:, 2] - g.uniform(1e-4, 3e-3, n)], 1)
    p1[100:200, 0] = 0.4  # hanging over the table edge: clipped manifold
    bad, nact = _compare(_pack(p1, m1, s1, p2, m2, s2), False, 0.0)
    assert nact >= 4 * 500


def test_box_box_separated_and_edge_edge():
    g = np.random.default_rng(2)
    n = 400
    s1, s2 = g.uniform(0.02, 0.1, (n, 3)), g.uniform(0.02, 0.1, (n, 3))
    m1, m2 = _rot(g, n), _rot(g, n)
    p2 = np.zeros((n, 3))
    u = g.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    p1 = u * (np.linalg.norm(s1, axis=1) + np.linalg.norm(s2, axis=1) + 0.01)[:, None]  # bounding spheres apart
    out = _gpu(_pack(p1, m1, s1, p2, m2, s2), False)
    assert (out[:, 0:4] >= 0).all()
    # edge-edge: two long thin bars crossed at right angles, each rotated 45 degrees about its own long axis so that
    # an edge (not a face) points at the other bar
    n = 200
    c, s = np.cos(np.pi / 4), np.sin(np.pi / 4)
    m1 = np.tile(np.array([[1, 0, 0], [0, c, -s], [0, s, c]], float), (n, 1, 1))         # bar along x
    m2 = np.tile(np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], float), (n, 1, 1))         # bar along y
    s1 = np.tile([0.3, 0.02, 0.02], (n, 1)); s2 = np.tile([0.02, 0.3, 0.02], (n, 1))
    p2 = np.zeros((n, 3))
    reach = 2 * 0.02 * np.sqrt(2)
    p1 = np.stack([g.uniform(-0.1, 0.1, n), g.uniform(-0.1, 0.1, n), reach - g.uniform(1e-3, 5e-3, n)], 1)
    pairs = _pack(p1, m1, s1, p2, m2, s2)
    bad, nact = _compare(pairs, False, 0.0)
    out = _gpu(pairs, False)
    assert ((out[:, 0:4] < 0).sum(1) == 1).all()          # one contact: the edge-edge branch
    assert (np.abs(out[:, 18]) > 0.99).all()               # normal along z


def test_plane_box():
    g = np.random.default_rng(3)
    n = 800
    s2 = g.uniform(0.02, 0.1, (n, 3))
    m2 = _rot(g, n)
    m2[:200] = np.eye(3)
    r = (np.abs(m2[:, 2, :]) * s2).sum(1)  # support along the plane normal (world z)
    p2 = np.stack([g.uniform(-1, 1, n), g.uniform(-1, 1, n), r - g.uniform(-5e-3, 5e-3, n)], 1)
    p1 = np.zeros((n, 3)); m1 = np.tile(np.eye(3), (n, 1, 1)); s1 = np.zeros((n, 3))
    bad, nact = _compare(_pack(p1, m1, s1, p2, m2, s2), True, 0.0)
    assert nact > 300
