"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE — see rsr_oracle.c).

Only tests/ (and the parity / golden-vector scripts under tools/ that serve them), __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import this module.  PARITY UNPINNED (no runnable MJX here).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rsr_mjx_b200.model import (MAXBODY, MAXGEOM, MAXQ, MAXSITE, MAXU, MAXV, MAXPAIR, EnvCfg, ModelBlob)

_HERE = os.path.dirname(os.path.abspath(__file__))
MAXOBS = 24
MAXCON = 4 * MAXPAIR
MAXEFC = 2 + 2 * MAXV + 6 * MAXCON
_d, _i = C.c_double, C.c_int32


class OrcData(C.Structure):
    _fields_ = [
        ("qpos", _d * MAXQ), ("qvel", _d * MAXV), ("ctrl", _d * MAXU), ("qacc_warmstart", _d * MAXV),
        ("time", _d),
        ("xpos", (_d * 3) * MAXBODY), ("xquat", (_d * 4) * MAXBODY),
        ("site_xpos", (_d * 3) * MAXSITE), ("geom_xpos", (_d * 3) * MAXGEOM),
        ("qacc", _d * MAXV), ("qacc_smooth", _d * MAXV), ("qfrc_constraint", _d * MAXV),
        ("qfrc_bias", _d * MAXV), ("qfrc_actuator", _d * MAXV),
        ("ncon", _i), ("ncon_active", _i), ("nefc", _i), ("nefc_active", _i), ("solver_niter", _i),
        ("ls_total", _i), ("pad", _i * 2),
    ]

    def arr(self, name):
        return np.ctypeslib.as_array(getattr(self, name))


class OrcEnvState(C.Structure):
    _fields_ = [
        ("d", OrcData), ("first", OrcData),
        ("obs", _d * MAXOBS), ("first_obs", _d * MAXOBS),
        ("reward", _d), ("done", _d), ("truncation", _d), ("steps", _d),
        ("target_pos", _d * 3), ("target2_pos", _d * 3), ("new_pos", _d * 2), ("site_pos", _d * 3),
        ("obj_pos", _d * 3), ("last_action", _d), ("xita", _d), ("target_w", _d),
        ("metrics", _d * 5),
    ]


class OrcContact(C.Structure):
    _fields_ = [("dist", _d), ("pos", _d * 3), ("frame", _d * 9), ("friction", _d * 5), ("solref", _d * 2),
                ("solimp", _d * 5), ("geom1", _i), ("geom2", _i)]


def build(force: bool = False):
    """Compile liboracle_f64.so / liboracle_f32.so next to the source."""
    targets = [os.path.join(_HERE, f"liboracle_{p}.so") for p in ("f64", "f32")]
    src = os.path.join(_HERE, "rsr_oracle.c")
    stale = force or any((not os.path.exists(t)) or os.path.getmtime(t) < os.path.getmtime(src) for t in targets)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, stdout=subprocess.DEVNULL)
    return targets


_LIBS = {}


def lib(precision: str = "f64"):
    if precision not in _LIBS:
        path = os.path.join(_HERE, f"liboracle_{precision}.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        P = C.POINTER
        L.oracle_env_reset.argtypes = [P(ModelBlob), P(EnvCfg), P(_d), P(_d), P(_d), P(OrcEnvState), _i]
        L.oracle_env_step.argtypes = [P(ModelBlob), P(EnvCfg), P(OrcEnvState), P(_d), _i]
        L.oracle_rollout.argtypes = [P(ModelBlob), _i, P(EnvCfg), P(OrcEnvState), _i, P(_d), _i, _i, _i]
        L.oracle_forward.argtypes = [P(ModelBlob), P(OrcData), _i]
        L.oracle_step.argtypes = [P(ModelBlob), P(OrcData), _i, _i]
        L.oracle_inspect.argtypes = [P(ModelBlob), P(OrcData), _i, P(_d), P(OrcContact), P(_i), P(_d), P(_d),
                                     P(_d), P(_d), P(_i), P(_d), P(_d)]
        L.oracle_box_box.argtypes = [P(_d)] * 9
        L.oracle_plane_box.argtypes = [P(_d)] * 8
        L.oracle_box_box.restype = None
        L.oracle_plane_box.restype = None
        assert L.oracle_sizeof_model() == C.sizeof(ModelBlob), (L.oracle_sizeof_model(), C.sizeof(ModelBlob))
        assert L.oracle_sizeof_cfg() == C.sizeof(EnvCfg)
        assert L.oracle_sizeof_data() == C.sizeof(OrcData)
        assert L.oracle_sizeof_env_state() == C.sizeof(OrcEnvState)
        _LIBS[precision] = L
    return _LIBS[precision]


def _p(a):
    return a.ctypes.data_as(C.POINTER(_d))


def _dbl(x, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    if n is not None and a.size < n:
        a = np.concatenate([a, np.zeros(n - a.size)])
    return a


def make_data(blob: ModelBlob, qpos, qvel, ctrl=None, warmstart=None) -> OrcData:
    d = OrcData()
    d.qpos[:blob.nq] = list(np.asarray(qpos, float))
    d.qvel[:blob.nv] = list(np.asarray(qvel, float))
    if ctrl is not None:
        d.ctrl[:blob.nu] = list(np.asarray(ctrl, float))
    if warmstart is not None:
        d.qacc_warmstart[:blob.nv] = list(np.asarray(warmstart, float))
    return d


def forward(blob, d: OrcData, precision="f64", dense=False):
    rc = lib(precision).oracle_forward(C.byref(blob), C.byref(d), int(dense))
    assert rc == 0
    return d


def step(blob, d: OrcData, nsteps=1, precision="f64", dense=False):
    rc = lib(precision).oracle_step(C.byref(blob), C.byref(d), nsteps, int(dense))
    assert rc == 0
    return d


def inspect(blob, d: OrcData, precision="f64", dense=False):
    """forward() + internals: dict(M, contacts, J, D, aref, force, cdof, subtree_com)."""
    nv = blob.nv
    M = np.zeros((nv, nv))
    con = (OrcContact * MAXCON)()
    ncon = _i(0)
    J = np.zeros((MAXEFC, nv))
    D = np.zeros(MAXEFC)
    aref = np.zeros(MAXEFC)
    force = np.zeros(MAXEFC)
    nefc = _i(0)
    cdof = np.zeros((nv, 6))
    sc = np.zeros((blob.nbody, 3))
    rc = lib(precision).oracle_inspect(C.byref(blob), C.byref(d), int(dense), _p(M), con, C.byref(ncon), _p(J), _p(D),
                                       _p(aref), _p(force), C.byref(nefc), _p(cdof), _p(sc))
    assert rc == 0
    n = nefc.value
    contacts = [dict(dist=c.dist, pos=np.array(c.pos), frame=np.array(c.frame).reshape(3, 3),
                     friction=np.array(c.friction), solref=np.array(c.solref), solimp=np.array(c.solimp),
                     geom1=c.geom1, geom2=c.geom2) for c in con[:ncon.value]]
    return dict(M=M, contacts=contacts, J=J[:n].copy(), D=D[:n].copy(), aref=aref[:n].copy(),
                force=force[:n].copy(), cdof=cdof, subtree_com=sc)


def box_box(p1, m1, s1, p2, m2, s2, precision="f64"):
    dist = np.zeros(4)
    pos = np.zeros((4, 3))
    nrm = np.zeros(3)
    args = [_dbl(p1), _dbl(m1), _dbl(s1), _dbl(p2), _dbl(m2), _dbl(s2)]
    lib(precision).oracle_box_box(*[_p(a) for a in args], _p(dist), _p(pos), _p(nrm))
    return dist, pos, nrm


def plane_box(pp, pm, bp, bm, size, precision="f64"):
    dist = np.zeros(4)
    pos = np.zeros((4, 3))
    nrm = np.zeros(3)
    args = [_dbl(pp), _dbl(pm), _dbl(bp), _dbl(bm), _dbl(size)]
    lib(precision).oracle_plane_box(*[_p(a) for a in args], _p(dist), _p(pos), _p(nrm))
    return dist, pos, nrm


def env_reset(blob, cfg: EnvCfg, qpos, qvel, ctrl, precision="f64", dense=False) -> OrcEnvState:
    s = OrcEnvState()
    q, v, c = _dbl(qpos, MAXQ), _dbl(qvel, MAXV), _dbl(ctrl, MAXU)
    rc = lib(precision).oracle_env_reset(C.byref(blob), C.byref(cfg), _p(q), _p(v), _p(c), C.byref(s), int(dense))
    assert rc == 0
    return s


def env_step(blob, cfg: EnvCfg, s: OrcEnvState, action, precision="f64", dense=False) -> OrcEnvState:
    a = _dbl(action, MAXU)
    rc = lib(precision).oracle_env_step(C.byref(blob), C.byref(cfg), C.byref(s), _p(a), int(dense))
    assert rc == 0
    return s


def rollout(blobs, cfg: EnvCfg, states, actions, precision="f64", dense=False, nthreads=0):
    """blobs: one ModelBlob (shared) or a ctypes array of N; states: ctypes array
    (OrcEnvState * N); actions: float64 [T, N, nu]."""
    N = len(states)
    actions = np.ascontiguousarray(actions, dtype=np.float64)
    T = actions.shape[0]
    assert actions.shape[1] == N
    if isinstance(blobs, ModelBlob):
        bp, stride = C.pointer(blobs), 0
    else:
        assert len(blobs) == N
        bp, stride = C.cast(blobs, C.POINTER(ModelBlob)), 1
    rc = lib(precision).oracle_rollout(bp, stride, C.byref(cfg), C.cast(states, C.POINTER(OrcEnvState)), N,
                                       _p(actions), T, int(dense), nthreads)
    assert rc == 0
    return states


def max_threads():
    return lib().oracle_max_threads()
