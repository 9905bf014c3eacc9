"""MJCF-subset model compiler (host side, float64 NumPy, no mujoco dependency).

The reference builds its physics model with the MuJoCo C compiler and brax:
``mujoco.MjModel.from_xml_path(path)`` then ``brax.io.mjcf.load_model``
(reference test/airbot.py:43-47, ppo_train/airbot_training/cube_env.py:37-43,
ppo_train/airbot_training/T_shape_env.py:39-45).  Neither library exists in this
image, so this module restates the compile semantics of ``mujoco==3.2.4`` for
exactly the MJCF features the three Airbot models use:

* ``<compiler angle inertiafromgeom inertiagrouprange>``, ``<option>``,
  nested ``<default class>`` trees (all ``<default>`` sections are read before
  ``<worldbody>``, whatever their position in the file),
* bodies with ``pos``/``quat``/``euler`` (intrinsic xyz), explicit ``<inertial>``
  or box inertia from geoms (density 1000, parallel-axis composition),
* hinge / slide / free joints, box and plane geoms, sphere sites,
* ``<position>`` actuators, joint ``<equality>``, ``<exclude>`` (parsed; irrelevant to
  the resulting pair list), fixed tendons are ignored (no dynamic effect),
* ``mj_setConst`` products: ``qpos0``, ``dof_invweight0``, ``body_invweight0``,
  ``stat.meaninertia``,
* the static collision-pair filter of ``mjx._src.collision_driver.geom_pairs``.

The result is a :class:`Model` of NumPy arrays named like ``mjModel`` fields.
"""
from __future__ import annotations

import dataclasses
import math
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional

import numpy as np

MJ_MINVAL = 1e-15

JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_PLANE, GEOM_SPHERE, GEOM_BOX = 0, 2, 6
_GEOM_TYPES = {"plane": 0, "hfield": 1, "sphere": 2, "capsule": 3, "ellipsoid": 4,
               "cylinder": 5, "box": 6, "mesh": 7}


# --------------------------------------------------------------------------- math
def quat_mul(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def quat_to_mat(q):
    w, x, y, z = np.asarray(q, float)
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat_to_quat(m):
    """Rotation matrix -> unit quaternion (w>=0)."""
    m = np.asarray(m, float)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
    q = np.array(q)
    if q[0] < 0:
        q = -q
    return q / np.linalg.norm(q)


def rotate(v, q):
    return quat_to_mat(q) @ np.asarray(v, float)


def euler_to_quat(e, seq="xyz"):
    """MuJoCo eulerseq semantics: lowercase = rotating (intrinsic) axes,
    accumulate by post-multiplication; uppercase = fixed axes."""
    q = np.array([1.0, 0, 0, 0])
    for ang, ax in zip(e, seq):
        t = np.array([math.cos(ang / 2), 0.0, 0.0, 0.0])
        t["xyz".index(ax.lower()) + 1] = math.sin(ang / 2)
        q = quat_mul(q, t) if ax.islower() else quat_mul(t, q)
    return q


def _floats(s, n=None):
    v = np.array([float(x) for x in s.split()], float)
    if n is not None and len(v) != n:
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


# ------------------------------------------------------------------- model record
@dataclasses.dataclass
class Model:
    """Compiled model; arrays are float64 / int32 and named like mjModel fields."""
    # sizes
    nbody: int = 0
    njnt: int = 0
    nq: int = 0
    nv: int = 0
    nu: int = 0
    ngeom: int = 0
    nsite: int = 0
    npair: int = 0
    neq: int = 0
    # options
    timestep: float = 0.002
    gravity: np.ndarray = None
    iterations: int = 100
    ls_iterations: int = 50
    tolerance: float = 1e-8
    ls_tolerance: float = 0.01
    impratio: float = 1.0
    integrator: str = "Euler"
    meaninertia: float = 1.0
    arrays: Dict[str, np.ndarray] = dataclasses.field(default_factory=dict)
    names: Dict[str, Dict[str, int]] = dataclasses.field(default_factory=dict)

    def __getattr__(self, k):
        arrays = self.__dict__.get("arrays", {})
        if k in arrays:
            return arrays[k]
        raise AttributeError(k)

    def body(self, name):
        return self.names["body"][name]

    def geom(self, name):
        return self.names["geom"][name]

    def site(self, name):
        return self.names["site"][name]

    def joint(self, name):
        return self.names["joint"][name]

    def replace_arrays(self, **kw) -> "Model":
        m = dataclasses.replace(self)
        m.arrays = dict(self.arrays)
        for k, v in kw.items():
            if k not in m.arrays:
                raise KeyError(k)
            m.arrays[k] = np.asarray(v, dtype=m.arrays[k].dtype).reshape(m.arrays[k].shape)
        return m


# ---------------------------------------------------------------------- defaults
class _Defaults:
    """One <default> class: per-element attribute dicts, inherited from parent."""

    def __init__(self, parent: Optional["_Defaults"] = None):
        self.attrs: Dict[str, Dict[str, str]] = {}
        if parent is not None:
            self.attrs = {k: dict(v) for k, v in parent.attrs.items()}

    def update(self, tag, a):
        self.attrs.setdefault(tag, {}).update(a)

    def get(self, tag):
        return self.attrs.get(tag, {})


def _read_defaults(root) -> Dict[str, _Defaults]:
    classes: Dict[str, _Defaults] = {"main": _Defaults()}

    def rec(el, cur: _Defaults):
        for ch in el:
            if ch.tag == "default":
                name = ch.get("class")
                if name is None:
                    raise ValueError("nested <default> needs a class")
                d = _Defaults(cur)
                classes[name] = d
                rec(ch, d)
            else:
                cur.update(ch.tag, dict(ch.attrib))

    # pass 1: top-level sections update 'main' first so children inherit them
    for sec in root.findall("default"):
        for ch in sec:
            if ch.tag != "default":
                classes["main"].update(ch.tag, dict(ch.attrib))
    for sec in root.findall("default"):
        for ch in sec:
            if ch.tag == "default":
                name = ch.get("class")
                d = _Defaults(classes["main"])
                classes[name] = d
                rec(ch, d)
    return classes


# ----------------------------------------------------------------------- parsing
def _orientation(a: Dict[str, str], eulerseq: str, angle_rad: bool) -> np.ndarray:
    if "quat" in a:
        q = _floats(a["quat"], 4)
        return q / np.linalg.norm(q)
    if "euler" in a:
        e = _floats(a["euler"], 3)
        if not angle_rad:
            e = np.deg2rad(e)
        return euler_to_quat(e, eulerseq)
    for k in ("axisangle", "xyaxes", "zaxis"):
        if k in a:
            raise NotImplementedError(f"orientation spec {k!r}")
    return np.array([1.0, 0, 0, 0])


def _box_inertia(size, mass):
    a, b, c = size
    return mass / 3.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])


def compile_mjcf(path_or_string: str, from_string: bool = False) -> Model:
    root = ET.fromstring(path_or_string) if from_string else ET.parse(path_or_string).getroot()
    if root.tag != "mujoco":
        raise ValueError("not an MJCF file")

    # ---- compiler / option
    comp = {}
    for c in root.findall("compiler"):
        comp.update(c.attrib)
    angle_rad = comp.get("angle", "degree") == "radian"
    eulerseq = comp.get("eulerseq", "xyz")
    inertiafromgeom = comp.get("inertiafromgeom", "auto")
    igr = _floats(comp.get("inertiagrouprange", "0 5"), 2).astype(int)
    autolimits = comp.get("autolimits", "true") == "true"

    opt = {}
    for o in root.findall("option"):
        opt.update(o.attrib)
        for fl in o.findall("flag"):
            for k, v in fl.attrib.items():
                if v != ("enable" if k in ("constraint", "equality", "frictionloss", "limit", "contact",
                                           "passive", "gravity", "clampctrl", "warmstart", "filterparent",
                                           "actuation", "refsafe", "sensor", "midphase", "eulerdamp") else "disable"):
                    raise NotImplementedError(f"<flag {k}={v}> is not supported")
    for k in ("solver", "cone", "jacobian"):
        if k in opt and opt[k].lower() not in ("newton", "pyramidal", "dense", "auto"):
            raise NotImplementedError(f"option {k}={opt[k]}")

    classes = _read_defaults(root)

    def resolve(tag, el, childclass):
        cls = el.get("class") or childclass or "main"
        if cls not in classes:
            raise ValueError(f"unknown default class {cls!r}")
        a = dict(classes[cls].get(tag))
        a.update(el.attrib)
        a.pop("class", None)
        return a

    # ---- kinematic tree (pre-order DFS, as mjCModel::MakeLists)
    bodies: List[dict] = [dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                               inertial=None, joints=[], geoms=[], sites=[])]
    joints: List[dict] = []
    geoms: List[dict] = []
    sites: List[dict] = []

    def parse_geom(el, bid, childclass):
        a = resolve("geom", el, childclass)
        gtype = _GEOM_TYPES[a.get("type", "sphere")]
        if gtype not in (GEOM_PLANE, GEOM_BOX):
            raise NotImplementedError(f"geom type {a.get('type')}")
        size = np.zeros(3)
        sz = _floats(a.get("size", "0 0 0"))
        size[:len(sz)] = sz
        solimp = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
        if "solimp" in a:
            v = _floats(a["solimp"])
            solimp[:len(v)] = v
        fr = np.array([1.0, 0.005, 0.0001])
        if "friction" in a:
            v = _floats(a["friction"])
            fr[:len(v)] = v
        g = dict(name=a.get("name"), body=bid, type=gtype, size=size,
                 pos=_floats(a.get("pos", "0 0 0"), 3), quat=_orientation(a, eulerseq, angle_rad),
                 contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
                 condim=int(a.get("condim", 3)), priority=int(a.get("priority", 0)),
                 group=int(a.get("group", 0)), friction=fr,
                 solref=_floats(a.get("solref", "0.02 1"), 2), solimp=solimp,
                 solmix=float(a.get("solmix", 1.0)), margin=float(a.get("margin", 0.0)),
                 gap=float(a.get("gap", 0.0)), density=float(a.get("density", 1000.0)),
                 mass=(float(a["mass"]) if "mass" in a else None))
        geoms.append(g)
        bodies[bid]["geoms"].append(len(geoms) - 1)

    def parse_joint(el, bid, childclass, free=False):
        a = resolve("joint", el, childclass) if not free else dict(el.attrib)
        jt = "free" if free else a.get("type", "hinge")
        jtype = {"free": JNT_FREE, "slide": JNT_SLIDE, "hinge": JNT_HINGE}.get(jt)
        if jtype is None:
            raise NotImplementedError(f"joint type {jt}")
        axis = _floats(a.get("axis", "0 0 1"), 3)
        axis = axis / np.linalg.norm(axis)
        has_range = "range" in a
        rng = _floats(a.get("range", "0 0"), 2)
        lim = a.get("limited", "auto")
        limited = (has_range and autolimits) if lim == "auto" else (lim == "true")
        if not angle_rad and jtype == JNT_HINGE:
            rng = np.deg2rad(rng)
        has_afr = "actuatorfrcrange" in a
        afr = _floats(a.get("actuatorfrcrange", "0 0"), 2)
        afl = a.get("actuatorfrclimited", "auto")
        actfrclimited = (has_afr and autolimits) if afl == "auto" else (afl == "true")
        for k in ("stiffness", "springref", "ref", "armature"):
            if k in a and float(a[k]) != 0.0 and k != "armature":
                raise NotImplementedError(f"joint {k}")
        solimp_l = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
        if "solimplimit" in a:
            v = _floats(a["solimplimit"])
            solimp_l[:len(v)] = v
        solimp_f = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
        if "solimpfriction" in a:
            v = _floats(a["solimpfriction"])
            solimp_f[:len(v)] = v
        j = dict(name=a.get("name"), body=bid, type=jtype, pos=_floats(a.get("pos", "0 0 0"), 3),
                 axis=axis, range=rng, limited=bool(limited), actfrcrange=afr,
                 actfrclimited=bool(actfrclimited), damping=float(a.get("damping", 0.0)),
                 frictionloss=float(a.get("frictionloss", 0.0)), armature=float(a.get("armature", 0.0)),
                 margin=float(a.get("margin", 0.0)),
                 solref_limit=_floats(a.get("solreflimit", "0.02 1"), 2), solimp_limit=solimp_l,
                 solref_friction=_floats(a.get("solreffriction", "0.02 1"), 2), solimp_friction=solimp_f)
        joints.append(j)
        bodies[bid]["joints"].append(len(joints) - 1)

    def parse_site(el, bid, childclass):
        a = resolve("site", el, childclass)
        sites.append(dict(name=a.get("name"), body=bid, pos=_floats(a.get("pos", "0 0 0"), 3),
                          quat=_orientation(a, eulerseq, angle_rad)))
        bodies[bid]["sites"].append(len(sites) - 1)

    def parse_body_children(el, bid, childclass):
        # elements of this body first (MakeLists adds a body's own elements
        # before recursing), then child bodies in order
        for ch in el:
            if ch.tag == "geom":
                parse_geom(ch, bid, childclass)
            elif ch.tag == "joint":
                parse_joint(ch, bid, childclass)
            elif ch.tag == "freejoint":
                parse_joint(ch, bid, childclass, free=True)
            elif ch.tag == "site":
                parse_site(ch, bid, childclass)
            elif ch.tag == "inertial":
                a = ch.attrib
                if "fullinertia" in a:
                    raise NotImplementedError("fullinertia")
                q = _orientation(a, eulerseq, angle_rad)
                bodies[bid]["inertial"] = dict(pos=_floats(a.get("pos", "0 0 0"), 3), quat=q,
                                               mass=float(a["mass"]),
                                               diaginertia=_floats(a.get("diaginertia", "0 0 0"), 3))
            elif ch.tag in ("light", "camera", "body"):
                pass
            else:
                raise NotImplementedError(f"<{ch.tag}> inside body")
        for ch in el:
            if ch.tag == "body":
                a = ch.attrib
                cc = a.get("childclass", childclass)
                bodies.append(dict(name=a.get("name"), parent=bid, pos=_floats(a.get("pos", "0 0 0"), 3),
                                   quat=_orientation(a, eulerseq, angle_rad), inertial=None,
                                   joints=[], geoms=[], sites=[]))
                parse_body_children(ch, len(bodies) - 1, cc)

    # MuJoCo merges repeated <worldbody> sections; elements of the world body
    # come first in the geom list because the world is body 0.
    world_sections = root.findall("worldbody")
    # first pass: world's own elements from every section, then child bodies
    for sec in world_sections:
        for ch in sec:
            if ch.tag == "geom":
                parse_geom(ch, 0, None)
            elif ch.tag == "site":
                parse_site(ch, 0, None)
    # child bodies: need body-major ordering of geoms, so re-number afterwards
    for sec in world_sections:
        for ch in sec:
            if ch.tag == "body":
                a = ch.attrib
                bodies.append(dict(name=a.get("name"), parent=0, pos=_floats(a.get("pos", "0 0 0"), 3),
                                   quat=_orientation(a, eulerseq, angle_rad), inertial=None,
                                   joints=[], geoms=[], sites=[]))
                parse_body_children(ch, len(bodies) - 1, a.get("childclass"))

    nbody = len(bodies)
    # geoms / joints / sites were appended body by body in DFS order already
    # (a body's elements are parsed before its children), so ids are final.

    # ---- ids and addresses
    njnt = len(joints)
    jnt_qposadr, jnt_dofadr = [], []
    nq = nv = 0
    for j in joints:
        jnt_qposadr.append(nq)
        jnt_dofadr.append(nv)
        if j["type"] == JNT_FREE:
            if bodies[j["body"]]["parent"] != 0:
                raise ValueError("free joint must be on a child of the world")
            nq += 7
            nv += 6
        else:
            nq += 1
            nv += 1

    body_parentid = np.array([b["parent"] for b in bodies], np.int32)
    body_jntnum = np.array([len(b["joints"]) for b in bodies], np.int32)
    body_jntadr = np.array([b["joints"][0] if b["joints"] else -1 for b in bodies], np.int32)
    body_dofnum = np.zeros(nbody, np.int32)
    body_dofadr = -np.ones(nbody, np.int32)
    for bi, b in enumerate(bodies):
        for ji in b["joints"]:
            if body_dofadr[bi] < 0:
                body_dofadr[bi] = jnt_dofadr[ji]
            body_dofnum[bi] += 6 if joints[ji]["type"] == JNT_FREE else 1
    body_weldid = np.zeros(nbody, np.int32)
    body_rootid = np.zeros(nbody, np.int32)
    body_depth = np.zeros(nbody, np.int32)
    for bi in range(1, nbody):
        p = body_parentid[bi]
        body_weldid[bi] = bi if body_jntnum[bi] > 0 else body_weldid[p]
        body_rootid[bi] = bi if p == 0 else body_rootid[p]
        body_depth[bi] = body_depth[p] + 1

    dof_bodyid = np.zeros(nv, np.int32)
    dof_jntid = np.zeros(nv, np.int32)
    for ji, j in enumerate(joints):
        w = 6 if j["type"] == JNT_FREE else 1
        dof_bodyid[jnt_dofadr[ji]:jnt_dofadr[ji] + w] = j["body"]
        dof_jntid[jnt_dofadr[ji]:jnt_dofadr[ji] + w] = ji
    dof_parentid = -np.ones(nv, np.int32)
    for d in range(nv):
        b = dof_bodyid[d]
        if d > body_dofadr[b]:
            dof_parentid[d] = d - 1
        else:
            p = body_parentid[b]
            while p > 0 and body_dofnum[p] == 0:
                p = body_parentid[p]
            if p > 0:
                dof_parentid[d] = body_dofadr[p] + body_dofnum[p] - 1

    # ---- inertial properties
    body_mass = np.zeros(nbody)
    body_inertia = np.zeros((nbody, 3))
    body_ipos = np.zeros((nbody, 3))
    body_iquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
    for bi, b in enumerate(bodies):
        if bi == 0:
            continue
        explicit = b["inertial"] is not None
        if explicit:
            it = b["inertial"]
            body_mass[bi] = it["mass"]
            body_inertia[bi] = it["diaginertia"]
            body_ipos[bi] = it["pos"]
            body_iquat[bi] = it["quat"]
        use_geoms = inertiafromgeom == "true" or (inertiafromgeom == "auto" and not explicit)
        if use_geoms:
            sel = [gi for gi in b["geoms"] if igr[0] <= geoms[gi]["group"] <= igr[1]]
            if not sel:
                continue  # InertiaFromGeom returns without touching the body
            gm, gI = [], []
            for gi in sel:
                g = geoms[gi]
                if g["type"] != GEOM_BOX:
                    raise NotImplementedError("inertia from non-box geom")
                vol = 8.0 * g["size"][0] * g["size"][1] * g["size"][2]
                mass = g["mass"] if g["mass"] is not None else vol * g["density"]
                gm.append(mass)
                gI.append(_box_inertia(g["size"], mass))
            if len(sel) == 1:
                g = geoms[sel[0]]
                body_mass[bi] = gm[0]
                body_inertia[bi] = gI[0]
                body_ipos[bi] = g["pos"]
                body_iquat[bi] = g["quat"]
            else:
                mtot = float(sum(gm))
                com = sum(m_ * geoms[gi]["pos"] for m_, gi in zip(gm, sel)) / max(mtot, MJ_MINVAL)
                I = np.zeros((3, 3))
                for m_, I_, gi in zip(gm, gI, sel):
                    R = quat_to_mat(geoms[gi]["quat"])
                    d = geoms[gi]["pos"] - com
                    I += R @ np.diag(I_) @ R.T + m_ * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
                w, V = np.linalg.eigh(I)
                order = np.argsort(-w)
                w, V = w[order], V[:, order]
                if np.linalg.det(V) < 0:
                    V[:, 2] = -V[:, 2]
                body_mass[bi] = mtot
                body_inertia[bi] = w
                body_ipos[bi] = com
                body_iquat[bi] = mat_to_quat(V)

    # ---- qpos0
    qpos0 = np.zeros(nq)
    for ji, j in enumerate(joints):
        if j["type"] == JNT_FREE:
            b = bodies[j["body"]]
            qpos0[jnt_qposadr[ji]:jnt_qposadr[ji] + 3] = b["pos"]
            qpos0[jnt_qposadr[ji] + 3:jnt_qposadr[ji] + 7] = b["quat"]

    # ---- actuators
    acts = []
    for sec in root.findall("actuator"):
        for el in sec:
            if el.tag != "position":
                raise NotImplementedError(f"actuator <{el.tag}>")
            a = resolve("position", el, None)
            if "joint" not in a:
                raise NotImplementedError("only joint transmission")
            kp = float(a.get("kp", 1.0))
            kv = float(a.get("kv", 0.0))
            has_cr = "ctrlrange" in a
            cl = a.get("ctrllimited", "auto")
            has_fr = "forcerange" in a
            fl = a.get("forcelimited", "auto")
            acts.append(dict(name=a.get("name"), joint=a["joint"], kp=kp, kv=kv,
                             ctrlrange=_floats(a.get("ctrlrange", "0 0"), 2),
                             ctrllimited=(has_cr and autolimits) if cl == "auto" else cl == "true",
                             forcerange=_floats(a.get("forcerange", "0 0"), 2),
                             forcelimited=(has_fr and autolimits) if fl == "auto" else fl == "true",
                             gear=_floats(a.get("gear", "1"))[0]))
    nu = len(acts)

    jname = {j["name"]: i for i, j in enumerate(joints) if j["name"]}
    bname = {b["name"]: i for i, b in enumerate(bodies) if b["name"]}
    gname = {g["name"]: i for i, g in enumerate(geoms) if g["name"]}
    sname = {s["name"]: i for i, s in enumerate(sites) if s["name"]}

    # ---- equality
    eqs = []
    for sec in root.findall("equality"):
        for el in sec:
            if el.tag != "joint":
                raise NotImplementedError(f"equality <{el.tag}>")
            a = resolve("equality", el, None)
            a.update(el.attrib)
            pc = np.array([0.0, 1.0, 0, 0, 0])
            if "polycoef" in a:
                v = _floats(a["polycoef"])
                pc[:len(v)] = v
            solimp = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
            if "solimp" in a:
                v = _floats(a["solimp"])
                solimp[:len(v)] = v
            eqs.append(dict(j1=jname[a["joint1"]], j2=jname[a["joint2"]] if "joint2" in a else -1,
                            data=pc, solref=_floats(a.get("solref", "0.02 1"), 2), solimp=solimp))
    neq = len(eqs)

    # ---- excludes
    exclude = set()
    for sec in root.findall("contact"):
        for el in sec:
            if el.tag == "exclude":
                b1, b2 = bname[el.get("body1")], bname[el.get("body2")]
                exclude.add((min(b1, b2) << 16) + max(b1, b2))
            else:
                raise NotImplementedError("explicit contact pairs")

    # ---- collision pair filter (mjx collision_driver.geom_pairs)
    ngeom = len(geoms)
    gcon = [g["contype"] | g["conaffinity"] for g in geoms]
    pairs = []
    for b1 in range(nbody):
        g1s = [g for g in bodies[b1]["geoms"] if gcon[g]]
        if not g1s:
            continue
        w1 = body_weldid[b1]
        w1p = body_weldid[body_parentid[w1]]
        for b2 in range(b1, nbody):
            g2s = [g for g in bodies[b2]["geoms"] if gcon[g]]
            if not g2s:
                continue
            if ((b1 << 16) + b2) in exclude:
                continue
            w2 = body_weldid[b2]
            if w1 == w2:
                continue
            w2p = body_weldid[body_parentid[w2]]
            if w1 != 0 and w2 != 0 and (w1 == w2p or w2 == w1p):
                continue
            for g1 in g1s:
                for g2 in g2s:
                    a_, b_ = g1, g2
                    if geoms[a_]["type"] > geoms[b_]["type"]:
                        a_, b_ = b_, a_
                    if geoms[a_]["type"] == GEOM_PLANE and geoms[b_]["type"] == GEOM_PLANE:
                        continue
                    mask = (geoms[a_]["contype"] & geoms[b_]["conaffinity"]) | (
                        geoms[b_]["contype"] & geoms[a_]["conaffinity"])
                    if not mask:
                        continue
                    if geoms[a_]["priority"] != geoms[b_]["priority"]:
                        raise NotImplementedError("geom priority")
                    pairs.append((a_, b_))
    # MJX groups contacts by collision function: plane-box first, then box-box
    pairs.sort(key=lambda p: (geoms[p[0]]["type"], geoms[p[1]]["type"]))

    m = Model(nbody=nbody, njnt=njnt, nq=nq, nv=nv, nu=nu, ngeom=ngeom, nsite=len(sites),
              npair=len(pairs), neq=neq)
    m.timestep = float(opt.get("timestep", 0.002))
    m.gravity = _floats(opt.get("gravity", "0 0 -9.81"), 3)
    m.iterations = int(opt.get("iterations", 100))
    m.ls_iterations = int(opt.get("ls_iterations", 50))
    m.tolerance = float(opt.get("tolerance", 1e-8))
    m.ls_tolerance = float(opt.get("ls_tolerance", 0.01))
    m.impratio = float(opt.get("impratio", 1.0))
    m.integrator = opt.get("integrator", "Euler")
    if m.integrator != "implicitfast":
        raise NotImplementedError("only integrator=implicitfast (all Airbot models use it)")

    A = m.arrays
    A["body_parentid"] = body_parentid
    A["body_rootid"] = body_rootid
    A["body_weldid"] = body_weldid
    A["body_jntadr"] = body_jntadr
    A["body_jntnum"] = body_jntnum
    A["body_dofadr"] = body_dofadr
    A["body_dofnum"] = body_dofnum
    A["body_depth"] = body_depth
    A["body_pos"] = np.array([b["pos"] for b in bodies])
    A["body_quat"] = np.array([b["quat"] for b in bodies])
    A["body_ipos"] = body_ipos
    A["body_iquat"] = body_iquat
    A["body_mass"] = body_mass
    A["body_inertia"] = body_inertia
    A["jnt_type"] = np.array([j["type"] for j in joints], np.int32)
    A["jnt_qposadr"] = np.array(jnt_qposadr, np.int32)
    A["jnt_dofadr"] = np.array(jnt_dofadr, np.int32)
    A["jnt_bodyid"] = np.array([j["body"] for j in joints], np.int32)
    A["jnt_limited"] = np.array([j["limited"] for j in joints], np.int32)
    A["jnt_actfrclimited"] = np.array([j["actfrclimited"] for j in joints], np.int32)
    A["jnt_pos"] = np.array([j["pos"] for j in joints])
    A["jnt_axis"] = np.array([j["axis"] for j in joints])
    A["jnt_range"] = np.array([j["range"] for j in joints])
    A["jnt_actfrcrange"] = np.array([j["actfrcrange"] for j in joints])
    A["jnt_solref"] = np.array([j["solref_limit"] for j in joints])
    A["jnt_solimp"] = np.array([j["solimp_limit"] for j in joints])
    A["jnt_margin"] = np.array([j["margin"] for j in joints])
    A["qpos0"] = qpos0
    A["dof_bodyid"] = dof_bodyid
    A["dof_jntid"] = dof_jntid
    A["dof_parentid"] = dof_parentid
    A["dof_damping"] = np.array([joints[j]["damping"] for j in dof_jntid])
    A["dof_frictionloss"] = np.array([joints[j]["frictionloss"] for j in dof_jntid])
    A["dof_armature"] = np.array([joints[j]["armature"] for j in dof_jntid])
    A["dof_solref"] = np.array([joints[j]["solref_friction"] for j in dof_jntid])
    A["dof_solimp"] = np.array([joints[j]["solimp_friction"] for j in dof_jntid])
    A["geom_type"] = np.array([g["type"] for g in geoms], np.int32)
    A["geom_bodyid"] = np.array([g["body"] for g in geoms], np.int32)
    A["geom_contype"] = np.array([g["contype"] for g in geoms], np.int32)
    A["geom_conaffinity"] = np.array([g["conaffinity"] for g in geoms], np.int32)
    A["geom_condim"] = np.array([g["condim"] for g in geoms], np.int32)
    A["geom_priority"] = np.array([g["priority"] for g in geoms], np.int32)
    A["geom_pos"] = np.array([g["pos"] for g in geoms])
    A["geom_quat"] = np.array([g["quat"] for g in geoms])
    A["geom_size"] = np.array([g["size"] for g in geoms])
    A["geom_friction"] = np.array([g["friction"] for g in geoms])
    A["geom_solref"] = np.array([g["solref"] for g in geoms])
    A["geom_solimp"] = np.array([g["solimp"] for g in geoms])
    A["geom_solmix"] = np.array([g["solmix"] for g in geoms])
    A["geom_margin"] = np.array([g["margin"] for g in geoms])
    A["geom_gap"] = np.array([g["gap"] for g in geoms])
    A["site_bodyid"] = np.array([s["body"] for s in sites], np.int32)
    A["site_pos"] = np.array([s["pos"] for s in sites]).reshape(-1, 3)
    A["site_quat"] = np.array([s["quat"] for s in sites]).reshape(-1, 4)
    A["pair_geom1"] = np.array([p[0] for p in pairs], np.int32)
    A["pair_geom2"] = np.array([p[1] for p in pairs], np.int32)
    A["act_trnid"] = np.array([jname[a["joint"]] for a in acts], np.int32)
    A["act_ctrllimited"] = np.array([a["ctrllimited"] for a in acts], np.int32)
    A["act_forcelimited"] = np.array([a["forcelimited"] for a in acts], np.int32)
    A["act_gear"] = np.array([a["gear"] for a in acts])
    A["act_gainprm"] = np.array([[a["kp"], 0, 0] for a in acts], float).reshape(-1, 3)
    A["act_biasprm"] = np.array([[0, -a["kp"], -a["kv"]] for a in acts], float).reshape(-1, 3)
    A["act_ctrlrange"] = np.array([a["ctrlrange"] for a in acts]).reshape(-1, 2)
    A["act_forcerange"] = np.array([a["forcerange"] for a in acts]).reshape(-1, 2)
    A["eq_obj1id"] = np.array([e["j1"] for e in eqs], np.int32)
    A["eq_obj2id"] = np.array([e["j2"] for e in eqs], np.int32)
    A["eq_data"] = np.array([e["data"] for e in eqs]).reshape(-1, 5)
    A["eq_solref"] = np.array([e["solref"] for e in eqs]).reshape(-1, 2)
    A["eq_solimp"] = np.array([e["solimp"] for e in eqs]).reshape(-1, 5)
    m.names = dict(body=bname, geom=gname, site=sname, joint=jname,
                   actuator={a["name"]: i for i, a in enumerate(acts) if a["name"]})

    for ji, j in enumerate(joints):
        if any(abs(a["kv"]) > 0 for a in acts):
            raise NotImplementedError("actuator kv (velocity-dependent bias)")
    _set_const(m)
    return m


# ------------------------------------------------------------------ mj_setConst
def mass_matrix(m: Model, qpos: np.ndarray):
    """Joint-space inertia at ``qpos`` by composite-rigid-body summation
    (float64, written independently of the oracle: M = sum_b J_b^T I_b J_b)."""
    kin = kinematics(m, qpos)
    nv = m.nv
    M = np.zeros((nv, nv))
    for b in range(1, m.nbody):
        if m.body_mass[b] == 0 and not m.body_inertia[b].any():
            continue
        jp, jr = jacobian(m, kin, kin["xipos"][b], b)
        Iw = kin["ximat"][b] @ np.diag(m.body_inertia[b]) @ kin["ximat"][b].T
        M += m.body_mass[b] * jp.T @ jp + jr.T @ Iw @ jr
    M += np.diag(m.dof_armature)
    return M, kin


def kinematics(m: Model, qpos: np.ndarray):
    nb = m.nbody
    xpos = np.zeros((nb, 3))
    xquat = np.tile(np.array([1.0, 0, 0, 0]), (nb, 1))
    xanchor = np.zeros((m.njnt, 3))
    xaxis = np.zeros((m.njnt, 3))
    for b in range(1, nb):
        p = m.body_parentid[b]
        pos = xpos[p] + rotate(m.body_pos[b], xquat[p])
        quat = quat_mul(xquat[p], m.body_quat[b])
        for k in range(m.body_jntnum[b]):
            j = m.body_jntadr[b] + k
            qa = m.jnt_qposadr[j]
            t = m.jnt_type[j]
            if t == JNT_FREE:
                pos = qpos[qa:qa + 3].copy()
                quat = qpos[qa + 3:qa + 7] / np.linalg.norm(qpos[qa + 3:qa + 7])
                xanchor[j] = pos
                xaxis[j] = [0, 0, 1]
            else:
                anchor = rotate(m.jnt_pos[j], quat) + pos
                axis = rotate(m.jnt_axis[j], quat)
                xanchor[j], xaxis[j] = anchor, axis
                dq = qpos[qa] - m.qpos0[qa]
                if t == JNT_HINGE:
                    ql = np.concatenate([[math.cos(dq / 2)], math.sin(dq / 2) * m.jnt_axis[j]])
                    quat = quat_mul(quat, ql)
                    pos = anchor - rotate(m.jnt_pos[j], quat)
                else:
                    pos = pos + axis * dq
        xpos[b], xquat[b] = pos, quat / np.linalg.norm(quat)
    xmat = np.array([quat_to_mat(q) for q in xquat])
    xipos = np.array([xpos[b] + xmat[b] @ m.body_ipos[b] for b in range(nb)])
    ximat = np.array([quat_to_mat(quat_mul(xquat[b], m.body_iquat[b])) for b in range(nb)])
    return dict(xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, ximat=ximat,
                xanchor=xanchor, xaxis=xaxis)


def jacobian(m: Model, kin, point, body):
    """World-frame translational/rotational Jacobians (3,nv) of ``point`` on ``body``."""
    jp = np.zeros((3, m.nv))
    jr = np.zeros((3, m.nv))
    b = body
    while b > 0:
        for k in range(m.body_jntnum[b]):
            j = m.body_jntadr[b] + k
            d = m.jnt_dofadr[j]
            t = m.jnt_type[j]
            if t == JNT_FREE:
                jp[:, d:d + 3] = np.eye(3)
                R = kin["xmat"][b]
                for a in range(3):
                    ax = R[:, a]
                    jr[:, d + 3 + a] = ax
                    jp[:, d + 3 + a] = np.cross(ax, point - kin["xpos"][b])
            elif t == JNT_HINGE:
                jr[:, d] = kin["xaxis"][j]
                jp[:, d] = np.cross(kin["xaxis"][j], point - kin["xanchor"][j])
            else:
                jp[:, d] = kin["xaxis"][j]
        b = m.body_parentid[b]
    return jp, jr


def _set_const(m: Model):
    """dof_invweight0, body_invweight0, stat.meaninertia at qpos0
    (engine_setconst.c: set0 + setStat)."""
    M, kin = mass_matrix(m, m.qpos0)
    Minv = np.linalg.inv(M) if m.nv else np.zeros((0, 0))
    dinv = np.diag(Minv).copy()
    for j in range(m.njnt):
        if m.jnt_type[j] == JNT_FREE:
            d = m.jnt_dofadr[j]
            dinv[d:d + 3] = dinv[d:d + 3].mean()
            dinv[d + 3:d + 6] = dinv[d + 3:d + 6].mean()
    binv = np.zeros((m.nbody, 2))
    for b in range(1, m.nbody):
        if m.body_weldid[b] == 0:
            continue
        jp, jr = jacobian(m, kin, kin["xipos"][b], b)
        binv[b, 0] = np.trace(jp @ Minv @ jp.T) / 3
        binv[b, 1] = np.trace(jr @ Minv @ jr.T) / 3
    m.arrays["dof_invweight0"] = dinv
    m.arrays["body_invweight0"] = binv
    m.meaninertia = max(float(np.mean(np.diag(M))) if m.nv else 1.0, MJ_MINVAL)


# ------------------------------------------------------------- flattened writer
def write_flat_mjcf(src_path: str, dst_path: str, header: str = ""):
    """Re-emit a model as 'flattened' MJCF: every default resolved into explicit
    attributes, render-only content dropped.  Used by tools/make_assets.py."""
    root = ET.parse(src_path).getroot()
    classes = _read_defaults(root)
    keep_geom = ("name", "type", "pos", "quat", "euler", "size", "contype", "conaffinity", "condim",
                 "friction", "solref", "solimp", "solmix", "margin", "gap", "mass", "density", "group")
    keep_joint = ("name", "type", "pos", "axis", "range", "limited", "damping", "frictionloss",
                  "armature", "actuatorfrcrange", "margin")

    out = ET.Element("mujoco", {"model": root.get("model", "model")})
    for tag in ("compiler", "option"):
        merged = {}
        for el in root.findall(tag):
            merged.update(el.attrib)
        if merged:
            ET.SubElement(out, tag, dict(sorted(merged.items())))

    def res(tag, el, cc):
        cls = el.get("class") or cc or "main"
        a = dict(classes[cls].get(tag))
        a.update(el.attrib)
        a.pop("class", None)
        return a

    def emit(src, dst, cc):
        for ch in src:
            if ch.tag == "geom":
                a = res("geom", ch, cc)
                ET.SubElement(dst, "geom", {k: a[k] for k in sorted(a) if k in keep_geom})
            elif ch.tag == "joint":
                a = res("joint", ch, cc)
                ET.SubElement(dst, "joint", {k: a[k] for k in sorted(a) if k in keep_joint})
            elif ch.tag == "freejoint":
                ET.SubElement(dst, "freejoint", dict(ch.attrib))
            elif ch.tag == "site":
                a = res("site", ch, cc)
                ET.SubElement(dst, "site", {k: a[k] for k in sorted(a) if k in ("name", "pos", "quat", "euler")})
            elif ch.tag == "inertial":
                ET.SubElement(dst, "inertial", dict(sorted(ch.attrib.items())))
            elif ch.tag == "body":
                a = {k: v for k, v in sorted(ch.attrib.items()) if k in ("name", "pos", "quat", "euler")}
                emit(ch, ET.SubElement(dst, "body", a), ch.get("childclass", cc))

    wb = ET.SubElement(out, "worldbody")
    for sec in root.findall("worldbody"):
        emit(sec, wb, None)
    for tag in ("contact", "equality"):
        secs = root.findall(tag)
        if secs:
            d = ET.SubElement(out, tag)
            for sec in secs:
                for el in sec:
                    ET.SubElement(d, el.tag, {k: v.strip() for k, v in sorted(el.attrib.items())})
    secs = root.findall("actuator")
    if secs:
        d = ET.SubElement(out, "actuator")
        for sec in secs:
            for el in sec:
                a = res(el.tag, el, None)
                a.pop("user", None)
                ET.SubElement(d, el.tag, dict(sorted(a.items())))
    ET.indent(out, space=" ")
    txt = ET.tostring(out, encoding="unicode")
    with open(dst_path, "w") as f:
        if header:
            f.write("<!-- " + header + " -->\n")
        f.write(txt + "\n")
