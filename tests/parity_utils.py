"""Helpers shared by the parity tests: marshal between the device State buffers of
the CUDA env and the oracle's per-env structs, and compare them."""
from __future__ import annotations

import ctypes as C

import numpy as np

from oracle import oracle as O
from rsr_mjx_b200 import _lib

I = _lib.INFO


def gpu_to_oracle_states(env, buf_np):
    """buf_np: dict of numpy copies of the device buffers -> list[OrcEnvState]"""
    L, m = env.layout, env.model
    out = []
    for e in range(buf_np["data"].shape[0]):
        s = O.OrcEnvState()
        for dst, key in ((s.d, "data"), (s.first, "first_data")):
            row = buf_np[key][e].astype(np.float64)
            dst.qpos[:m.nq] = list(row[L.qpos:L.qpos + m.nq])
            dst.qvel[:m.nv] = list(row[L.qvel:L.qvel + m.nv])
            dst.ctrl[:m.nu] = list(row[L.ctrl:L.ctrl + m.nu])
            dst.qacc_warmstart[:m.nv] = list(row[L.qacc_warmstart:L.qacc_warmstart + m.nv])
            dst.time = float(row[L.time])
            np.ctypeslib.as_array(dst.xpos)[:m.nbody] = row[L.xpos:L.xpos + 3 * m.nbody].reshape(-1, 3)
            np.ctypeslib.as_array(dst.xquat)[:m.nbody] = row[L.xquat:L.xquat + 4 * m.nbody].reshape(-1, 4)
            np.ctypeslib.as_array(dst.site_xpos)[:m.nsite] = row[L.site_xpos:L.site_xpos + 3 * m.nsite].reshape(-1, 3)
            np.ctypeslib.as_array(dst.geom_xpos)[:m.ngeom] = row[L.geom_xpos:L.geom_xpos + 3 * m.ngeom].reshape(-1, 3)
        s.obs[:L.obs_stride] = list(buf_np["obs"][e].astype(np.float64))
        s.first_obs[:L.obs_stride] = list(buf_np["first_obs"][e].astype(np.float64))
        s.reward = float(buf_np["reward"][e])
        s.done = float(buf_np["done"][e])
        info = buf_np["info"][e].astype(np.float64)
        s.steps = info[I["STEPS"]]
        s.truncation = info[I["TRUNCATION"]]
        s.target_pos[:] = list(info[I["TARGET"]:I["TARGET"] + 3])
        s.target2_pos[:] = list(info[I["TARGET2"]:I["TARGET2"] + 3])
        s.new_pos[:] = list(info[I["NEWPOS"]:I["NEWPOS"] + 2])
        s.site_pos[:] = list(info[I["SITE"]:I["SITE"] + 3])
        s.obj_pos[:] = list(info[I["OBJ"]:I["OBJ"] + 3])
        s.last_action = info[I["LAST_ACTION"]]
        s.xita = info[I["XITA"]]
        s.target_w = info[I["TARGET_W"]]
        s.metrics[:5] = list(buf_np["metrics"][e][:5].astype(np.float64))
        out.append(s)
    return out


def buffers_to_numpy(state):
    return {k: v.detach().cpu().numpy().copy() for k, v in state._buf.items()}


def oracle_row(env, d):
    """OrcData -> float64 data row in the device layout"""
    L, m = env.layout, env.model
    row = np.zeros(L.data_stride)
    row[L.qpos:L.qpos + m.nq] = np.array(d.qpos)[:m.nq]
    row[L.qvel:L.qvel + m.nv] = np.array(d.qvel)[:m.nv]
    row[L.ctrl:L.ctrl + m.nu] = np.array(d.ctrl)[:m.nu]
    row[L.qacc_warmstart:L.qacc_warmstart + m.nv] = np.array(d.qacc_warmstart)[:m.nv]
    row[L.time] = d.time
    row[L.xpos:L.xpos + 3 * m.nbody] = np.ctypeslib.as_array(d.xpos)[:m.nbody].ravel()
    row[L.xquat:L.xquat + 4 * m.nbody] = np.ctypeslib.as_array(d.xquat)[:m.nbody].ravel()
    row[L.site_xpos:L.site_xpos + 3 * m.nsite] = np.ctypeslib.as_array(d.site_xpos)[:m.nsite].ravel()
    row[L.geom_xpos:L.geom_xpos + 3 * m.ngeom] = np.ctypeslib.as_array(d.geom_xpos)[:m.ngeom].ravel()
    return row


def oracle_info(s):
    info = np.zeros(_lib.INFO_STRIDE)
    info[I["TARGET"]:I["TARGET"] + 3] = list(s.target_pos)
    info[I["TARGET2"]:I["TARGET2"] + 3] = list(s.target2_pos)
    info[I["NEWPOS"]:I["NEWPOS"] + 2] = list(s.new_pos)
    info[I["SITE"]:I["SITE"] + 3] = list(s.site_pos)
    info[I["OBJ"]:I["OBJ"] + 3] = list(s.obj_pos)
    info[I["LAST_ACTION"]] = s.last_action
    info[I["XITA"]] = s.xita
    info[I["TARGET_W"]] = s.target_w
    info[I["STEPS"]] = s.steps
    info[I["TRUNCATION"]] = s.truncation
    return info


def copy_state(s):
    t = O.OrcEnvState()
    C.memmove(C.byref(t), C.byref(s), C.sizeof(s))
    return t


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(floor, |b|_inf-per-field)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(floor, float(np.max(np.abs(b))) if b.size else 0.0))


# ---- vectorised marshalling (long teacher-forced runs): numpy structured views over ctypes arrays of OrcEnvState
def oracle_state_array(n):
    """(ctypes array of n OrcEnvState, numpy structured view onto the same memory)"""
    arr = (O.OrcEnvState * n)()
    return arr, np.ctypeslib.as_array(arr)


def fill_oracle_states(env, view, buf_np, idx=None):
    """Vectorised gpu_to_oracle_states: writes the device buffers (numpy copies, rows `idx`) into the structured view."""
    L, m = env.layout, env.model
    sel = slice(None) if idx is None else idx
    for dst, key in ((view["d"], "data"), (view["first"], "first_data")):
        row = buf_np[key][sel].astype(np.float64)
        dst["qpos"][:, :m.nq] = row[:, L.qpos:L.qpos + m.nq]
        dst["qvel"][:, :m.nv] = row[:, L.qvel:L.qvel + m.nv]
        dst["ctrl"][:, :m.nu] = row[:, L.ctrl:L.ctrl + m.nu]
        dst["qacc_warmstart"][:, :m.nv] = row[:, L.qacc_warmstart:L.qacc_warmstart + m.nv]
        dst["time"][:] = row[:, L.time]
        dst["xpos"][:, :m.nbody] = row[:, L.xpos:L.xpos + 3 * m.nbody].reshape(-1, m.nbody, 3)
        dst["xquat"][:, :m.nbody] = row[:, L.xquat:L.xquat + 4 * m.nbody].reshape(-1, m.nbody, 4)
        dst["site_xpos"][:, :m.nsite] = row[:, L.site_xpos:L.site_xpos + 3 * m.nsite].reshape(-1, m.nsite, 3)
        dst["geom_xpos"][:, :m.ngeom] = row[:, L.geom_xpos:L.geom_xpos + 3 * m.ngeom].reshape(-1, m.ngeom, 3)
    view["obs"][:, :L.obs_stride] = buf_np["obs"][sel]
    view["first_obs"][:, :L.obs_stride] = buf_np["first_obs"][sel]
    view["reward"][:] = buf_np["reward"][sel]
    view["done"][:] = buf_np["done"][sel]
    info = buf_np["info"][sel].astype(np.float64)
    view["steps"][:] = info[:, I["STEPS"]]
    view["truncation"][:] = info[:, I["TRUNCATION"]]
    view["target_pos"][:] = info[:, I["TARGET"]:I["TARGET"] + 3]
    view["target2_pos"][:] = info[:, I["TARGET2"]:I["TARGET2"] + 3]
    view["new_pos"][:] = info[:, I["NEWPOS"]:I["NEWPOS"] + 2]
    view["site_pos"][:] = info[:, I["SITE"]:I["SITE"] + 3]
    view["obj_pos"][:] = info[:, I["OBJ"]:I["OBJ"] + 3]
    view["last_action"][:] = info[:, I["LAST_ACTION"]]
    view["xita"][:] = info[:, I["XITA"]]
    view["target_w"][:] = info[:, I["TARGET_W"]]
    view["metrics"][:, :5] = buf_np["metrics"][sel][:, :5]


def oracle_states_to_buffers(env, view):
    """The inverse: dict of float64 arrays in the device layout (data rows, obs, reward, done, info, metrics)."""
    L, m = env.layout, env.model
    n = view.shape[0]
    d = view["d"]
    data = np.zeros((n, L.data_stride))
    data[:, L.qpos:L.qpos + m.nq] = d["qpos"][:, :m.nq]
    data[:, L.qvel:L.qvel + m.nv] = d["qvel"][:, :m.nv]
    data[:, L.ctrl:L.ctrl + m.nu] = d["ctrl"][:, :m.nu]
    data[:, L.qacc_warmstart:L.qacc_warmstart + m.nv] = d["qacc_warmstart"][:, :m.nv]
    data[:, L.time] = d["time"]
    data[:, L.xpos:L.xpos + 3 * m.nbody] = d["xpos"][:, :m.nbody].reshape(n, -1)
    data[:, L.xquat:L.xquat + 4 * m.nbody] = d["xquat"][:, :m.nbody].reshape(n, -1)
    data[:, L.site_xpos:L.site_xpos + 3 * m.nsite] = d["site_xpos"][:, :m.nsite].reshape(n, -1)
    data[:, L.geom_xpos:L.geom_xpos + 3 * m.ngeom] = d["geom_xpos"][:, :m.ngeom].reshape(n, -1)
    info = np.zeros((n, _lib.INFO_STRIDE))
    info[:, I["TARGET"]:I["TARGET"] + 3] = view["target_pos"]
    info[:, I["TARGET2"]:I["TARGET2"] + 3] = view["target2_pos"]
    info[:, I["NEWPOS"]:I["NEWPOS"] + 2] = view["new_pos"]
    info[:, I["SITE"]:I["SITE"] + 3] = view["site_pos"]
    info[:, I["OBJ"]:I["OBJ"] + 3] = view["obj_pos"]
    info[:, I["LAST_ACTION"]] = view["last_action"]
    info[:, I["XITA"]] = view["xita"]
    info[:, I["TARGET_W"]] = view["target_w"]
    info[:, I["STEPS"]] = view["steps"]
    info[:, I["TRUNCATION"]] = view["truncation"]
    return dict(data=data, obs=view["obs"][:, :L.obs_stride].copy(), reward=view["reward"].copy(),
                done=view["done"].copy(), info=info, metrics=view["metrics"].copy())


def elem_err(a, b, floor=1.0):
    """per-element mixed error |a - b| / max(floor, |b|): the worst element (1e-4 relative for |x| >= 1, 1e-4 absolute
    below — the np.allclose(rtol = atol) form of the north star's "1e-4 relative in fp32")"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(floor, np.abs(b)))) if a.size else 0.0


def elem_err_rows(a, b, floor=1.0):
    """elem_err per leading row"""
    a = np.asarray(a, np.float64).reshape(len(a), -1)
    b = np.asarray(b, np.float64).reshape(len(b), -1)
    return np.max(np.abs(a - b) / np.maximum(floor, np.abs(b)), axis=1)


def oracle_step_batch(blob_or_blobs, cfg, arr, actions, precision="f32"):
    """one env.step of every state in the ctypes array (OpenMP over envs inside the oracle); actions [n, nu]"""
    return O.rollout(blob_or_blobs, cfg, arr, np.asarray(actions, np.float64)[None], precision=precision)
