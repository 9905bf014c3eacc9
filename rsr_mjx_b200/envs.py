"""brax-shaped `Env.reset / step / State` API over the fused CUDA stepper.

`AirbotPlayBase` mirrors the three reference env classes
  test/airbot.py::AirbotPlayBase                       -> kind='sf'
  ppo_train/airbot_training/cube_env.py::AirbotPlayBase -> kind='cube'
  ppo_train/airbot_training/T_shape_env.py::AirbotPlayBase -> kind='T'
and *is already* the training stack brax builds around them —
AutoResetWrapper(EpisodeWrapper(VmapWrapper | DomainRandomizationVmapWrapper(env)))
(`brax.envs.training.wrap`, called at reference RSR/train.py:220-229) — batched
over `num_envs`: `reset(rng[N,2]) -> State`, `step(state, action[N,5]) -> State`.
All State leaves are torch CUDA tensors with a leading N axis, carrying the same
names the reference's callers read.  One `step` = one kernel
launch.  The State returned by `step` aliases the env's device buffers
(equivalent to jax buffer donation); `state.clone()` detaches a copy.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Any, Callable, Dict, Optional

import numpy as np
import torch

from . import _lib, airbot_spec
from .mjcf import Model
from .model import pack_model


@dataclasses.dataclass
class PipelineState:
    """The mjx.Data subset the reference's callers read; all are views of `data`."""
    data: torch.Tensor  # [N, data_stride]
    layout: Any

    def _v(self, off, n, shape=None):
        t = self.data[:, off:off + n]
        return t if shape is None else t.unflatten(1, shape)

    @property
    def qpos(self):
        return self._v(self.layout.qpos, self.layout.nq)

    @property
    def qvel(self):
        return self._v(self.layout.qvel, self.layout.nv)

    @property
    def ctrl(self):
        return self._v(self.layout.ctrl, self.layout.nu)

    @property
    def qacc_warmstart(self):
        return self._v(self.layout.qacc_warmstart, self.layout.nv)

    @property
    def time(self):
        return self.data[:, self.layout.time]

    @property
    def xpos(self):
        return self._v(self.layout.xpos, self.layout.nbody * 3, (self.layout.nbody, 3))

    @property
    def xquat(self):
        return self._v(self.layout.xquat, self.layout.nbody * 4, (self.layout.nbody, 4))

    @property
    def site_xpos(self):
        return self._v(self.layout.site_xpos, self.layout.nsite * 3, (self.layout.nsite, 3))

    @property
    def geom_xpos(self):
        return self._v(self.layout.geom_xpos, self.layout.ngeom * 3, (self.layout.ngeom, 3))

    # brax aliases
    @property
    def q(self):
        return self.qpos

    @property
    def qd(self):
        return self.qvel


@dataclasses.dataclass
class State:
    """brax.envs.base.State with a leading env axis on every leaf."""
    pipeline_state: PipelineState
    obs: torch.Tensor
    reward: torch.Tensor
    done: torch.Tensor
    metrics: Dict[str, torch.Tensor]
    info: Dict[str, Any]
    # raw device buffers behind the views (what the C-ABI steps in place)
    _buf: Dict[str, torch.Tensor] = dataclasses.field(default_factory=dict, repr=False)

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)

    def clone(self) -> "State":
        buf = {k: v.clone() for k, v in self._buf.items()}
        return _state_from_buffers(buf, self.pipeline_state.layout, self._kind)

    _kind: str = "sf"


def _state_from_buffers(buf, layout, kind) -> State:
    I = _lib.INFO
    info_t = buf["info"]
    info: Dict[str, Any] = {}
    if kind == "T":
        info["target_base_pos"] = info_t[:, I["TARGET"]:I["TARGET"] + 3]
        info["target_vertical_pos"] = info_t[:, I["TARGET2"]:I["TARGET2"] + 3]
        info["target_w"] = info_t[:, I["TARGET_W"]]
        info["new_T_pos"] = info_t[:, I["NEWPOS"]:I["NEWPOS"] + 2]
        info["site_pos"] = info_t[:, I["SITE"]:I["SITE"] + 3]
        info["T_pos"] = info_t[:, I["OBJ"]:I["OBJ"] + 3]
        info["xita"] = info_t[:, I["XITA"]]
    else:
        info["target_pos"] = info_t[:, I["TARGET"]:I["TARGET"] + 3]
        info["new_cube_pos"] = info_t[:, I["NEWPOS"]:I["NEWPOS"] + 2]
        info["site_pos"] = info_t[:, I["SITE"]:I["SITE"] + 3]
        info["cube_pos"] = info_t[:, I["OBJ"]:I["OBJ"] + 3]
        info["reached_box"] = torch.zeros_like(info_t[:, 0])
        if kind == "sf":
            info["last_action"] = info_t[:, I["LAST_ACTION"]]
    info["steps"] = info_t[:, I["STEPS"]]
    info["truncation"] = info_t[:, I["TRUNCATION"]]
    info["first_pipeline_state"] = PipelineState(buf["first_data"], layout)
    info["first_obs"] = buf["first_obs"][:, :layout.obs_size]
    metrics = {k: buf["metrics"][:, i] for i, k in enumerate(airbot_spec.METRIC_KEYS[kind])}
    st = State(pipeline_state=PipelineState(buf["data"], layout), obs=buf["obs"][:, :layout.obs_size],
               reward=buf["reward"], done=buf["done"], metrics=metrics, info=info, _buf=buf)
    st._kind = kind
    return st


class System:
    """Minimal stand-in for brax's `System`: the compiled model with the leaves
    callers touch (`geom_friction`, `body_mass`, `dof_damping`, `dof_frictionloss`,
    `qpos0`, `nq/nv/nu`, `opt.timestep`) and `tree_replace` for the four arrays
    the reference randomises (domain_randomize.py:85-90, rsr_pipeline.py:103-106)."""
    _PER_ENV = ("geom_friction", "body_mass", "dof_damping", "dof_frictionloss")

    def __init__(self, model: Model, overrides: Optional[Dict[str, torch.Tensor]] = None):
        self.mj_model = model
        self._ov = dict(overrides or {})

    def __getattr__(self, k):
        if k in ("mj_model", "_ov"):
            raise AttributeError(k)
        if k in self._ov:
            return self._ov[k]
        return getattr(self.mj_model, k)

    def tree_replace(self, params: Dict[str, Any]) -> "System":
        ov = dict(self._ov)
        for k, v in params.items():
            if k not in self._PER_ENV:
                raise NotImplementedError(f"tree_replace({k!r}): only {self._PER_ENV} can be replaced")
            ov[k] = v
        return System(self.mj_model, ov)


class AirbotPlayBase:
    """Batched, wrapped Airbot env (see module docstring).

    Args mirror the reference constructors (reward weights, reset ranges, ...)
    plus the wrapper arguments of `brax.envs.training.wrap`:
      num_envs, episode_length, action_repeat, randomization_fn.
    `randomization_fn(sys, rng) -> (sys_batched, in_axes)` follows the reference
    contract (domain_randomize.py:26); see rsr_mjx_b200.domain_randomize.
    """

    def __init__(self, kind: str = "sf", num_envs: int = 1, episode_length: int = 1000, action_repeat: int = 1,
                 model_path: Optional[str] = None, device: str | torch.device = "cuda",
                 randomization_fn: Optional[Callable] = None, randomization_rng=None, **kwargs):
        if kind not in airbot_spec.KINDS:
            raise ValueError(f"kind must be one of {list(airbot_spec.KINDS)}")
        if action_repeat != 1:
            raise NotImplementedError("action_repeat != 1 (the reference never uses it)")
        self.kind = kind
        self.num_envs = int(num_envs)
        self.episode_length = int(episode_length)
        self.action_repeat = int(action_repeat)
        self._params = dict(kwargs)
        self._model_path = model_path
        self._randomization_fn = randomization_fn
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("AirbotPlayBase runs only on a CUDA device (no CPU fallback)")
        self.model = airbot_spec.load_model(kind, model_path)
        self.sys = System(self.model)
        self._ids = airbot_spec.env_ids(self.model, kind)
        for k, v in self._ids.items():
            setattr(self, k, v)
        self.cfg = airbot_spec.make_env_cfg(self.model, kind, self._params, episode_length, action_repeat)
        self._n_frames = self.cfg.n_frames
        L = _lib.lib()
        self._blob = pack_model(self.model)
        handle = C.c_void_p()
        _lib.check(L.rsrx_model_create(C.byref(self._blob), C.sizeof(self._blob), C.byref(self.cfg), C.byref(handle)),
                   "rsrx_model_create")
        self._handle = handle
        self.layout = _lib.Layout()
        _lib.check(L.rsrx_model_layout(self._handle, C.byref(self.layout)))
        self._per_env = _lib.PerEnvC()
        self._per_env_tensors: Dict[str, torch.Tensor] = {}
        if randomization_fn is not None:
            if randomization_rng is None:
                raise ValueError("randomization_fn needs randomization_rng (keys [num_envs, 2])")
            self.randomize(randomization_fn, randomization_rng)
        self._lowers = torch.tensor(self.model.act_ctrlrange[:, 0], dtype=torch.float32, device=self.device)
        self._uppers = torch.tensor(self.model.act_ctrlrange[:, 1], dtype=torch.float32, device=self.device)

    def randomize(self, randomization_fn: Callable, rng) -> None:
        """DomainRandomizationVmapWrapper (wrapper.py:139-165): `randomization_fn(sys, rng[N, 2]) -> (sys_v, in_axes)`
        installs the batched model leaves of `sys_v` on this env."""
        sys_v, _ = randomization_fn(self.sys, rng)
        self.set_per_env(**{k: sys_v._ov[k] for k in System._PER_ENV if k in sys_v._ov})
        self._randomization_fn = randomization_fn

    def clone(self, num_envs: int, randomization_fn: Optional[Callable] = None, randomization_rng=None) -> "AirbotPlayBase":
        """A second env of the same kind / reward parameters / wrapper settings with its own batch size (the eval env
        of RSR/train.py:428-439 is the training env re-wrapped with `num_eval_envs` randomisation keys)."""
        return AirbotPlayBase(self.kind, num_envs=num_envs, episode_length=self.episode_length,
                              action_repeat=self.action_repeat, model_path=self._model_path, device=self.device,
                              randomization_fn=randomization_fn, randomization_rng=randomization_rng, **self._params)

    # ------------------------------------------------------------------ properties
    @property
    def observation_size(self) -> int:
        return airbot_spec.OBS_SIZE[self.kind]

    @property
    def action_size(self) -> int:
        return self.model.nu

    @property
    def dt(self) -> float:
        return self.model.timestep * self._n_frames

    @property
    def unwrapped(self):
        return self

    @property
    def backend(self) -> str:
        return "rsrx-sm100a"

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().rsrx_model_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # --------------------------------------------------------------- per-env model
    def set_per_env(self, **arrays):
        """Per-env model leaves [N, ...] (geom_friction[N,ngeom,3], body_mass[N,nbody],
        dof_damping[N,nv], dof_frictionloss[N,nv]); None restores the nominal value."""
        shapes = dict(geom_friction=(self.model.ngeom, 3), body_mass=(self.model.nbody,),
                      dof_damping=(self.model.nv,), dof_frictionloss=(self.model.nv,))
        for k, v in arrays.items():
            if k not in shapes:
                raise KeyError(k)
            if v is None:
                self._per_env_tensors.pop(k, None)
                setattr(self._per_env, k, None)
                continue
            t = torch.as_tensor(v, dtype=torch.float32, device=self.device).contiguous()
            if tuple(t.shape) != (self.num_envs, *shapes[k]):
                raise ValueError(f"{k}: expected shape {(self.num_envs, *shapes[k])}, got {tuple(t.shape)}")
            self._per_env_tensors[k] = t
            setattr(self._per_env, k, t.data_ptr())

    # ------------------------------------------------------------------ reset/step
    def _alloc(self) -> Dict[str, torch.Tensor]:
        N, L, dev = self.num_envs, self.layout, self.device
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
        return dict(data=z(N, L.data_stride), first_data=z(N, L.data_stride), obs=z(N, L.obs_stride),
                    first_obs=z(N, L.obs_stride), reward=z(N), done=z(N), info=z(N, L.info_stride),
                    metrics=z(N, L.metrics_stride), status=z(N, dt=torch.int32))

    @staticmethod
    def _cstate(buf) -> _lib.StateC:
        return _lib.StateC(*[buf[k].data_ptr() for k in ("data", "first_data", "obs", "first_obs", "reward", "done",
                                                         "info", "metrics", "status")])

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, rng) -> State:
        """rng: jax-style keys, uint32 [N, 2] (numpy or torch).  Sampling follows
        the reference's reset with a NumPy restatement of jax.random (host side)."""
        keys = rng.cpu().numpy() if isinstance(rng, torch.Tensor) else np.asarray(rng)
        keys = keys.astype(np.uint32).reshape(-1, 2)
        if keys.shape[0] != self.num_envs:
            raise ValueError(f"reset expects {self.num_envs} keys, got {keys.shape[0]}")
        qpos, qvel, ctrl = airbot_spec.sample_reset(self.model, self.kind, keys, self._params)
        return self.reset_from(qpos, qvel, ctrl)

    def reset_from(self, qpos, qvel, ctrl) -> State:
        """pipeline_init + env bookkeeping from explicit initial conditions."""
        dev = self.device
        q = torch.as_tensor(qpos, dtype=torch.float32, device=dev).contiguous()
        v = torch.as_tensor(qvel, dtype=torch.float32, device=dev).contiguous()
        c = torch.as_tensor(ctrl, dtype=torch.float32, device=dev).contiguous()
        N = self.num_envs
        if q.shape != (N, self.model.nq) or v.shape != (N, self.model.nv) or c.shape != (N, self.model.nu):
            raise ValueError("reset_from: qpos/qvel/ctrl must be [N,nq], [N,nv], [N,nu]")
        buf = self._alloc()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().rsrx_env_reset(self._handle, N, q.data_ptr(), v.data_ptr(), c.data_ptr(),
                                                 C.byref(self._per_env), self._cstate(buf), self._stream()),
                       "rsrx_env_reset")
        self._keep = (q, v, c)
        return _state_from_buffers(buf, self.layout, self.kind)

    def step(self, state: State, action) -> State:
        a = torch.as_tensor(action, dtype=torch.float32, device=self.device)
        if a.shape != (self.num_envs, self.model.nu):
            raise ValueError(f"action must be [{self.num_envs}, {self.model.nu}], got {tuple(a.shape)}")
        a = a.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rsrx_env_step(self._handle, self.num_envs, self._cstate(state._buf), a.data_ptr(),
                                                C.byref(self._per_env), self._stream()), "rsrx_env_step")
        return state

    def step_host(self, state: State, host_action: torch.Tensor, host_obs: Optional[torch.Tensor] = None,
                  host_reward: Optional[torch.Tensor] = None, host_done: Optional[torch.Tensor] = None) -> State:
        """`step` for a host-side caller: `host_action` [N, nu] (CPU float32, ideally pinned) goes in, the optional CPU
        buffers `host_obs` [N, obs_stride], `host_reward` [N], `host_done` [N] receive the results; copies and launch are
        queued by one C call (`rsrx_env_step_host`).  Synchronise the stream before reading the host buffers."""
        N = self.num_envs
        def chk(t, shape, name):
            if t is None:
                return None
            if t.device.type != "cpu" or t.dtype != torch.float32 or tuple(t.shape) != shape or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous CPU float32 tensor of shape {shape}")
            return t.data_ptr()
        pa = chk(host_action, (N, self.model.nu), "host_action")
        if pa is None:
            raise ValueError("host_action is required")
        if getattr(self, "_act_staging", None) is None:
            self._act_staging = torch.empty(N, self.model.nu, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rsrx_env_step_host(
                self._handle, N, self._cstate(state._buf), pa, self._act_staging.data_ptr(),
                chk(host_obs, (N, self.layout.obs_stride), "host_obs"), chk(host_reward, (N,), "host_reward"),
                chk(host_done, (N,), "host_done"), C.byref(self._per_env), self._stream()), "rsrx_env_step_host")
        return state

    def step_raw(self, buf: Dict[str, torch.Tensor], action_ptr: int):
        """Launch-only path for benchmarks / CUDA-graph capture (no tensor checks)."""
        _lib.check(_lib.lib().rsrx_env_step(self._handle, self.num_envs, self._cstate(buf), action_ptr,
                                            C.byref(self._per_env), self._stream()), "rsrx_env_step")

    # pipeline_step / mjx.step on raw data rows (used by tests and the friction sweep)
    def physics_step(self, data: torch.Tensor, nsteps: int = 1, status: Optional[torch.Tensor] = None):
        N = data.shape[0]
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rsrx_physics_step(self._handle, N, data.data_ptr(), int(nsteps), C.byref(self._per_env),
                                                    status.data_ptr() if status is not None else None, self._stream()),
                       "rsrx_physics_step")
        return data

    def physics_step_debug(self, data: torch.Tensor) -> torch.Tensor:
        N = data.shape[0]
        dump = torch.zeros(N, _lib.lib().rsrx_debug_stride(), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rsrx_physics_step_debug(self._handle, N, data.data_ptr(), C.byref(self._per_env),
                                                          dump.data_ptr(), self._stream()), "rsrx_physics_step_debug")
        return dump

    def status(self, state: State) -> torch.Tensor:
        return state._buf["status"]

    def render(self, *a, **k):
        raise NotImplementedError("rendering is out of scope (SURVEY.md §2: render-only paths)")
