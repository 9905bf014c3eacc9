#!/usr/bin/env python
"""bench.py — Airbot cube env-steps/sec (BASELINE.json metric) on N B200s.

One "step" = one `env.step` of the wrapped, batched Airbot cube env
(test/airbot.py + sf.xml: action shaping, 4 x mjx.step, reward/obs/done,
episode + auto-reset) over ENVS_PER_GPU environments on every rank — one kernel
launch per rank.  Envs shard across GPUs with no data-path collective (weak
scaling: per-GPU work is fixed).

  python bench.py --gpus 1 --steps K --warmup W            # our arm
  python bench.py --impl reference ...                      # CPU arm: the oracle port, all host threads
  torchrun --nproc-per-node N bench.py --gpus N ...         # N > 1

Prints ONE JSON line (rank 0).  Timing: W >= 3 warm-up steps, then K steps each
bracketed by CUDA events on the launching stream with a 256 MiB L2 flush between
timed steps (outside the events); barrier + synchronize on both sides; max over
ranks.  `value` = whole-job env-steps/s with state and actions resident in HBM;
`e2e` = the same through the public API with HOST action/observation buffers
(pinned H2D of actions + D2H of obs/reward/done inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

ALGO_BYTES = {"sf": 1232, "cube": 1232, "T": 580}  # algorithmic bytes per env-step
METRIC = "airbot_cube_env_steps_per_sec"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash():
    """sha256 over the CUDA sources of the stepper: the committed ncu summary is only quoted for the build it measured"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "rsr_mjx_b200", "csrc")
    for f in sorted(os.listdir(d)):
        # (rsrx_api.cu holds every trainer entry point too and changes with them; the launch geometry it chooses is in the
        # summary's .meta.json)
        if f in ("rsrx_device.cuh", "rsrx_physics.cuh", "rsrx_env.cuh", "rsrx_redo.cu", "rsrx_redo.h", "rsrx_mid.cu", "rsrx_mid.h"):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


NCU_SUMMARY = os.path.join(ROOT, "profiles", "r2_step_kernel_ncu_full_summary.csv")


def read_ncu_summary(kind, n_envs):
    """The committed `ncu --set full` summary of one step_kernel launch (tools/ncu_summary.py writes it together with the
    hash of the kernel sources it was captured from).  Returns None when there is none for this workload, and
    {'stale': True} when the kernel sources have changed since — stale counters are never quoted."""
    meta_p = NCU_SUMMARY.replace(".csv", ".meta.json")
    if not (os.path.exists(NCU_SUMMARY) and os.path.exists(meta_p)):
        return None
    with open(meta_p) as f:
        meta = json.load(f)
    if meta.get("kind") != kind or meta.get("envs") != n_envs:
        return None
    if meta.get("source_hash") != kernel_source_hash():
        return {"stale": True, "captured_from_source_hash": meta.get("source_hash"), "current_source_hash": kernel_source_hash()}
    raw = {}
    with open(NCU_SUMMARY) as f:
        for ln in f:
            parts = ln.strip().split(",")
            if len(parts) == 3 and parts[0] != "metric":
                scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(parts[1], 1.0)  # -> bytes, ms
                try:
                    raw[parts[0]] = float(parts[2]) * scale
                except ValueError:
                    pass
    return {"stale": False, "raw": raw, "meta": meta}


def compute_view(summary, n_envs, kern_ms, sm_mhz, n_sms=148):
    """the compute-side view of step_kernel — what actually bounds an issue/latency-bound kernel.
    Counter values come from the committed ncu capture of THIS build (hash-checked); the issue-slot fraction is
    re-derived from this run's own kernel time: warp-instructions per launch (a property of the workload, stable to
    ~1 % between launches of the same distribution) / (launch time x SM clock x SMs x 4 schedulers)."""
    if summary is None:
        return None
    if summary["stale"]:
        return {"stale": True, "note": "kernel sources changed since the committed ncu capture; re-run tools/ncu_summary.py",
                **{k: v for k, v in summary.items() if k != "stale"}}
    raw = summary["raw"]
    want = {"sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fp32_fma_pipe_pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
            "sm__inst_executed.avg.per_cycle_active": "ipc_per_sm",
            "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
            "smsp__inst_issued.sum": "warp_instructions_per_launch",
            "gpu__time_duration.sum": "ncu_launch_ms"}
    out = {"stale": False, "source": "profiles/" + os.path.basename(NCU_SUMMARY) + " (ncu --set full, one launch)",
           "source_hash": summary["meta"]["source_hash"]}
    for k, name in want.items():
        if k in raw:
            out[name] = raw[k]
    try:
        per_cycle = (raw["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed"]
                     + raw["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed"]
                     + 2.0 * raw["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed"])
        flop = per_cycle * raw["sm__cycles_elapsed.max"]
        out["fp32_flop_per_env_step"] = flop / n_envs
        out["fp32_tflops_peak_at_run_clock"] = n_sms * 128 * 2 * sm_mhz * 1e6 / 1e12 if sm_mhz else None
        out["fp32_tflops_achieved_this_run"] = flop / (kern_ms * 1e-3) / 1e12
    except KeyError:
        pass
    if "smsp__inst_issued.sum" in raw and sm_mhz:
        out["issue_slot_frac_this_run"] = raw["smsp__inst_issued.sum"] / (kern_ms * 1e-3 * sm_mhz * 1e6 * n_sms * 4)
    return out


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_oracle_rate(kind, n_envs, n_steps, dense, threads=0, precision="f32", stationary=True):
    """env-steps/s of the CPU oracle (OpenMP over envs) on a bounded sample of the same workload.  stationary: like the
    GPU arm, the envs are first advanced (untimed) by one episode with staggered step counters, so that the timed steps
    see every episode phase (contact load) in the proportion a running job does."""
    import ctypes as C
    from oracle import oracle as O
    from rsr_mjx_b200 import airbot_spec as A, prng
    from rsr_mjx_b200.model import pack_model
    O.build()
    if threads <= 0:
        threads = os.cpu_count() or 1  # all host threads (torchrun exports OMP_NUM_THREADS=1; do not inherit that)
    m = A.load_model(kind)
    blob, cfg = pack_model(m), A.make_env_cfg(m, kind, episode_length=1200)
    keys = prng.split(prng.PRNGKey(0), n_envs)
    q, v, c = A.sample_reset(m, kind, keys)
    states = (O.OrcEnvState * n_envs)()
    for i in range(n_envs):
        s = O.env_reset(blob, cfg, q[i], v[i], c[i], precision=precision)
        C.memmove(C.byref(states[i]), C.byref(s), C.sizeof(s))
    actions = np.random.default_rng(1).uniform(-1, 1, (n_steps, n_envs, m.nu))
    if stationary:
        view = np.ctypeslib.as_array(states)
        view["steps"][:] = np.random.default_rng(2).integers(0, 1200, n_envs).astype(np.float64)
        pre = np.random.default_rng(3).uniform(-1, 1, (64, n_envs, m.nu))
        for c in range(0, 1200, 64):  # active-set rows for the untimed advance, whatever the timed variant
            O.rollout(blob, cfg, states, pre[:min(64, 1200 - c)], precision=precision, dense=False, nthreads=threads)
    O.rollout(blob, cfg, states, actions[:2], precision=precision, dense=dense, nthreads=threads)  # warm-up
    t0 = time.perf_counter()
    O.rollout(blob, cfg, states, actions, precision=precision, dense=dense, nthreads=threads)
    dt = time.perf_counter() - t0
    return n_envs * n_steps / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count()
    n_envs, n_steps = args.cpu_envs, max(args.steps, 1)
    # warm-up steps are part of cpu_oracle_rate (2 untimed steps); W extra rollouts are not needed on a CPU
    rate, dt, thr = cpu_oracle_rate(args.kind, n_envs, n_steps, dense=False)
    rate_dense, dt_dense, _ = cpu_oracle_rate(args.kind, max(n_envs // 8, thr), max(n_steps // 2, 1), dense=True)
    sample = (f"{n_envs} envs x {n_steps} steps of the same reset/action law on the stationary episode-phase distribution (one "
              "untimed episode with staggered step counters first), oracle port (C, float32, OpenMP over envs), active-set rows")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": n_steps,
        "warmup": args.warmup, "ms_per_step": dt / n_steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"airbot_{args.kind} env.step (test/airbot.py + sf.xml), CPU oracle port, {n_envs} envs",
                   "episode_phase": "stationary (same protocol as the GPU arm)",
                   "note": "MJX itself cannot be installed here (SURVEY.md §8c); this is the oracle restatement, not MJX"},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": thr, "kind": "port", "sample": sample,
                         "mjx_work_pattern_value": rate_dense,
                         "mjx_work_pattern_note": "same oracle keeping every geom pair's 4 contact slots x 6 rows dense, as MJX executes"},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores": ncores,
    }
    print(json.dumps(line))


def stagger_episode_phases(env, state, gen):
    """Put env i at a uniformly random phase of its episode: the step counter of the EpisodeWrapper gets a random
    offset, so env i is truncated and auto-reset 1200 - offset_i steps from now; after 1200 further steps every env has
    been reset once and sits offset_i steps into a fresh episode — the stationary distribution of a running job."""
    import torch
    from rsr_mjx_b200 import _lib
    off = torch.randint(0, env.episode_length, (env.num_envs,), device=env.device, generator=gen).float()
    state._buf["info"][:, _lib.INFO["STEPS"]] = off
    state._buf["done"].zero_()  # a set done flag makes the next step restart the counter at 0 (AutoReset pre-step)


def time_env(env, state, actions_fn, W, K, flush, barrier, episode_profile=True):
    """Timing protocol of one env (all on the current stream, CUDA events):
       1. (the caller has run W warm-up steps on a throw-away state; `state` is fresh from reset)
       2. (episode_profile) a whole episode from reset, 1200 steps, timed as one block and in windows at t = 0 / 300 /
          600 / 900 -> the from-reset profile and the full-episode mean;
       3. episode phases staggered, 1200 more steps (timed as one block: the stationary mean, L2 warm);
       4. the K timed steps of the bench contract on that stationary population, one event pair per step with an L2
          flush between steps."""
    import torch
    EP = env.episode_length
    out = {}
    barrier()
    if episode_profile:
        marks = sorted({0, 16, 300, 316, 600, 616, 900, 916, EP})
        ev = {m: torch.cuda.Event(enable_timing=True) for m in marks}
        for t in range(EP):
            if t in ev:
                ev[t].record()
            env.step(state, actions_fn(t))
        ev[EP].record()
        barrier()
        out["episode_from_reset_mean_ms"] = ev[0].elapsed_time(ev[EP]) / EP
        out["from_reset_window_ms"] = {f"t{m}": ev[m].elapsed_time(ev[m + 16]) / 16 for m in (0, 300, 600, 900)}
    gen = torch.Generator(device=env.device).manual_seed(12345)
    stagger_episode_phases(env, state, gen)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(EP):
        env.step(state, actions_fn(t + 7))
    e1.record()
    barrier()
    out["stationary_1200_step_mean_ms_l2_warm"] = e0.elapsed_time(e1) / EP
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    # cudaProfilerStart/Stop around the timed region: `ncu --profile-from-start off` lists exactly these launches
    # (profiles/r2_launches_bench_sf8192.csv); a no-op without a profiler attached
    torch.cuda.profiler.start()
    for t in range(K):
        if flush is not None:
            flush.fill_(t & 0xFF)
        evs[t][0].record()
        env.step(state, actions_fn(t + 3))
        evs[t][1].record()
    barrier()
    torch.cuda.profiler.stop()
    out["step_ms"] = np.array([a.elapsed_time(b) for a, b in evs])
    return out


def synthetic_rsr_files(obs_size, act_size, n=51, seed=2, rollout=None):
    """BASELINE config 4: six synthetic tables (real / past-sim / current-sim observations and actions, 51 rows each)
    -> the five arrays `policy_params_training` takes.  `rollout` = (obs [n, obs_size], actions [n, act_size]) of one env
    under random actions puts the tables where the policy's transitions live, so the KDE sees the online batch and the RSR
    term is non-zero (tables far from the data give exactly 0: every online row underflows out of the density); without it,
    smooth random walks around 0.  "real" / "past sim" / "current sim" differ by noise of 1e-3 / 5e-3 / 2e-3."""
    rng = np.random.default_rng(seed)
    if rollout is not None:
        base, act = np.asarray(rollout[0], np.float32)[:n], np.asarray(rollout[1], np.float32)[:n]
    else:
        base = np.cumsum(rng.normal(0, 0.01, (n, obs_size)), axis=0).astype(np.float32)
        act = rng.uniform(-1, 1, (n, act_size)).astype(np.float32)
    real = base + rng.normal(0, 1e-3, base.shape).astype(np.float32)
    past = base + rng.normal(0, 5e-3, base.shape).astype(np.float32)
    cur = base + rng.normal(0, 2e-3, base.shape).astype(np.float32)
    m = min(n - 1, 50)
    return real[:m], act[:m], real[1:m + 1], past[1:m + 1], cur[1:m + 1]


# KDE bandwidth of the bench's RSR term.  With the reference default (h = 0.1, 10 grid points uniform in [-3, 3]^51,
# rsr_pipeline.py:286-289) the squared distances data -> grid points differ by ~30 / (2 h^2) = 1500 nats, every softmax over
# the grid is one-hot for the reference AND the online density and the term is exactly 0 (true of the reference too); h = 3
# keeps it alive so the gradient path is exercised.  The kernels' cost does not depend on h.
RSR_BENCH_BANDWIDTH = 3.0


def random_action_rollout(kind, dev, n=51, seed=5):
    """observations and actions of env 0 over n steps of U(-1,1) actions (input of synthetic_rsr_files)"""
    import torch
    from rsr_mjx_b200 import prng
    from rsr_mjx_b200.envs import AirbotPlayBase
    env = AirbotPlayBase(kind, num_envs=32, episode_length=1200, device=dev)
    st = env.reset(prng.split(prng.PRNGKey(seed), 32))
    g = torch.Generator(device=dev).manual_seed(seed)
    obs, act = [], []
    for _ in range(n):
        a = torch.rand(32, env.action_size, device=dev, generator=g) * 2 - 1
        obs.append(st.obs[0].cpu().numpy().copy())
        act.append(a[0].cpu().numpy().copy())
        st = env.step(st, a)
    return np.stack(obs), np.stack(act)


def ppo_sub_record(dev, rank, world, training_steps=9):
    """BASELINE config 3 + 4: PPO exactly as ppo_train/airbot_training/train.py:45-55 (1024 envs globally, unroll 10,
    32 x 256 minibatches, 8 updates per batch, lr 1e-4, gamma 0.96, entropy 2e-2, reward scaling 0.1, obs normalisation,
    domain randomisation) WITH the RSR term (past_data, RSR/losses.py:186-195); `training/sps` of RSR/train.py:378-385 =
    env-steps consumed / wall time of a training step, first (graph-capturing) step dropped."""
    from rsr_mjx_b200 import domain_randomize as DR, ppo, prng, rsr_pipeline as RP
    from rsr_mjx_b200.envs import AirbotPlayBase
    N = 1024 // world
    env = AirbotPlayBase("cube", num_envs=N, episode_length=1200, device=dev, randomization_fn=DR.domain_randomize,
                         randomization_rng=prng.split(prng.PRNGKey(1), N))
    past = RP.build_policy_rsr_data(*synthetic_rsr_files(env.observation_size, env.action_size,
                                                         rollout=random_action_rollout("cube", dev)),
                                    bandwidth=RSR_BENCH_BANDWIDTH, device=dev)
    sps, split, rsr = [], [], []
    ppo.train(env, num_timesteps=10**9, episode_length=1200, past_data=past, num_envs=N, learning_rate=1e-4,
              entropy_cost=2e-2, discounting=0.96, unroll_length=10, batch_size=256 // world, num_minibatches=32,
              num_updates_per_batch=8, num_evals=training_steps, normalize_observations=True, reward_scaling=0.1,
              max_training_steps=training_steps, run_evals=False,
              training_step_fn=lambda n, m: (sps.append(m["training/sps"]), rsr.append(m.get("training/sim2real_loss")),
                                             split.append((m["training/collect_s"], m["training/update_s"]))))
    steady = sps[1:] if len(sps) > 1 else sps
    return {"metric": "ppo_train_env_steps_per_sec", "value": float(np.mean(steady)), "unit": "env-steps/s", "n_gpus": world,
            "training_steps_timed": len(steady), "env_steps_per_training_step": 256 * 10 * 32,
            "collect_s": float(np.mean([c for c, _ in split[1:] or split])), "update_s": float(np.mean([u for _, u in split[1:] or split])),
            "rsr_term": True, "sim2real_loss_last": rsr[-1],
            "config": {"workload": "PPO on cube_env + domain randomisation with the RSR KDE+Wasserstein term (BASELINE configs 3+4)",
                       "global_envs": 1024, "unroll_length": 10, "minibatches": "32 x 256 sequences", "updates_per_batch": 8,
                       "past_data": "synthetic six-table set from a 51-step random-action rollout of one env + noise, 50 transitions, grid 10 x 51, h = 3.0 (the default 0.1 makes the term identically 0)"}}


def sac_sub_record(dev, training_steps=300):
    """SAC on the sf env at the reference's RSR defaults (test/rsr_policy_training.py:60-68: 512 envs, batch
    128, min_replay 10 000, max_replay 200 000; one gradient update per actor step), `training/sps` over the steps after
    the replay prefill.  Single GPU only (the trainer has no multi-rank path)."""
    from rsr_mjx_b200 import sac
    from rsr_mjx_b200.envs import AirbotPlayBase
    env = AirbotPlayBase("sf", num_envs=512, episode_length=1200, device=dev)
    seen = []
    sac.train(env, num_timesteps=10**9, episode_length=1200, num_envs=512, batch_size=128, min_replay_size=10_000,
              max_replay_size=200_000, num_evals=5, max_training_steps=training_steps, use_cuda_graph=True, run_evals=False,
              progress_fn=lambda n, m: seen.append(m["training/sps"]))
    return {"metric": "sac_train_env_steps_per_sec", "value": float(seen[-1]), "unit": "env-steps/s", "n_gpus": 1,
            "training_steps_timed": training_steps,
            "config": {"workload": "SAC on airbot_sf, 512 envs, batch 128, 1 gradient update per actor step, replay ring on the device",
                       "update": "torch modules inside CUDA graphs (not the hand-written PPO kernels)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kind", default="sf", choices=["sf", "cube", "T"])
    ap.add_argument("--envs", type=int, default=8192, help="envs per GPU")
    ap.add_argument("--cpu-envs", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the T-shape and PPO sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from rsr_mjx_b200 import sharding
    from rsr_mjx_b200.envs import AirbotPlayBase

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the stepper has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K, N = max(args.warmup, 3), args.steps, args.envs
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make(kind):
        env = AirbotPlayBase(kind, num_envs=N, episode_length=1200, device=dev)
        # env i of the global job = rank * N + i: disjoint reset keys per rank, no communication
        keys = sharding.shard_keys(0, N, rank, world)
        gen = torch.Generator(device=dev).manual_seed(1 + rank)
        actions = torch.rand(64, N, env.action_size, device=dev, generator=gen) * 2 - 1  # U(-1,1)^5, cycled
        warm = env.reset(keys)
        for t in range(W):
            env.step(warm, actions[t % 64])
        return env, env.reset(keys), actions

    env, state, actions = make(args.kind)
    sampler = ClockSampler(local)
    sampler.start()
    tm = time_env(env, state, lambda t: actions[t % 64], W, K, flush, barrier)
    step_ms = tm.pop("step_ms")
    total_ms = float(step_ms.sum())
    # ---- end-to-end timing: host action buffer in, host obs/reward/done out, every step, through the host-buffer entry
    # of the C-ABI (rsrx_env_step_host: H2D action, launch, D2H obs/reward/done queued by one call); same stationary
    # population of envs
    h_act = torch.empty(K, N, env.action_size, dtype=torch.float32).pin_memory()
    h_act.copy_(actions[torch.arange(K) % 64].cpu())
    h_obs = torch.empty(N, env.layout.obs_stride, dtype=torch.float32).pin_memory()
    h_rew = torch.empty(N, dtype=torch.float32).pin_memory()
    h_done = torch.empty(N, dtype=torch.float32).pin_memory()
    for t in range(3):
        env.step_host(state, h_act[t % K], h_obs, h_rew, h_done)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(K):
        env.step_host(state, h_act[t], h_obs, h_rew, h_done)
        torch.cuda.current_stream().synchronize()  # the caller consumes obs before choosing the next action
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    from rsr_mjx_b200 import _lib
    status = state._buf["status"]
    status_counts = {name: int(((status & bit) != 0).sum().item()) for name, bit in
                     (("nonfinite", _lib.STATUS_NONFINITE), ("contact_dropped", _lib.STATUS_CONTACT_OVERFLOW),
                      ("newton_iteration_cap", _lib.STATUS_SOLVER_CAP), ("large_capacity_redo", _lib.STATUS_CONTACT_REDO))}

    # ---- sub-records (same JSON line): BASELINE config 2 (T-shape env, 8192 envs) and configs 3+4 (PPO with the RSR term)
    sub = {}
    if not args.no_sub:
        if args.kind != "T":
            envT, stateT, actT = make("T")
            tT = time_env(envT, stateT, lambda t: actT[t % 64], W, K, flush, barrier, episode_profile=False)
            msT = float(sharding.reduce_max([float(tT["step_ms"].sum())], device=dev)[0])
            sub["airbot_T_env_steps_per_sec"] = {
                "value": N * world * K / (msT * 1e-3), "unit": "env-steps/s", "ms_per_step": msT / K, "n_gpus": world,
                "config": {"workload": "airbot_T env.step (T_shape_env.py + T_shape.xml, box-box contacts), stationary episode phases",
                           "envs_per_gpu": N}}
            del envT, stateT, actT
        barrier()
        sub["ppo_train_env_steps_per_sec"] = ppo_sub_record(dev, rank, world)
        barrier()
        if world == 1:
            sub["sac_train_env_steps_per_sec"] = sac_sub_record(dev)

    total_ms, e2e_ms = sharding.reduce_max([total_ms, e2e_ms], device=dev)
    if rank == 0:
        peak, peak_src = read_peaks()
        value = N * world * K / (total_ms * 1e-3)
        kern_ms = float(np.mean(step_ms))
        achieved = ALGO_BYTES[args.kind] * N / (kern_ms * 1e-3) / 1e9
        clocks = sampler.summary()
        summary = read_ncu_summary(args.kind, N)
        cv = compute_view(summary, N, kern_ms, clocks.get("sm_mhz"))
        traffic = summary["raw"].get("dram__bytes_read.sum", 0.0) + summary["raw"].get("dram__bytes_write.sum", 0.0) \
            if summary and not summary["stale"] else None
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"airbot_{args.kind} env.step (test/airbot.py + sf.xml; 4 x mjx.step + reward/obs/done + episode/auto-reset)",
                       "envs_per_gpu": N, "global_envs": N * world, "actions": "U(-1,1)^5 pre-generated on device",
                       "reset": "jax-style keys split(PRNGKey(0)), reference reset law", "domain_randomization": False,
                       "episode_phase": "stationary: envs staggered uniformly over the 1200-step episode before the timed steps "
                                        "(a whole episode from reset and 1200 staggering steps run first, outside the timed region)",
                       "parallelism": f"dp{world} (envs sharded, no collective)",
                       "l2": "256 MiB flush between timed steps" if flush is not None else "no flush"},
            "clocks": clocks,
            "e2e": {"value": N * world * K / (e2e_ms * 1e-3), "unit": "env-steps/s",
                    "h2d_bytes_per_step": N * env.action_size * 4, "d2h_bytes_per_step": N * (env.layout.obs_stride + 2) * 4},
            # two launches of ours per env.step: step_kernel and the large-capacity pass over the (normally empty) redo list
            "gpu_launches": 2 * K,
            "episode_profile": {k: (v if isinstance(v, dict) else float(v)) for k, v in tm.items()},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": ALGO_BYTES[args.kind],
                         "issue_slot_frac": (cv or {}).get("issue_slot_frac_this_run"),
                         "compute_view": cv,
                         "note": "scan-like state-in/state-out step: far below the HBM roof by design (SURVEY.md §8d); the "
                                 "kernel is issue/latency-bound, so the issue-slot fraction next to `frac` is the one to watch"},
            "status_flagged_envs": status_counts,
            "sub": sub,
        }
        if not args.no_cpu_baseline and world == 1:
            rate, dt, thr = cpu_oracle_rate(args.kind, args.cpu_envs, 300, dense=False)  # ~14 s untimed advance + ~4 s timed on 16 threads
            line["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": thr, "kind": "port",
                                    "sample": f"{args.cpu_envs} envs x 300 steps on the stationary episode-phase distribution (one untimed "
                                              f"episode first), oracle port (C float32, OpenMP), active-set rows; {dt:.1f} s timed"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
