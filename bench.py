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


def read_traffic(kind, n_envs):
    """dram bytes per launch of step_kernel from the committed ncu summary (captured at 8192 envs), if any"""
    p = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(p) and n_envs == 8192:
        with open(p) as f:
            return json.load(f).get(kind)
    return None


def read_compute_view(n_envs):
    """the compute-side view of step_kernel from the committed `ncu --set full` summary
    (fp32 pipe, issue slots, IPC, resident warps) — the numbers that actually explain an issue/latency-bound kernel"""
    p = os.path.join(ROOT, "profiles", "r1_step_kernel_ncu_full_summary.csv")
    if not os.path.exists(p) or n_envs != 8192:
        return None
    want = {"sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fp32_fma_pipe_pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
            "sm__inst_executed.avg.per_cycle_active": "ipc_per_sm",
            "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
            "smsp__inst_issued.sum": "warp_instructions_per_launch"}
    out = {"source": "profiles/r1_step_kernel_ncu_full_summary.csv (ncu --set full, one launch, 8192 envs)"}
    raw = {}
    with open(p) as f:
        for ln in f:
            parts = ln.strip().split(",")
            if len(parts) == 3:
                raw[parts[0]] = parts[2]
                if parts[0] in want:
                    out[want[parts[0]]] = float(parts[2])
    # executed fp32 work: thread-level FADD + FMUL + 2 x FFMA per elapsed cycle, summed over the chip
    try:
        per_cycle = (float(raw["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed"])
                     + float(raw["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed"])
                     + 2.0 * float(raw["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed"]))
        flop = per_cycle * float(raw["sm__cycles_elapsed.max"])
        out["fp32_flop_per_env_step"] = flop / n_envs
        out["fp32_tflops_achieved"] = flop / (float(raw["gpu__time_duration.sum"]) * 1e-3) / 1e12
        out["fp32_tflops_peak"] = 148 * 128 * 2 * float(raw["sm__cycles_elapsed.avg.per_second"]) * 1e9 / 1e12
    except (KeyError, ValueError):
        pass
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


def cpu_oracle_rate(kind, n_envs, n_steps, dense, threads=0, precision="f32"):
    """env-steps/s of the CPU oracle (OpenMP over envs) on a bounded sample of the same workload"""
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
    sample = f"{n_envs} envs x {n_steps} steps of the same reset/action law, oracle port (C, float32, OpenMP over envs), active-set rows"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": n_steps,
        "warmup": args.warmup, "ms_per_step": dt / n_steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"airbot_{args.kind} env.step (test/airbot.py + sf.xml), CPU oracle port, {n_envs} envs",
                   "note": "MJX itself cannot be installed here (SURVEY.md §8c); this is the oracle restatement, not MJX"},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": thr, "kind": "port", "sample": sample,
                         "mjx_work_pattern_value": rate_dense,
                         "mjx_work_pattern_note": "same oracle keeping every geom pair's 4 contact slots x 6 rows dense, as MJX executes"},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores": ncores,
    }
    print(json.dumps(line))


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

    env = AirbotPlayBase(args.kind, num_envs=N, episode_length=1200, device=dev)
    # env i of the global job = rank * N + i: disjoint reset keys per rank, no communication
    keys = sharding.shard_keys(0, N, rank, world)
    state = env.reset(keys)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    actions = torch.rand(W + K, N, env.action_size, device=dev, generator=gen) * 2 - 1
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    # ---- device-resident timing
    for t in range(W):
        env.step(state, actions[t])
    barrier()
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for t in range(K):
        if flush is not None:
            flush.fill_(t & 0xFF)
        evs[t][0].record()
        env.step(state, actions[W + t])
        evs[t][1].record()
    barrier()
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    total_ms = float(step_ms.sum())
    # ---- end-to-end timing: host action buffer in, host obs/reward/done out, every step, through the host-buffer entry
    # of the C-ABI (rsrx_env_step_host: H2D action, launch, D2H obs/reward/done queued by one call)
    h_act = torch.empty(K, N, env.action_size, dtype=torch.float32).pin_memory()
    h_act.copy_(actions[W:].cpu())
    h_obs = torch.empty(N, env.layout.obs_stride, dtype=torch.float32).pin_memory()
    h_rew = torch.empty(N, dtype=torch.float32).pin_memory()
    h_done = torch.empty(N, dtype=torch.float32).pin_memory()
    for t in range(3):
        env.step_host(state, h_act[t], h_obs, h_rew, h_done)
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
    status_bad = int((state._buf["status"] != 0).sum().item())

    total_ms, e2e_ms = sharding.reduce_max([total_ms, e2e_ms], device=dev)
    if rank == 0:
        peak, peak_src = read_peaks()
        value = N * world * K / (total_ms * 1e-3)
        kern_ms = float(np.mean(step_ms))
        achieved = ALGO_BYTES[args.kind] * N / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"airbot_{args.kind} env.step (test/airbot.py + sf.xml; 4 x mjx.step + reward/obs/done + episode/auto-reset)",
                       "envs_per_gpu": N, "global_envs": N * world, "actions": "U(-1,1)^5 pre-generated on device",
                       "reset": "jax-style keys split(PRNGKey(0)), reference reset law", "domain_randomization": False,
                       "parallelism": f"dp{world} (envs sharded, no collective)",
                       "l2": "256 MiB flush between timed steps" if flush is not None else "no flush"},
            "clocks": sampler.summary(),
            "e2e": {"value": N * world * K / (e2e_ms * 1e-3), "unit": "env-steps/s",
                    "h2d_bytes_per_step": N * env.action_size * 4, "d2h_bytes_per_step": N * (env.layout.obs_stride + 2) * 4},
            "gpu_launches": K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": read_traffic(args.kind, N), "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": ALGO_BYTES[args.kind],
                         "compute_view": read_compute_view(N),
                         "note": "scan-like state-in/state-out step: far below the HBM roof by design (SURVEY.md §8d); "
                                 "issue/latency-bound, see profiles/"},
            "status_flagged_envs": status_bad,
        }
        if not args.no_cpu_baseline and world == 1:
            rate, dt, thr = cpu_oracle_rate(args.kind, args.cpu_envs, 1000, dense=False)  # ~10 s on 16 host threads
            line["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": thr, "kind": "port",
                                    "sample": f"{args.cpu_envs} envs x 1000 steps, oracle port (C float32, OpenMP), active-set rows; {dt:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
