#!/usr/bin/env python
"""`ncu -i <rep> --page raw --csv` -> profiles/<name>.csv (metric,unit,value for the metrics DESIGN.md / bench.py quote)
+ <name>.meta.json carrying the hash of the kernel sources the capture was taken from (bench.py refuses to quote a
summary whose hash differs from the current sources).
usage: ncu_summary.py <raw.csv> <kernel-name substring> <out.csv> <kind> <envs> [launch index among the matches, default last]"""
import csv, json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench

WANT = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__occupancy_limit_shared_mem launch__occupancy_limit_registers
launch__waves_per_multiprocessor sm__warps_active.avg.pct_of_peak_sustained_active dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__inst_executed.avg.per_cycle_active smsp__issue_active.avg.pct_of_peak_sustained_active smsp__warps_eligible.avg.per_cycle_active
smsp__inst_issued.sum smsp__inst_executed.sum sm__icc_request_hit_rate.pct sm__icc_requests.sum
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active smsp__thread_inst_executed_per_inst_executed.ratio
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__sass_average_branch_targets_threads_uniform.pct sm__cycles_elapsed.avg.per_second sm__cycles_elapsed.max
smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed
smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed""".split()

raw, kern, out, kind, envs = sys.argv[1:6]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
match = [r for r in rows[2:] if kern in r[ki]]
if not match:
    sys.exit(f"no kernel matching {kern!r} in {raw}")
r = match[int(sys.argv[6]) if len(sys.argv) > 6 else -1]
col = {}
for i, h in enumerate(hdr):
    col.setdefault(h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1][:1].isupper() else h, i)
with open(out, "w") as f:
    f.write("metric,unit,value\n")
    for m in WANT:
        i = col.get(m, hdr.index(m) if m in hdr else None)
        if i is None:
            cands = [j for j, h in enumerate(hdr) if h.endswith(m)]
            i = cands[0] if cands else None
        if i is not None and r[i] != "":
            f.write(f"{m},{units[i]},{r[i].replace(',', '')}\n")
json.dump({"kind": kind, "envs": int(envs), "kernel": r[ki], "block": r[hdr.index('Block Size')], "grid": r[hdr.index('Grid Size')],
           "source_hash": bench.kernel_source_hash(), "raw_csv": os.path.basename(raw)},
          open(out.replace(".csv", ".meta.json"), "w"), indent=1)
print("wrote", out)
