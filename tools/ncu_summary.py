"""Summarise an `ncu --set full` capture of b2k_step_kernel: the counters DESIGN.md quotes, per env-step where that
makes sense.  Usage: python tools/ncu_summary.py <report.ncu-rep> <env-steps in the launch> [title]"""
import csv
import subprocess
import sys

rep, nes = sys.argv[1], float(sys.argv[2])
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__sass_inst_executed_op_shared_ld.sum",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
stalls = ["long_scoreboard", "barrier", "wait", "short_scoreboard", "no_instruction", "branch_resolving", "math_pipe_throttle",
          "not_selected", "dispatch_stall", "lg_throttle", "mio_throttle"]
print(f"ncu --set full --clock-control none of ONE b2k_step_kernel launch: {title}")
print("numbers under a profiler are for SHARES and counters, not throughput\n")
for k in want:
    if k in d:
        print(f"{k:80s} {d[k][0]:>22s} {d[k][1]}")
for s in stalls:
    k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
    if k in d:
        print(f"{k:80s} {d[k][0]:>22s}")


def num(k):
    return float(d[k][0].replace(",", "")) if k in d and d[k][0] else 0.0


def to_bytes(k):
    v, u = num(k), d.get(k, ("", ""))[1].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
print(f"\nper env-step ({nes:.0f} env-steps in this launch): warp instructions {num('smsp__inst_executed.sum') / nes:.0f}, "
      f"DRAM bytes read {rd / nes:.0f} + written {wr / nes:.0f} = {(rd + wr) / nes:.0f}")
print(f"local-memory loads {num('smsp__sass_inst_executed_op_local_ld.sum') / nes:.0f} / stores "
      f"{num('smsp__sass_inst_executed_op_local_st.sum') / nes:.0f} per env-step; shared loads "
      f"{num('smsp__sass_inst_executed_op_shared_ld.sum') / nes:.0f}; global loads {num('smsp__sass_inst_executed_op_global_ld.sum') / nes:.0f}")
