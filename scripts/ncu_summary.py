"""Summarise an .ncu-rep (read here, no GPU): per-kernel key metrics + hot SASS blocks.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--blocks]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_inst0.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum", "sm__sass_thread_inst_executed_op_fadd_pred_on.sum",
]
for r in rows[2:]:
    print("=" * 100)
    for k in KEYS:
        if k in hdr:
            v = r[hdr.index(k)]
            print("%-88s %s %s" % (k, v[:90], rows[1][hdr.index(k)]))
if "--blocks" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    # the source page concatenates kernels; split on the "Kernel Name" marker rows
    chunks = src.split('"Kernel Name",')
    for ch in chunks[1:]:
        lines = list(csv.reader(io.StringIO(ch)))
        print("#" * 100)
        print(lines[0][0][:160])
        h = lines[1]
        data = [l for l in lines[2:] if len(l) == len(h)]
        iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        tot = sum(int(r[iE]) for r in data)
        totS = max(1, sum(int(r[iN]) for r in data))
        print("total warp-instructions %d, SASS lines %d" % (tot, len(data)))
        segs, cur = [], None
        for k, r in enumerate(data):
            e = int(r[iE])
            if cur and abs(e - cur["e"]) <= 0.02 * max(e, cur["e"], 1):
                cur["n"] += 1; cur["sum"] += e; cur["samp"] += int(r[iN])
            else:
                cur = dict(start=k, e=e, n=1, sum=e, samp=int(r[iN])); segs.append(cur)
        for s in segs:
            if s["sum"] < 0.006 * tot:
                continue
            ops = {}
            for r in data[s["start"]: s["start"] + s["n"]]:
                t = r[iS].split()
                op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
                ops[op] = ops.get(op, 0) + 1
            top = sorted(ops.items(), key=lambda x: -x[1])[:8]
            print("  idx %5d n %4d exec/inst %9d share %5.1f%% samples %5.1f%%  %s" % (s["start"], s["n"], s["e"], 100 * s["sum"] / tot, 100 * s["samp"] / totS, top))
