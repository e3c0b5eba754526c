"""Summarise an .ncu-rep: key metrics + top stall reasons + hottest source lines.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index("Kernel Name")][:100] if "Kernel Name" in hdr else "")
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"  {h:90s} {r[i]:>16s} {units[i]}")
        stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        if not stalls:
            stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if "warp_issue_stalled" in h and h.endswith(".pct")]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {h:84s} {v:10.3f}")
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        hdr = rows[0]
        print(hdr)
        try:
            ci = hdr.index("Warp Stall Sampling (All Samples)")
        except ValueError:
            ci = None
        if ci is not None:
            body = [r for r in rows[1:] if len(r) > ci and r[ci].isdigit()]
            body.sort(key=lambda r: -int(r[ci]))
            tot = sum(int(r[ci]) for r in body)
            for r in body[:n]:
                print(f"{int(r[ci]) / tot * 100:6.2f}%  {r[hdr.index('Source')][:120] if 'Source' in hdr else r[:3]}")


if __name__ == "__main__":
    main()
