"""Extracts the judged numbers from an ncu report into a small text summary (run where ncu is installed):
   python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/<name>_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "sm__cycles_elapsed.max"]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"kernel: {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:90s} {r[i]:>16s} {units[i]}")
        st = [(float(r[i] or 0), h[len(STALL):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr)
              if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")]
        for v, n in sorted(st, reverse=True)[:8]:
            print(f"  stall per issue: {n:40s} {v:8.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
