"""Extracts the judged numbers from an ncu report into a small text summary (run where ncu is installed):
   python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/<name>_summary.txt
   python profiles/summarize.py gpurun_out/prof.ncu-rep --json profiles/r02_ncu_constants.json <inner iterations of
          the profiled launch> <what was profiled>      also writes the per-launch constants bench.py quotes"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
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
        last = (hdr, r)
    if "--json" in sys.argv:
        i = sys.argv.index("--json")
        dst, n_inner, what = sys.argv[i + 1], float(sys.argv[i + 2]), sys.argv[i + 3]
        hdr, r = last
        val = lambda k: float(r[hdr.index(k)].replace(",", "")) if k in hdr and r[hdr.index(k)] else None
        unit = lambda k: units[hdr.index(k)] if k in hdr else ""
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = sum((val(k) or 0.0) * scale.get(unit(k), 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        wf = val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        json.dump({"source": f"{path} ({what}); kernel as of commit {commit}",
                   "dram_bytes_per_launch": dram,
                   "smem_wavefronts_per_inner_iteration": (wf / n_inner) if wf else None,
                   "smem_pct_of_peak_sustained_elapsed": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                   "fp64_pipe_pct": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                   "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
                   "eligible_warps_per_cycle": val("smsp__warps_eligible.avg.per_cycle_active"),
                   "registers_per_thread": val("launch__registers_per_thread"),
                   "inner_iterations_of_profiled_launch": n_inner}, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
