#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep [OUT.md] -- compact, committable summary of one `ncu --set full` capture:
the launch/occupancy/pipe/memory metrics that the roofline discussion in DESIGN.md cites plus the stall-reason
distribution over the sampled warps (runs here on the CPU box: `ncu -i` only reads the report)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__ops_path_tensor_src_fp64.sum",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        out.append("## %s  (grid %s x block %s)\n" % (name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        out.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append("| %s | %s | %s |" % (k, r[i], units[i]))
        out.append("\nstall reasons (warps stalled per issued instruction, `%s*_per_issue_active.ratio`):\n" % STALLS)
        st = []
        for i, h in enumerate(hdr):
            if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h[len(STALLS):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        out.append("| reason | ratio |\n|---|---|")
        for v, n in sorted(st, reverse=True):
            if v >= 0.01:
                out.append("| %s | %.3f |" % (n, v))
        out.append("")
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    else:
        print(text)


if __name__ == "__main__":
    main()
