"""Summarise an .ncu-rep (read here, no GPU needed) into the small text/JSON files kept under profiles/.

  python tools/ncu_summary.py gpurun_out/prof_c2.ncu-rep profiles/r1_c2   -> r1_c2_ncu.md, traffic json entry
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.per_cycle_active", "sm__cycles_elapsed.max",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1}


def main(rep, out_prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu --set full summary of `%s`" % rep, "",
             "(values are per launch, cold-cache & serialised under the profiler; compare shares, not absolutes)", ""]
    traffic = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        lines.append("## %s" % name)
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        vals = {}
        for k in KEYS:
            if k in col:
                lines.append("| %s | %s | %s |" % (k, r[col[k]], units[col[k]]))
                vals[k] = (r[col[k]], units[col[k]])
        try:
            rd = float(vals["dram__bytes_read.sum"][0].replace(",", "")) * UNIT_SCALE.get(vals["dram__bytes_read.sum"][1], 1)
            wr = float(vals["dram__bytes_write.sum"][0].replace(",", "")) * UNIT_SCALE.get(vals["dram__bytes_write.sum"][1], 1)
            short = name.replace("void ", "").split("<")[0].split("(")[0]
            traffic[short] = int(rd + wr)
            lines.append("")
            lines.append("DRAM traffic (read+write) = %d bytes" % int(rd + wr))
        except Exception:
            pass
        lines.append("")
    open(out_prefix + "_ncu.md", "w").write("\n".join(lines))
    json.dump(traffic, open(out_prefix + "_traffic.json", "w"), indent=1)
    print("wrote", out_prefix + "_ncu.md", traffic)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
