"""Dev tool: per-instruction stall samples of one kernel from an .ncu-rep (run here on the CPU box).
usage: python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [min_samples]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; mins = int(sys.argv[2]) if len(sys.argv) > 2 else 3
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, v = r[0], r[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for i, n in enumerate(h):
    if n in want or (n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio") and float(v[i] or 0) > 0.02):
        print("%-90s %s" % (n, v[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
print("kernels:", [rows[i][1][:90] for i in starts[:-1]])
rows = rows[starts[which]:starts[which + 1]]
h = rows[1]; idx = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = 0
for r in rows[2:]:
    if len(r) >= len(h): tot += int(r[idx["# Samples"]] or 0)
print("total samples", tot)
for k, r in enumerate(rows[2:]):
    if len(r) < len(h): continue
    s = int(r[idx["# Samples"]] or 0)
    if s >= mins:
        st = {n[6:]: int(r[idx[n]] or 0) for n in stalls if int(r[idx[n]] or 0) > 0}
        print(k, r[idx["Source"]].strip()[:64].ljust(64), s, r[idx["Instructions Executed"]], st)
