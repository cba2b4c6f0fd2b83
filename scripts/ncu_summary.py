"""Summarise `ncu --set full` reports (read here, no GPU): python scripts/ncu_summary.py out.md rep1.ncu-rep [rep2 ...]"""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
out, reps = sys.argv[1], sys.argv[2:]
cols = []
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        cols.append((name, {k: (d.get(k, ""), units[hdr.index(k)] if k in hdr else "") for k in WANT}))
with open(out, "w") as f:
    f.write("| metric | " + " | ".join(c[0] for c in cols) + " |\n|---|" + "---|" * len(cols) + "\n")
    for k in WANT:
        f.write(f"| {k} | " + " | ".join(f"{c[1][k][0]} {c[1][k][1]}".strip() for c in cols) + " |\n")
print(open(out).read())
