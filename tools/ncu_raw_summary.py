"""Summarise `ncu --page raw --csv` output: python tools/ncu_raw_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("---", d.get("Kernel Name", "")[:100], d.get("Grid Size"), d.get("Block Size"))
    for w in want:
        if w in d:
            print(f"   {w:75s} {d[w]} {units[hdr.index(w)]}")
    st = {k: float(v.replace(",", "")) for k, v in d.items()
          if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v}
    print("   stalls:", ", ".join(f"{k[34:-23]}={v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
