"""Compact per-launch summary of an .ncu-rep (`--set full`): the metrics profiles/README.md quotes.
Usage: python tools/ncu_summary.py report.ncu-rep [max_rows]"""
import csv
import io
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
     ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue act %"),
     ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
     ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe %"),
     ("smsp__inst_executed.sum", "warp instr"),
     ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
     ("lts__t_sector_hit_rate.pct", "L2 hit %"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts")]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio")
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(body)
for r in body[:n]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    print(f"## {name}")
    parts = []
    for key, label in M:
        if key in col and r[col[key]] != "":
            parts.append(f"{label} {r[col[key]]} {units[col[key]]}".strip())
    print("   " + "; ".join(parts))
    st = []
    for h, i in col.items():
        m = STALL.match(h)
        if m and r[i] not in ("", "0"):
            try:
                st.append((float(r[i]), m.group(1)))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("   stalls per issue: " + ", ".join(f"{k} {v:.2f}" for v, k in st[:7]))
