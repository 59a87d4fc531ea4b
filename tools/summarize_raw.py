"""Compact table of an `ncu --page raw --csv` export (one row per kernel launch) -> markdown on stdout.
    python tools/summarize_raw.py <raw.csv> [title]"""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor inst"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
use = [(k, n) for k, n in COLS if k in idx]
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}\n")
print("| kernel | " + " | ".join(f"{n} ({units[idx[k]]})" if units[idx[k]] else n for k, n in use) + " |")
print("|---|" + "---|" * len(use))
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
    name = re.sub(r"void |ovo::|\(anonymous namespace\)::|<unnamed>::", "", name)[:48]
    print(f"| `{name}` | " + " | ".join(r[idx[k]] for k, n in use) + " |")
