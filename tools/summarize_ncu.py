"""Writes compact, committed summaries of ncu outputs into profiles/ (the .ncu-rep files stay in gpurun_out/).
    python tools/summarize_ncu.py <tag> <launches.csv> [rep1.ncu-rep ...]"""
import collections
import csv
import re
import subprocess
import sys

tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row.get("Metric Unit", "")
    v = v / 1e3 if u in ("nsecond", "ns") else (v * 1e3 if u in ("msecond", "ms") else v)
    k = re.sub(r"\(.*", "", row["Kernel Name"])
    k = re.sub(r"void |ovo::", "", k)[:80]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_launches_summary.md", "w") as f:
    f.write(f"# {tag} — ncu launch list ({sum(v[0] for v in agg.values())} launches of a bench.py run)\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file ... python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n")
    f.write("(cold-cache, serialised per-launch times: compare SHARES with bench.py's `roofline.step_breakdown_ms`, not absolutes)\n\n")
    f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        f.write(f"| `{k}` | {c} | {t:.1f} | {100 * t / tot:.1f}% | {t / c:.1f} |\n")
    g = 100 * sum(t for k, (c, t) in agg.items() if "gemm_bf16" in k) / tot
    f.write(f"\nTotal {tot / 1e3:.1f} ms. GEMM class (`gemm_bf16_tn_kernel<*>`) share: {g:.1f}%.\n")
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size']
if reps:
    with open(f"profiles/{tag}_ncu_full_summary.md", "w") as f:
        f.write(f"# {tag} — `ncu --set full --clock-control none --import-source on` captures (tools/profile_target.py)\n\n")
        for rep in reps:
            out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(out.splitlines()))
            hdr, units = rows[0], rows[1]
            idx = [(w, hdr.index(w)) for w in want if w in hdr]
            f.write(f"## {rep.split('/')[-1]}\n\n| " + " | ".join('kernel' if w == 'Kernel Name' else w.split('.')[0] for w, _ in idx) + " |\n|" + "---|" * len(idx) + "\n")
            f.write("| " + " | ".join(units[i] for _, i in idx) + " |\n")
            for r in rows[2:]:
                f.write("| " + " | ".join(r[i][:48] for _, i in idx) + " |\n")
            f.write("\n")
print("written")
