"""profiles/<tag>_traffic.json from `ncu --set full` captures (gpurun_out/<tag>_{gemm,attn,query,map,fuse}.ncu-rep, tools/r2_capture.sh):
per kernel the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, duration, tensor-pipe %, L2 bytes.  bench.py
reads this file for `roofline.traffic` (a capture of the build whose source hash is recorded here, not a literal).
    python tools/traffic_from_captures.py r2"""
import csv
import hashlib
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def src_hash():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "ovo_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


def to_us(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}[u]


out = {"tag": tag, "csrc_sha16": src_hash(), "how": "ncu --set full --clock-control none (cold caches, one replayed launch at a time)", "kernels": []}
for part in ("gemm", "attn", "query", "map", "fuse"):
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_{part}.ncu-rep")
    if not os.path.exists(rep):
        continue
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def col(r, name, conv=None):
        if name not in ix:
            return None
        v, u = r[ix[name]], units[ix[name]]
        try:
            return conv(v, u) if conv else float(v.replace(",", ""))
        except Exception:
            return None
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        name = re.sub(r"void |ovo::", "", name)
        rd, wr = col(r, "dram__bytes_read.sum", to_bytes), col(r, "dram__bytes_write.sum", to_bytes)
        out["kernels"].append({"capture": part, "kernel": name, "grid": col(r, "launch__grid_size"), "block": col(r, "launch__block_size"),
                               "us": col(r, "gpu__time_duration.sum", to_us), "dram_read_bytes": rd, "dram_write_bytes": wr,
                               "dram_bytes": (rd or 0) + (wr or 0), "l2_bytes": col(r, "lts__t_bytes.sum", to_bytes),
                               "tensor_pipe_pct": col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                               "dram_pct": col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                               "warps_active_pct": col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                               "registers": col(r, "launch__registers_per_thread"),
                               "smem_bank_conflicts": col(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")})
json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
print(f"profiles/{tag}_traffic.json: {len(out['kernels'])} launches, csrc {out['csrc_sha16']}")
