"""Summarise an ``ncu --set full`` report (.ncu-rep) into profiles/: one markdown row per captured launch (duration, DRAM
bytes and throughput, tensor-pipe activity, L2 throughput, occupancy limiter inputs) and, for the dominant kernel
(conv3_kd3_kernel on 64->64 @ 8x80x96x80), profiles/top_kernel_ncu.json -- the file bench.py reads ``roofline.traffic``
from.  Runs where ``ncu`` is installed (the authoring container can import reports without a GPU).

    python tools/ncu_summarize.py gpurun_out/hot.ncu-rep profiles/r02_ncu_hot_kernels.md --commit $(git rev-parse --short HEAD)
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WANT = {
    "dur_us": r"gpu__time_duration\.sum$",
    "dram_rd": r"dram__bytes_read\.sum$",
    "dram_wr": r"dram__bytes_write\.sum$",
    "dram_pct": r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    "tensor_pct": r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active$",
    "tensor_pct_alt": r"sm__inst_executed_pipe_tensor.*pct_of_peak_sustained_active$",
    "l2_pct": r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    "sm_pct": r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    "regs": r"launch__registers_per_thread$",
    "smem_kb": r"launch__shared_mem_per_block_dynamic$",
    "tc_smem_pct": r"l1tex__data_pipe_tc_wavefronts_mem_shared.*pct",
}
UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "s": 1e6,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def load(rep):
    if rep.endswith(".csv"):       # the raw page exported on the GPU box (ncu -i X.ncu-rep --page raw --csv)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    rows = [r for r in rows if r and not r[0].startswith("==")]
    head, units, body = rows[0], rows[1], rows[2:]
    cols = {}
    for key, pat in WANT.items():
        for i, h in enumerate(head):
            if re.search(pat, h):
                cols.setdefault(key, i)
    name_i, grid_i, block_i = head.index("Kernel Name"), head.index("Grid Size"), head.index("Block Size")
    res = []
    for r in body:
        d = {"kernel": re.sub(r"\(.*", "", r[name_i]).replace("void sivae::", "").replace("sivae::", ""), "grid": r[grid_i],
             "block": r[block_i]}
        for key, i in cols.items():
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            u = units[i]
            if key == "dur_us" or key.startswith("dram_r") or key.startswith("dram_w"):
                v *= UNIT.get(u, 1.0)
            d[key] = v
        if "tensor_pct" not in d and "tensor_pct_alt" in d:
            d["tensor_pct"] = d["tensor_pct_alt"]
        res.append(d)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out_md")
    ap.add_argument("--commit", default="unknown")
    ap.add_argument("--command", default="")
    ap.add_argument("--top-kernel", default="conv3_kd3_kernel")
    ap.add_argument("--top-shape", default="64->64 @ 8x80x96x80")
    ap.add_argument("--no-top", action="store_true", help="do not rewrite profiles/top_kernel_ncu.json")
    a = ap.parse_args()
    rows = load(a.rep)
    lines = [f"# ncu --set full --clock-control none, source commit {a.commit}", "",
             f"Report: `{os.path.basename(a.rep)}` (gpurun scratch, not committed){'; command: `' + a.command + '`' if a.command else ''}.",
             "Per launch; ncu serialises launches and flushes caches, so durations are cold-cache (bench.py's CUDA-event numbers",
             "are the steady-state ones).", "",
             "| kernel | grid | duration us | DRAM read MB | DRAM write MB | DRAM % of peak | tensor pipe % active | L2 % | SM % | regs |",
             "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for d in rows:
        lines.append(f"| `{d['kernel']}` | {d['grid']} | {d.get('dur_us', float('nan')):.1f} | {d.get('dram_rd', 0) / 1e6:.1f} | "
                     f"{d.get('dram_wr', 0) / 1e6:.1f} | {d.get('dram_pct', float('nan')):.1f} | "
                     f"{d.get('tensor_pct', float('nan')):.1f} | {d.get('l2_pct', float('nan')):.1f} | "
                     f"{d.get('sm_pct', float('nan')):.1f} | {int(d.get('regs', 0))} |")
    with open(a.out_md, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
    top = [d for d in rows if d["kernel"].startswith(a.top_kernel)]
    if top and not a.no_top:
        d = max(top, key=lambda r: r.get("dur_us", 0.0))
        js = {"kernel": a.top_kernel, "shape": a.top_shape, "duration_us": round(d.get("dur_us", 0.0), 1),
              "dram_bytes_per_launch": d.get("dram_rd", 0.0) + d.get("dram_wr", 0.0),
              "dram_read_bytes": d.get("dram_rd"), "dram_write_bytes": d.get("dram_wr"),
              "tensor_pipe_active_pct": round(d.get("tensor_pct", float("nan")), 1),
              "dram_pct_of_peak": round(d.get("dram_pct", float("nan")), 1), "commit": a.commit,
              "summary": os.path.basename(a.out_md)}
        with open(os.path.join(ROOT, "profiles", "top_kernel_ncu.json"), "w") as f:
            json.dump(js, f, indent=1)
        print("wrote profiles/top_kernel_ncu.json:", js)


if __name__ == "__main__":
    sys.exit(main())
