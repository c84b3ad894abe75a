"""Summarise an ncu launch list (gpu__time_duration.sum CSV) of bench.py: per kernel / per grid time of ONE step,
delimited by the advance_seed_counter_kernel marker launches."""
import collections
import csv
import re
import sys


def main(path, by_grid=True, top=34):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    idx = [i for i, r in enumerate(rows) if "advance_seed_counter" in r["Kernel Name"]]
    sel = rows[idx[0]:idx[1]] if len(idx) > 1 else rows
    agg = collections.defaultdict(list)
    for row in sel:
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void sivae::", "").replace("sivae::", "")[:44]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v if u == "us" else v * 1e3
        agg[(name, row["Grid Size"] if by_grid else "")].append(v)
    tot = sum(sum(v) for v in agg.values())
    out = [f"one step: {len(sel)} launches, sum of kernel durations {tot / 1e3:.2f} ms", "",
           "| ms | share | n | avg us | kernel | grid |", "|---:|---:|---:|---:|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
        out.append(f"| {sum(v) / 1e3:.3f} | {100 * sum(v) / tot:.1f}% | {len(v)} | {sum(v) / len(v):.1f} | `{k[0]}` | {k[1]} |")
    return "\n".join(out)


if __name__ == "__main__":
    print(main(sys.argv[1], by_grid=(len(sys.argv) < 3)))
