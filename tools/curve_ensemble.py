"""Chaos or bias?    Trains R replicas of every arm of tests/loss_curve.py -- fp32 oracle,
control (oracle under bf16 autocast), ours -- whose initial weights differ by a relative 1e-6 perturbation (identical
across the arms of one replica; batches / noise / eps / masks identical everywhere), and reports per arm the spread of
the loss terms over the replicas at several steps.  A systematic error of the CUDA path shows as a mean shift of
"ours" outside the replica spread of fp32 / control; sensitivity of the recipe shows as a replica spread of the
same size inside every arm.  Test infrastructure (trains the oracle).

    python tools/curve_ensemble.py --vol 40 48 40 --batch 4 --steps 200 --replicas 4 --out profiles/r02_ensemble
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import loss_curve as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vol", type=int, nargs=3, default=[40, 48, 40])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--replicas", type=int, default=4)
    ap.add_argument("--rel", type=float, default=1e-6)
    ap.add_argument("--out", default=None)
    ap.add_argument("--env", nargs="*", default=[])
    ap.add_argument("--warm-start", type=int, default=0)
    a = ap.parse_args()
    for kv in a.env:
        k_, v_ = kv.split("=", 1)
        os.environ[k_] = v_
    runs = []
    for r in range(a.replicas):
        c = L.run(a.steps, tuple(a.vol), a.batch, 4, control=True, perturb=None if r == 0 else (1000 + r, a.rel),
                  warm_start=a.warm_start)
        runs.append(c)
        print(f"replica {r}: loss_rec@last  fp32 {c['oracle']['loss_rec'][-1]:.6g}  control {c['control']['loss_rec'][-1]:.6g}"
              f"  ours {c['ours']['loss_rec'][-1]:.6g}", flush=True)
    marks = [s for s in (0, 1, 5, 10, 25, 50, 100, 150, a.steps - 1) if s < a.steps]
    lines = [f"# Replica ensemble, {a.replicas} replicas (initial weights x (1 + {a.rel:g} N(0,1)), replica 0 unperturbed), "
             f"{a.steps} steps, volumes {a.vol}, batch {a.batch}" + (f", env {a.env}" if a.env else "")
             + (f", warm start {a.warm_start} fp32 steps" if a.warm_start else ""), ""]
    for term in ("loss_rec", "kl_real", "lossE", "lossD"):
        lines += [f"## {term}: 10-step mean around the step, per arm: mean over replicas [min .. max]", "",
                  "| step | fp32 | control | ours |", "|---:|---|---|---|"]
        for s in marks:
            row = [f"| {s} "]
            for arm in ("oracle", "control", "ours"):
                vals = []
                for c in runs:
                    seg = c[arm][term][max(0, s - 4): s + 6]
                    vals.append(sum(seg) / len(seg))
                row.append(f"| {statistics.mean(vals):.5g} [{min(vals):.5g} .. {max(vals):.5g}] ")
            lines.append("".join(row) + "|")
        lines.append("")
    # deviation of every run from the unperturbed fp32 trajectory (the quantity tests/test_loss_curve.py bounds)
    base = runs[0]["oracle"]
    lines += ["## median over steps of |x - fp32 replica 0| / |fp32 replica 0| (loss_rec / kl_real)", "",
              "| replica | fp32 | control | ours |", "|---:|---|---|---|"]
    for r, c in enumerate(runs):
        row = [f"| {r} "]
        for arm in ("oracle", "control", "ours"):
            cell = []
            for term in ("loss_rec", "kl_real"):
                rel = [abs(x - y) / max(abs(y), 1e-30) for x, y in zip(c[arm][term], base[term])]
                cell.append(f"{statistics.median(rel):.3f}")
            row.append("| " + " / ".join(cell) + " ")
        lines.append("".join(row) + "|")
    print("\n".join(lines))
    if a.out:
        with open(a.out + ".md", "w") as f:
            f.write("\n".join(lines) + "\n")
        with open(a.out + ".json", "w") as f:
            json.dump(dict(config=vars(a), runs=runs), f)


if __name__ == "__main__":
    main()
