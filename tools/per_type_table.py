"""Per-type-pair table for profiles/: kernel time per launch from an `ncu --metrics gpu__time_duration.sum` launch list of
one timed bench step (each launch alone on the GPU, serialised, cold cache), pairs/s alone, the model flops per pair
(SURVEY.md section 8(d) with the mean iteration count of that type pair, from the oracle on a 40,960-pair sample of the same
distribution) and the resulting fraction of the measured FP64 peak.

    python tools/per_type_table.py launches.csv out.md [--peak-tflops 36.4] [--pairs-per-launch 209715]
"""
import csv
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

NAMES = {(100, 6): "box", (0, 8): "polytope8", (0, 6): "polytope6", (0, 0): "polytope-n", (1, 0): "capsule", (2, 0): "cylinder",
         (3, 0): "cone", (4, 0): "sphere", (5, 5): "polygon5", (5, 0): "polygon-n", (6, 0): "ellipsoid"}


def main():
    args = sys.argv[1:]
    peak, ppl = 36.4, 209715

    def opt(name, cast, default):
        if name in args:
            i = args.index(name)
            v = cast(args[i + 1])
            del args[i:i + 2]
            return v
        return default
    peak = opt("--peak-tflops", float, peak)
    ppl = opt("--pairs-per-launch", int, ppl)
    launches, out = args[0], args[1]
    import oracle as O
    from dcol_trajectory_optimization_b200 import workloads as W
    from dcol_trajectory_optimization_b200.shapes import POLYTOPE, flatten_shapes, flop_model, problem_dims
    n = 40 * 1024
    shapes, i1, i2, p1, p2 = W.config4_batch(n, seed=5)
    rec, A, b = flatten_shapes(shapes)
    its = O.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=O.GRAD_NONE)["iters"]
    kinds = {}
    for s_idx in range(len(rec)):
        t, f = int(rec["type"][s_idx]), int(rec["n_faces"][s_idx])
        is_box = t == POLYTOPE and f == 6
        kinds[s_idx] = (100, 6) if is_box else ((t, f) if t in (0, 5) else (t, 0))
    model = {}
    for g in range(40):
        sel = np.arange(g, n, 40)
        a, c = int(i1[g]), int(i2[g])
        m_ort, q1, q2, nn = problem_dims(rec[a], rec[c])
        fa = (m_ort, q1, q2, nn, int(rec["n_faces"][a]), int(rec["n_faces"][c]), int(rec["type"][a]) == POLYTOPE,
              int(rec["type"][c]) == POLYTOPE)
        f0 = flop_model(*fa, 0)
        fit = flop_model(*fa, 1) - f0
        mean_it = float(its[sel].mean())
        wm = np.array([its[sel][w:w + 32].max() for w in range(0, len(sel), 32)]).mean()
        model[(kinds[a], kinds[c])] = (f0 + mean_it * fit, mean_it, float(mean_it / wm))
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
    hdr = rows[0]
    iN, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
    lines = ["| primitive 1 | primitive 2 | time per launch (us) | M pairs/s alone | mean PDIP iterations | lane efficiency (mean / per-warp max) | model flop / pair | fraction of FP64 peak |",
             "|---|---|---|---|---|---|---|---|"]
    tot_t = tot_f = 0.0
    seen = set()
    for r in rows[1:]:
        if "pair_kernel" not in r[iN]:
            continue
        m = re.findall(r"Prim<(\d+), (\d+)", r[iN])
        k = ((int(m[0][0]), int(m[0][1])), (int(m[1][0]), int(m[1][1])))
        if k not in model or k in seen:      # a launch list may run into the next step: first occurrence only
            continue
        seen.add(k)
        t_us = float(r[iV].replace(",", "")) / 1e3
        flops, mean_it, eff = model[k]
        frac = flops * ppl / (t_us * 1e-6) / (peak * 1e12)
        tot_t += t_us
        tot_f += flops * ppl
        lines.append(f"| {NAMES[k[0]]} | {NAMES[k[1]]} | {t_us:.1f} | {ppl / t_us:.0f} | {mean_it:.2f} | {eff:.2f} | {flops:.0f} | {frac:.2f} |")
    lines.append(f"| **all {len(lines) - 2}** | | {tot_t:.0f} | {ppl * (len(lines) - 2) / tot_t:.0f} (serialised) | | | | {tot_f / (tot_t * 1e-6) / (peak * 1e12):.2f} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
