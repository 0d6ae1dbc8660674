#!/usr/bin/env python
"""Emit the markdown tables of DESIGN.md §7 from the committed raw bench lines under profiles/."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
R1 = {"rmat22 clustering": "28.1 ms / 22.2 G", "rmat24 clustering": "100.6 ms / 25.2 G", "rmat22 refinement": "19.3 ms / 30.8 G"}
NAMES = [("rmat22", "clustering", "R-MAT 22, k=16, clustering (config 2, default)"),
         ("rmat24", "clustering", "R-MAT 24, k=64, clustering (config 4 on one GPU)"),
         ("grid512", "clustering", "3-D grid 512^3, k=64, clustering (config 3)"),
         ("road", "clustering", "road-like, k=256, clustering (config 5)"),
         ("rgg24", "clustering", "RGG2D 2^24, k=64, clustering"),
         ("rmat22", "refinement", "R-MAT 22, k=16, refinement"),
         ("grid512", "refinement", "grid 512^3, k=64, refinement"),
         ("road", "refinement", "road-like, k=256, refinement")]


def fmt(x):
    return f"{x:,}".replace(",", " ")


def main():
    rows = [json.loads(l) for l in open(os.path.join(P, "r2_bench_table.jsonl")) if l.strip() and "failed" not in l]
    print("| workload / mode | n (LP-visible) | m directed | ms / step | edges/s resident (`value`) | edges/s e2e | launches / step | dominant sweep kernel | its algorithmic GB/s = fraction of measured HBM peak | all sweeps, GB/s | all sweeps / L2-gather bound | round 1 |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for w, mode, label in NAMES:
        d = next((r for r in rows if r["config"]["workload"] == w and r["config"]["mode"] == mode), None)
        if d is None:
            continue
        c, r = d["config"], d["roofline"]
        e2e = (d.get("e2e") or {}).get("value", 0) / 1e9
        print(f"| {label} | {fmt(c['n'])} | {fmt(c['m_directed'])} | {d['ms_per_step']:.1f} | **{d['value'] / 1e9:.1f} G** | {e2e:.1f} G | "
              f"{d['gpu_launches'] / d['steps']:.0f} | `{r['kernel']}` | {r['achieved']:.0f} = {r['frac']:.3f} | {r['all_sweeps']['achieved']:.0f} | "
              f"{r['gather_bound']['all_sweeps_frac_of_gather_bound']:.2f} | {R1.get(w + ' ' + mode, '—')} |")
    # where a step goes (rmat22 clustering)
    d = next(r for r in rows if r["config"]["workload"] == "rmat22" and r["config"]["mode"] == "clustering")
    r = d["roofline"]
    print()
    print("R-MAT 22 clustering, per-tier CUDA-event times (tiers serialised):",
          ", ".join(f"tier {i} {x:.2f} ms ({e / 1e6:.0f} M edges)" for i, (x, e) in enumerate(zip(r["all_sweeps"]["per_group_ms"], r["all_sweeps"]["per_group_edges"]))),
          f"; commit {r['commit_ms']:.2f} ms; {r['measured_in']}")
    sc = []
    for n in (1, 2, 4, 8):
        f = os.path.join(P, f"r2_scale_n{n}.json")
        if os.path.exists(f):
            try:
                sc.append(json.load(open(f)))
            except Exception:
                pass
    if sc:
        print()
        print("| GPUs | ms / step | edges/s | speed-up vs 1 GPU | efficiency |")
        print("|---|---|---|---|---|")
        base = next((x for x in sc if x["n_gpus"] == 1), sc[0])
        for x in sc:
            sp = base["ms_per_step"] / x["ms_per_step"]
            print(f"| {x['n_gpus']} | {x['ms_per_step']:.1f} | {x['value'] / 1e9:.1f} G | {sp:.2f} | {sp / x['n_gpus'] * base['n_gpus']:.2f} |")


if __name__ == "__main__":
    sys.exit(main())
