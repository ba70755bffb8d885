#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for the LAST bench step.
usage: launch_shares.py launches.csv"""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
# one step = from a k_init_cells launch (first kernel of the band phase) to the launch before the next mesh upload /
# band phase; the LAST step of the run is taken (the e2e legs of bench.py run the same kernels as the device leg)
starts = [i for i, r in enumerate(rows) if "k_init_cells" in r[4]]
rows = rows[starts[-1]:]
per_step = len(rows)
agg, order = {}, []
for r in rows:
    name = re.sub(r"^.*::", "", r[4].split("(")[0]).replace("unnamed>", "").strip()
    t = float(r[-1]) / 1e6                       # ns -> ms
    if name not in agg: agg[name] = [0, 0.0]; order.append(name)
    agg[name][0] += 1; agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print(f"# ncu launch list of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e` (last step, {per_step} launches)")
print("# gpu__time_duration.sum per launch, --clock-control none; times are serialised/cold-cache: compare SHARES\n")
for n in sorted(order, key=lambda k: -agg[k][1]):
    print(f"{n:24s} launches {agg[n][0]:3d}  total {agg[n][1]:9.3f} ms  share {100*agg[n][1]/tot:5.1f}%")
print(f"{'step total':24s} launches {per_step:3d}  total {tot:9.3f} ms\n")
for kern in ("k_sweep_columns", "k_relax_rounds", "k_relax_scan"):
    ts = [float(r[-1]) / 1e6 for r in rows if kern in r[4]]
    if ts: print(f"per {kern} launch (ms): " + " ".join(f"{t:.2f}" for t in ts))
