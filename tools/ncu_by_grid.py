"""Per (kernel, grid size) durations from an `ncu --metrics gpu__time_duration.sum,launch__grid_size --csv` list."""
import csv
import sys
from collections import defaultdict

rows = defaultdict(dict)
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    rows[r["ID"]]["name"] = r["Kernel Name"].split("(")[0].replace("void ", "").replace("atspeed::", "")
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows[r["ID"]]["us"] = v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
    elif r["Metric Name"] == "launch__grid_size":
        rows[r["ID"]]["grid"] = int(v)
agg = defaultdict(lambda: [0, 0.0])
for r in rows.values():
    k = (r["name"][:40], r.get("grid", 0))
    agg[k][0] += 1
    agg[k][1] += r.get("us", 0.0)
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1e3:.3f} ms over {len(rows)} launches")
for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{name:<42} grid={grid:<6} n={n:<4} avg={us / n:8.2f} us  total={us / 1e3:7.3f} ms  {us / tot:6.1%}")
