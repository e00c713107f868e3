"""Reduce an `ncu --metrics gpu__time_duration.sum --csv` launch list to a per-kernel table (profiles/)."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
            rows.append((r["Kernel Name"].split("(")[0], ns))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        agg[k][0] += 1
        agg[k][1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total device time (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':<60} {'launches':>8} {'total_us':>10} {'avg_us':>9} {'share':>7}")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:<60} {n:>8} {ns / 1e3:>10.1f} {ns / 1e3 / n:>9.2f} {ns / tot:>7.1%}")


if __name__ == "__main__":
    main(sys.argv[1])
