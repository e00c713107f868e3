"""Reduce one or more `ncu --set full` reports to profiles/ncu_traffic.json + a readable table:
per kernel, DRAM bytes (read + write) and duration per launch, DRAM / tensor-pipe utilisation.
usage: python tools/ncu_traffic.py out.json MODE rep1.ncu-rep [rep2.ncu-rep ...]   (runs `ncu -i ... --page raw --csv`)
MODE names the configuration the capture was taken in ("cohort" | "single" | "standalone"); entries are stored under
"<kernel>/<MODE>" and MERGED into out.json, so bench.py can only ever pick the capture of the configuration it timed."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3,
         "usecond": 1.0, "msecond": 1e3, "second": 1e6}
COLS = {"rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "us": "gpu__time_duration.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pct": "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l2_hit": "lts__t_sector_hit_rate.pct", "regs": "launch__registers_per_thread", "grid": "launch__grid_size",
        "lts_bytes": "lts__t_bytes.sum"}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main(out_json, mode, reps):
    agg = defaultdict(lambda: defaultdict(list))
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        for r in rows[2:]:
            name = r[ki].split("(")[0].replace("void ", "").replace("atspeed::", "")
            for key, col in COLS.items():
                if col in hdr:
                    i = hdr.index(col)
                    v = num(r[i])
                    if v is not None:
                        agg[name][key].append(v * UNITS.get(units[i], 1))
    res = {}
    print(f"{'kernel':<44} {'n':>3} {'us':>8} {'dramMB':>8} {'GB/s':>7} {'dram%':>6} {'tens%':>6} {'L2hit%':>7} {'regs':>5}")
    for k, d in sorted(agg.items(), key=lambda kv: -sum(kv[1]["us"])):
        n = len(d["us"])
        mean = lambda key: sum(d[key]) / len(d[key]) if d[key] else None
        by = (mean("rd") or 0) + (mean("wr") or 0)
        res[k + "/" + mode] = {"dram_bytes_per_launch": by, "dram_read_bytes_per_launch": mean("rd"), "dram_write_bytes_per_launch": mean("wr"),
                  "launches_captured": n, "avg_us_under_ncu": mean("us"), "dram_throughput_pct": mean("dram_pct"),
                  "tensor_pipe_pct": mean("tensor_pct"), "l2_hit_pct": mean("l2_hit"), "lts_bytes_per_launch": mean("lts_bytes"),
                  "source": [r.split("/")[-1] for r in reps]}
        print(f"{k[:44]:<44} {n:>3} {mean('us'):>8.1f} {by / 1e6:>8.2f} {by / mean('us') / 1e3:>7.0f} "
              f"{(mean('dram_pct') or 0):>6.1f} {(mean('tensor_pct') or 0):>6.1f} {(mean('l2_hit') or 0):>7.1f} {int(mean('regs') or 0):>5}")
    old = {}
    if os.path.exists(out_json):
        try:
            old = {k: v for k, v in json.load(open(out_json)).items() if "/" in k}      # un-tagged (round 1) entries are dropped
        except Exception:
            old = {}
    old.update(res)
    json.dump(old, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
