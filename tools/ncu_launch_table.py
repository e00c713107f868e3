#!/usr/bin/env python
"""Per-LAUNCH table from an `ncu --set full` report: duration, DRAM bytes read/written, L2->SM bytes, achieved DRAM GB/s,
tensor-pipe utilisation, L2 hit rate, registers (runs `ncu -i REP --page raw --csv`).  The tables committed under profiles/
come from this.   usage: python tools/ncu_launch_table.py REP.ncu-rep [label]"""
import csv
import io
import subprocess
import sys

COLS = [("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs", "launch__registers_per_thread"),
        ("us", "gpu__time_duration.sum"), ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
        ("l2_to_sm_MB", "l1tex__m_xbar2l1tex_read_bytes.sum"), ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor_pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("l2_hit_pct", "lts__t_sector_hit_rate.pct"), ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active")]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0,
         "msecond": 1e3, "second": 1e6}


def main(rep, label=""):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# {label or rep}: one row per captured launch (ncu --set full --clock-control none; durations are under replay)")
    print(f"{'kernel':<44}" + "".join(f"{n:>12}" for n, _ in COLS) + f"{'dram_GB/s':>11}")
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("atspeed::", "")[:43]
        vals = {}
        for n, c in COLS:
            if c in hdr:
                i = hdr.index(c)
                try:
                    vals[n] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
                except ValueError:
                    vals[n] = None
            else:
                vals[n] = None
        gbs = ((vals["dram_rd_MB"] or 0) + (vals["dram_wr_MB"] or 0)) / vals["us"] * 1e3 if vals["us"] else 0
        print(f"{name:<44}" + "".join(f"{(vals[n] if vals[n] is not None else float('nan')):>12.2f}" for n, _ in COLS) + f"{gbs:>11.0f}")


if __name__ == "__main__":
    main(*sys.argv[1:3])
