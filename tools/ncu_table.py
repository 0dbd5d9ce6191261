#!/usr/bin/env python
"""Markdown table of an ``ncu --set full`` capture:  ncu -i x.ncu-rep --page raw --csv | python tools/ncu_table.py [title]

One row per captured kernel launch with the numbers the roofline discussion needs: duration, grid / block, registers,
dynamic shared memory, achieved occupancy, DRAM bytes read / written and the resulting DRAM throughput as a fraction of
the measured copy bandwidth (MEASURED_PEAKS.json), L2 / L1 hit rates, SM throughput %, tensor-pipe % where reported."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("time [us]", "gpu__time_duration.sum", 1e-3), ("grid", "launch__grid_size", 1), ("block", "launch__block_size", 1),
        ("regs", "launch__registers_per_thread", 1), ("dyn smem [KB]", "launch__shared_mem_per_block_dynamic", 1e-3),
        ("occ %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("DRAM rd [MB]", "dram__bytes_read.sum", 1e-6),
        ("DRAM wr [MB]", "dram__bytes_write.sum", 1e-6), ("L2 hit %", "lts__t_sector_hit_rate.pct", 1),
        ("L1 hit %", "l1tex__t_sector_hit_rate.pct", 1), ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("tensor %", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 1)]
UNIT_SCALE = {"nsecond": 1.0, "ns": 1.0, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9, "s": 1e9,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def unit_scale(u):
    """ncu prints e.g. 'us', 'Mbyte', 'Kbyte/block', 'byte/second' - scale by the numerator's prefix."""
    return UNIT_SCALE.get(u.split("/")[0].strip(), 1.0)


def main():
    title = sys.argv[1] if len(sys.argv) > 1 else "capture"
    rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
    head, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(head)}
    peak = 6462.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    print(f"## {title}  (ncu --set full --clock-control none; DRAM % is of the measured {peak:.0f} GB/s copy bandwidth)\n")
    names = [c[0] for c in COLS] + ["DRAM GB/s", "DRAM % of peak"]
    print("| kernel | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for r in data:
        vals = {}
        for label, metric, scale in COLS:
            i = idx.get(metric)
            if i is None or r[i] in ("", "n/a"):
                vals[label] = None
                continue
            x = float(r[i].replace(",", "")) * unit_scale(units[i])
            vals[label] = x * scale
        t_us = vals["time [us]"] or 0.0
        rd, wr = vals["DRAM rd [MB]"] or 0.0, vals["DRAM wr [MB]"] or 0.0
        gbs = (rd + wr) / (t_us * 1e-6) / 1e3 if t_us else 0.0
        cells = [("-" if vals[c[0]] is None else f"{vals[c[0]]:.2f}") for c in COLS] + [f"{gbs:.0f}", f"{100 * gbs / peak:.1f}"]
        print(f"| {r[idx['Kernel Name']][:58]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
