"""ncu report -> markdown table: one row per kernel (summed over its launches): time, DRAM bytes, achieved GB/s and the
fraction of the measured copy peak (MEASURED_PEAKS.json, fallback 6650 GB/s), occupancy, registers.

    python scripts/ncu_summarize.py report.ncu-rep [more.ncu-rep ...] > profiles/<name>.md
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback"


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    for line in r[2:]:
        yield {h: (v, u) for h, v, u in zip(hdr, line, units)}


def num(cell):
    v, u = cell
    try:
        return float(v.replace(",", "")) * UNIT.get(u, 1.0)
    except ValueError:
        return 0.0


def main():
    pk, src = peak()
    acc = OrderedDict()
    for rep in sys.argv[1:]:
        for row in rows_of(rep):
            name = row["Kernel Name"][0].split("(")[0]
            a = acc.setdefault(name, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "regs": 0, "occ": 0.0})
            a["n"] += 1
            a["t"] += num(row["gpu__time_duration.sum"])
            a["rd"] += num(row.get("dram__bytes_read.sum", ("0", "")))
            a["wr"] += num(row.get("dram__bytes_write.sum", ("0", "")))
            a["regs"] = max(a["regs"], int(num(row.get("launch__registers_per_thread", ("0", "")))))
            a["occ"] += num(row.get("sm__warps_active.avg.pct_of_peak_sustained_active", ("0", "")))
    print(f"| kernel | launches | time (ms) | DRAM read + written (GB) | GB/s | of the {pk:.0f} GB/s copy peak ({src}) | warps active % | regs |")
    print("|---|---|---|---|---|---|---|---|")
    for name, a in sorted(acc.items(), key=lambda kv: -kv[1]["t"]):
        gb = (a["rd"] + a["wr"]) / 1e9
        gbs = gb / a["t"] if a["t"] else 0.0
        print(f"| `{name}` | {a['n']} | {a['t'] * 1e3:.3f} | {a['rd'] / 1e9:.2f} + {a['wr'] / 1e9:.2f} | {gbs:.0f} | "
              f"{100 * gbs / pk:.1f} % | {a['occ'] / a['n']:.0f} | {a['regs']} |")


if __name__ == "__main__":
    main()
