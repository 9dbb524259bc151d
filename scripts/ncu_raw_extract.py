"""ncu report -> the transposed raw-metric CSV kept under profiles/ (one row per metric: name, unit, one column per launch).

    python scripts/ncu_raw_extract.py gpurun_out/prof.ncu-rep > profiles/<name>_ncu_raw.csv
"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, launches = rows[0], rows[1], rows[2:]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"launch_{i + 1}" for i in range(len(launches))])
skip = {"ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC", "Section Name", "Metric Name",
        "Metric Unit", "Metric Value", "Rule Name", "Rule Type", "Rule Description"}
for i, h in enumerate(hdr):
    if h in skip:
        continue
    w.writerow([h, units[i]] + [r[i] for r in launches])
