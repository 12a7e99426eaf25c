#!/usr/bin/env python
"""Write profiles/ncu_traffic.json (read by bench.py's roofline.traffic) from `ncu --set full` captures.

usage: python scripts/ncu_traffic.py <kernel-name> <grid> <columns-per-launch> <file.ncu-rep> [more quadruples ...]
Each entry: dram__bytes_read.sum + dram__bytes_write.sum of the ONE launch in the capture."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_traffic.json")
_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    entries = []
    if os.path.exists(OUT):
        entries = json.load(open(OUT))
    a = sys.argv[1:]
    for kernel, grid, cols, rep in zip(a[0::4], a[1::4], a[2::4], a[3::4]):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, r = rows[0], rows[1], rows[2]
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i].replace(",", "")) * _SCALE[units[i]]
        entries = [e for e in entries if not (e["kernel"] == kernel and e["grid"] == int(grid) and e["columns_per_launch"] == int(cols))]
        entries.append({"kernel": kernel, "grid": int(grid), "columns_per_launch": int(cols), "dram_bytes_per_launch": tot,
                        "ms": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]],
                        "source": f"ncu --set full --clock-control none, one launch of {r[hdr.index('Kernel Name')][:60]} "
                                  f"(dram__bytes_read.sum + dram__bytes_write.sum; {os.path.basename(rep)})"})
    json.dump(entries, open(OUT, "w"), indent=1)
    print(json.dumps(entries, indent=1))


if __name__ == "__main__":
    main()
