#!/usr/bin/env python
"""DRAM traffic of the profiled launches of an .ncu-rep as JSON (read on the CPU box with `ncu -i`):
    python profiles/ncu_traffic.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...] > profiles/rNN/traffic.json
bench.py reads `roofline.traffic` from this file (dram__bytes_read.sum + dram__bytes_write.sum per launch of the
roofline kernel) instead of carrying a constant."""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        def val(k):
            i = col[k]
            return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        out.append({"kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
                    "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
                    "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                    "duration_us_under_ncu": float(r[col["gpu__time_duration.sum"]].replace(",", "")) *
                    {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[col["gpu__time_duration.sum"]], 1.0),
                    "source": path})
    return out


if __name__ == "__main__":
    res = []
    for p in sys.argv[1:]:
        res.extend(launches(p))
    json.dump({"how": "ncu --set full --clock-control none, one launch per entry; dram__bytes_read.sum + dram__bytes_write.sum",
               "launches": res}, sys.stdout, indent=1)
