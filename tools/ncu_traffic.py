#!/usr/bin/env python3
"""profiles/traffic.json from an .ncu-rep: DRAM bytes per launch of every captured kernel, keyed by the kernel sources they
were measured on (bench.source_sha16), so bench.py quotes them only while the sources are unchanged.
usage: python tools/ncu_traffic.py REP ITEMS_PER_LAUNCH > profiles/traffic.json"""
import csv, io, json, os, re, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench

rep, items = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
kernels = {}
for r in rows[2:]:
    name = re.match(r"[A-Za-z0-9_]+", r[idx["Kernel Name"]]).group(0)
    rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
    wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
    kernels[name] = {"dram_read_bytes": int(rd), "dram_write_bytes": int(wr), "duration_ms_under_ncu": float(r[idx["gpu__time_duration.sum"]])}
json.dump({"source": "%s (ncu --set full --clock-control none; tools/gpu_ncu.sh)" % os.path.basename(rep), "source_sha16": bench.source_sha16(),
           "items_per_launch": items, "kernels": kernels,
           "note": "almost all of it is per-thread stack write-back: the algorithmic HBM traffic of a verification is 225 bytes"}, sys.stdout, indent=1)
