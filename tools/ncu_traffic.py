#!/usr/bin/env python
"""DRAM traffic of the scan kernel of every BASELINE config at its FULL shape, measured by ncu (dram__bytes_read.sum +
dram__bytes_write.sum of one launch after a warm-up launch), written to profiles/r02_traffic.json -- bench.py reports
`roofline.traffic` from this file only when the launch size matches exactly.  Needs a B200: run under gpurun.
Usage: python tools/ncu_traffic.py [config indices ...]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

which = [int(x) for x in sys.argv[1:]] or [0, 1, 2, 3, 4]
out_path = os.path.join(ROOT, "profiles", "r02_traffic.json")
try:
    res = json.load(open(out_path))
except Exception:
    res = {}
cfgs = bench.baseline_configs()
for ci in which:
    cfg = cfgs[ci]
    cmd = ["ncu", "--csv", "--clock-control", "none", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum",
           "-k", "regex:scan_kernel", "-s", "1", "-c", "1", sys.executable, os.path.join(ROOT, "tools", "prof_cfg.py"), str(ci), "2"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    txt = p.stdout.decode()
    rows = [r for r in csv.reader(io.StringIO(txt[txt.find('"ID"'):])) if len(r) > 5] if '"ID"' in txt else []
    vals = {}
    for r in rows[1:]:
        h = dict(zip(rows[0], r))
        v = float(h["Metric Value"].replace(",", ""))
        unit = h["Metric Unit"].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1, "us": 1e3, "ms": 1e6, "second": 1e9, "s": 1e9}.get(unit, 1)
        vals[h["Metric Name"]] = v * mult
        kernel = h["Kernel Name"]
    if "dram__bytes_read.sum" not in vals:
        print("config %d: ncu gave nothing (%s)" % (ci, p.stderr.decode()[-300:]))
        continue
    nbytes = cfg["n_files"] * cfg["file_len"]
    res[cfg["key"]] = {"algorithmic_bytes_per_launch": nbytes, "dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals["dram__bytes_write.sum"],
                       "dram_bytes_per_launch": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"], "kernel": kernel,
                       "ncu_duration_ms": vals.get("gpu__time_duration.sum", 0) / 1e6,
                       "ratio": (vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]) / nbytes,
                       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, 2nd launch of tools/prof_cfg.py %d (tools/ncu_traffic.py)" % ci}
    print(cfg["key"], json.dumps(res[cfg["key"]]))
json.dump(res, open(out_path, "w"), indent=1)
if os.path.isdir(os.path.join(ROOT, "gpurun_out")):  # the GPU box only sends gpurun_out/ back
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r02_traffic.json"), "w"), indent=1)
