#!/usr/bin/env python
"""Runs the scan of ONE BASELINE config (bench.baseline_configs()[i], full shape) a few times on a device-resident corpus:
the target process for ncu (tools/ncu_traffic.py).  Usage: python tools/prof_cfg.py CONFIG_INDEX [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import grab_b200 as G  # noqa: E402

ci = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = bench.baseline_configs()[ci]
n, flen = cfg["n_files"], cfg["file_len"]
ctx = G.Context(0)
d = ctx.device_alloc(n * flen)
ctx.synth_corpus(d, cfg["seed"], 0, n, flen, needle=cfg["needle"].encode() if cfg["needle"] else None, needle_every=bench.NEEDLE_EVERY if cfg["needle"] else 0)
batch = ctx.batch_create(G.Context.device_units(d, n, flen))
p = G.Pattern(cfg["pattern"], literal=cfg["literal"])
mode = {"ALL": G.MODE_ALL, "FIRST": G.MODE_FIRST}[cfg["mode"]]
for _ in range(reps):
    r = ctx.batch_scan(p, batch, mode, copy=False)
    st = ctx.stats()
    print("%s: %d matches, kernel %.3f ms = %.0f GB/s" % (cfg["key"], len(r), st["scan_kernel_ms"], n * flen / st["scan_kernel_ms"] / 1e6), flush=True)
