#!/usr/bin/env python
"""Tuning helper: gscan_scan_batch on PAGEABLE host memory (what the CLI's mmap windows are) for several numbers of
staging threads (GSCAN_STAGE_THREADS).  Usage: python tools/stage_sweep.py [gib]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import grab_b200 as G  # noqa: E402

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4
FILE_LEN = 1 << 20
n = int(gib * 1024)
ctx = G.Context(0)
d = ctx.device_alloc(n * FILE_LEN)
ctx.synth_corpus(d, 2, 0, n, FILE_LEN, needle=b"foobardoesexist", needle_every=64)
pag = np.empty(n * FILE_LEN, dtype=np.uint8)
G.lib().gscan_memcpy_d2h(ctx._h, pag.ctypes.data, d, n * FILE_LEN)
pat = G.Pattern("foobardoesexist")
for unit_mib in (1, 64):
    k = n // unit_mib
    units = np.zeros(k, dtype=G.UNIT_DTYPE)
    units["ptr"] = pag.ctypes.data + np.arange(k, dtype=np.uint64) * np.uint64(FILE_LEN * unit_mib)
    units["len"] = FILE_LEN * unit_mib
    units["file_id"] = np.arange(k, dtype=np.uint32)
    for th in (4, 8, 12, 16, 24, 32):
        os.environ["GSCAN_STAGE_THREADS"] = str(th)
        ctx.scan_units(pat, units)
        best = None
        for _ in range(4):
            t0 = time.perf_counter()
            r = ctx.scan_units(pat, units)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print("units of %3d MiB, %2d staging threads: %6.1f GB/s (%d matches)" % (unit_mib, th, n * FILE_LEN / best / 1e9, len(r)), flush=True)
