#!/usr/bin/env python
"""The other BASELINE.json configs' patterns beside the benchmarked one, same corpus, same box: kernel-only GB/s on a
device-resident 16 GiB shard, end-to-end GB/s through gscan_scan_batch from pinned host memory (4 GiB), and the
unmodified reference on the host cores (tmpfs sample).  One JSON object per pattern on stdout.
Usage: python tools/config_bench.py [resident_gib] [cpu_sample_files]        (needs a B200; written after round 1's last
GPU session -- first numbers come from whoever runs it next)"""
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import corpus  # noqa: E402
import grab_b200 as G  # noqa: E402

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
cpu_files = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
FILE_LEN = bench.FILE_LEN
n = int(gib * 1024)
ctx = G.Context(0)
d = ctx.device_alloc(n * FILE_LEN)
ctx.synth_corpus(d, bench.SEED, 0, n, FILE_LEN, needle=bench.PATTERN.encode(), needle_every=bench.NEEDLE_EVERY)
batch = ctx.batch_create(G.Context.device_units(d, n, FILE_LEN))
e_files = min(4096, n)
hptr = G.lib().gscan_host_alloc(e_files * FILE_LEN)
G.lib().gscan_memcpy_d2h(ctx._h, hptr, d, e_files * FILE_LEN)
hunits = np.zeros(e_files, dtype=G.UNIT_DTYPE)
hunits["ptr"] = hptr + np.arange(e_files, dtype=np.uint64) * np.uint64(FILE_LEN)
hunits["len"] = FILE_LEN
hunits["file_id"] = np.arange(e_files, dtype=np.uint32)
sample = bench.materialise(bench.baseline_configs()[1], cpu_files, from_device=(ctx, d))
cores = bench.reference_cores()
try:
    for name, pat, literal in (("configs[1] literal", bench.PATTERN, True), ("configs[2] alternation (non-capturing spelling)", "foo|bar|baz|quux", False),
                               ("configs[3] class run", "[A-Za-z0-9_]{16,}", False), ("configs[4] 100 literals", corpus.literals100(), False)):
        p = G.Pattern(pat, literal=literal)
        for _ in range(2):
            ctx.batch_scan(p, batch)
        ks, ts = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            r = ctx.batch_scan(p, batch)
            ts.append(time.perf_counter() - t0)
            ks.append(ctx.stats()["scan_kernel_ms"])
        ctx.scan_units(p, hunits)
        es = []
        for _ in range(5):
            t0 = time.perf_counter()
            re2e = ctx.scan_units(p, hunits)
            es.append(time.perf_counter() - t0)
        args = [bench.REF_BIN] + (["-n", str(cores)] if cores > 1 else []) + ["-r", "-O", "-l"] + (["-S"] if False else []) + [pat, sample]
        cpu = None
        if os.path.exists(bench.REF_BIN):
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                q = subprocess.run(args, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=600)
                dt = time.perf_counter() - t0
                if q.returncode == 0:
                    best = dt if best is None else min(best, dt)
            cpu = cpu_files * FILE_LEN / best / 1e9 if best else None
        print(json.dumps({"config": name, "pattern": pat if len(pat) < 60 else pat[:57] + "...", "matches": int(len(r)),
                          "kernel_gbs": n * FILE_LEN / (min(ks) * 1e-3) / 1e9, "resident_call_gbs": n * FILE_LEN / min(ts) / 1e9,
                          "e2e_pinned_gbs": e_files * FILE_LEN / min(es) / 1e9, "e2e_matches": int(len(re2e)),
                          "cpu_reference_gbs": cpu, "cpu_cores": cores, "cpu_sample_files": cpu_files}), flush=True)
finally:
    shutil.rmtree(sample, ignore_errors=True)
    G.lib().gscan_host_free(hptr)
