#!/usr/bin/env python
"""Runs one pattern over a device-resident synthetic corpus a few times (target for ncu).
Usage: python tools/prof_one.py PATTERN [gib] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus  # noqa: E402
import grab_b200 as G  # noqa: E402

pat = sys.argv[1]
if pat == "@lits100":
    pat = corpus.literals100()
gib = float(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = G.Context(0)
n = int(gib * 1024)
d = ctx.device_alloc(n << 20)
ctx.synth_corpus(d, 2, 0, n, 1 << 20, needle=b"foobardoesexist", needle_every=64)
batch = ctx.batch_create(G.Context.device_units(d, n, 1 << 20))
p = G.Pattern(pat)
for _ in range(reps):
    r = ctx.batch_scan(p, batch)
    st = ctx.stats()
    print("%d matches, kernel %.3f ms = %.0f GB/s" % (len(r), st["scan_kernel_ms"], (n << 20) / st["scan_kernel_ms"] / 1e6))
