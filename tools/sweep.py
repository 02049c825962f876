#!/usr/bin/env python
"""Kernel timing helper (tuning only): scan-kernel GB/s for a handful of patterns on a device-resident
synthetic corpus.  Usage: python tools/sweep.py [corpus_gib]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, json, os
sys.path.insert(0, %r)
sys.path.insert(0, %r + "/tests")
import numpy as np
import grab_b200 as G
import corpus
gib = float(sys.argv[1])
ctx = G.Context(0)
n = int(gib * 1024)
d = ctx.device_alloc(n << 20)
ctx.synth_corpus(d, 2, 0, n, 1 << 20, needle=b"foobardoesexist", needle_every=64)
batch = ctx.batch_create(G.Context.device_units(d, n, 1 << 20))
out = {}
for name, pat, lit in (("literal", "foobardoesexist", True), ("alt4", "foo|bar|baz|quux", False), ("run16", "[A-Za-z0-9_]{16,}", False), ("lit2", "qz", False), ("icase", "(?i)linus", False), ("lits8", "alpha|bravo|charlie|delta|echo|foxtrot|golf|hotel", False), ("digits", r"\\d{3}-\\d{4}", False), ("run4", "[0-9]{4,}", False), ("lits100", corpus.literals100(), False), ("literal2", "foobardoesexist", True), ("nomatch", "foobardoesnotexist", True)):
    p = G.Pattern(pat, literal=lit)
    ms = []
    for i in range(6):
        r = ctx.batch_scan(p, batch)
        ms.append(ctx.stats()["scan_kernel_ms"])
    st = ctx.stats()
    out[name] = {"best_ms": min(ms[1:]), "med_ms": sorted(ms[1:])[len(ms[1:]) // 2], "gbs": (n << 20) / (min(ms[1:]) * 1e-3) / 1e9,
                 "matches": int(len(r)), "resolve_ms": st["resolve_ms"], "total_ms": st["total_ms"]}
probe = min(ctx.read_probe(d, n << 20)[0] for _ in range(3))
out["read_probe_gbs"] = (n << 20) / (probe * 1e-3) / 1e9
for g in (0, 1):
    out["tma_probe_" + str(g)] = (n << 20) / (min(ctx.tma_probe(batch, g) for _ in range(4)) * 1e-3) / 1e9
print(json.dumps(out))
''' % (ROOT, ROOT)


def main():
    gib = sys.argv[1] if len(sys.argv) > 1 else "16"
    p = subprocess.run([sys.executable, "-c", CHILD, gib], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    if p.returncode != 0:
        print("FAILED", p.stderr.decode()[-500:])
        return
    o = json.loads(p.stdout.decode().strip().splitlines()[-1])
    print("read probe %6.0f GB/s; TMA ring with null consumer: stream geom %6.0f, balanced geom %6.0f GB/s" % (o["read_probe_gbs"], o["tma_probe_0"], o["tma_probe_1"]))
    for k, v in o.items():
        if isinstance(v, dict):
            print("%-8s %6.0f GB/s  (%8d matches, kernel %.3f ms, resolve %.2f ms, call %.2f ms)" % (k, v["gbs"], v["matches"], v["best_ms"], v["resolve_ms"], v["total_ms"]), flush=True)


if __name__ == "__main__":
    main()
