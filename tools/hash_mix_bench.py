#!/usr/bin/env python
"""Hashed engine on text the class prefilter cannot thin out: kernel GB/s of the 100-literal set over (a) printable noise,
(b) lowercase prose (every position passes the class test: the kernel must notice and run its dense path), (c) 64 KiB
stretches of each alternating -- with the sparse path on and off (GSCAN_HASH_PRE=0).  2 GiB device buffers, built on the host.
One JSON line per case.  Needs a B200."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import corpus  # noqa: E402
import grab_b200 as G  # noqa: E402

MiB = 1 << 20


def texts(n):
    rng = np.random.default_rng(5)
    noise = rng.integers(32, 127, n, dtype=np.uint8)
    prose = rng.integers(ord("a"), ord("z") + 1, n, dtype=np.uint8)
    prose[rng.random(n) < 0.15] = 32
    mixed = noise.copy()
    m = mixed.reshape(-1, 65536)
    m[1::2] = prose.reshape(-1, 65536)[1::2]
    return {"noise": noise, "prose": prose, "mixed64k": mixed}


def main():
    chunk = 64 * MiB
    reps = 32  # 2 GiB
    ctx = G.Context(0)
    d = ctx.device_alloc(chunk * reps)
    out = []
    for name, a in texts(chunk).items():
        for k in range(reps):
            ctx.h2d(d + k * chunk, a)
        for pre in ("1", "0"):
            os.environ["GSCAN_HASH_PRE"] = pre
            p = G.Pattern(corpus.literals100())
            b = ctx.batch_create(G.Context.device_units(d, reps * 64, MiB))
            ks, n = [], 0
            for i in range(5):
                r = ctx.batch_scan(p, b)
                n = len(r)
                if i:
                    ks.append(ctx.stats()["scan_kernel_ms"])
            b.free()
            row = {"text": name, "sparse_path": pre == "1", "kernel_gbs": round(chunk * reps / (min(ks) * 1e-3) / 1e9, 1), "matches": int(n)}
            out.append(row)
            print(json.dumps(row), flush=True)
    by = {}
    for r in out:
        by.setdefault(r["text"], {})[r["sparse_path"]] = r["matches"]
    assert all(v[True] == v[False] for v in by.values()), by
    ctx.device_free(d)
    ctx.close()


if __name__ == "__main__":
    main()
