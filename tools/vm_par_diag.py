#!/usr/bin/env python
"""Diagnostic (GPU): general patterns on the chain path against the per-unit walk and the oracle, several repetitions
(are the parallel attempts deterministic? is a VM limit reported whenever records are missing?)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import grab_b200 as G  # noqa: E402
import oracle_py as O  # noqa: E402
import test_gpu_random_patterns as R  # noqa: E402

import random  # noqa: E402
rnd = random.Random(99)  # the inputs of tests/test_gpu_random_patterns.py
inputs = []
for k in range(40):
    alpha = [b"abc", b"abcx \n", b"ab", b"abc abc\n\n", b"aabbcc_x1 \t\n"][k % 5]
    ln = rnd.choice([1, 2, 3, 7, 16, 17, 33, 64, 130, 511, 513, 1000, 2050, 4100, 9000])
    inputs.append(bytes(rnd.choice(alpha) for _ in range(ln)))
ctx = G.Context(0)
pats = ["[a-c]{2}[ab]*?\\w[ab]|[ab]{1,3}a{2,}|(?:a *)xa$", "ab+a", "a[ab]*b", "e\\w+ "] + list(R.PATTERNS)[:40]
bad = 0
for pat in pats:
    try:
        p, o = G.Pattern(pat), O.Regex(pat)
    except Exception as e:  # noqa: BLE001
        continue
    if o.nullable or p.info["engine"] != 4:
        continue
    for mode in (G.MODE_ALL, G.MODE_LINE, G.MODE_FIRST):
        want = {i: o.scan_window(b, mode=mode) for i, b in enumerate(inputs)}
        outs = []
        for force in ("0", "1", "1", "1"):
            os.environ["GSCAN_CHAIN"] = force
            r = ctx.scan(p, inputs, mode=mode)
            st = ctx.stats()
            got = {}
            for fid, s, l in zip(r["file_id"].tolist(), r["start"].tolist(), r["match_len"].tolist()):
                got.setdefault(fid, []).append((s, l))
            wrong = [i for i in range(len(inputs)) if got.get(i, []) != want[i]]
            outs.append((force, len(r), st["vm_limit_hit"], st["total_launches"], wrong[:4]))
        flag = any(w for f, n, lim, nl, w in outs if not lim)
        if flag or len({(n, lim) for f, n, lim, nl, w in outs}) > 1:
            bad += flag
            print(("WRONG " if flag else "note  ") + repr(pat), "mode", mode, outs, flush=True)
            if flag:
                f, n, lim, nl, w = [x for x in outs if x[4] and not x[2]][0]
                i = w[0]
                print("   unit", i, "len", len(inputs[i]), "want tail", want[i][-3:], flush=True)
print("VM_PAR_DIAG", "OK" if bad == 0 else "WRONG x%d" % bad, flush=True)
