#!/usr/bin/env python
"""Diagnostic (GPU): a match that starts on the last byte of a 4 KiB tile of a unit that ends 4 bytes later -- fresh context,
second scan, the unit alone, one big tile, simpler patterns.  GSCAN_LIB picks the library under test."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import grab_b200 as G  # noqa: E402
import oracle_py as O  # noqa: E402

rnd = random.Random(99)
inputs = []
for k in range(40):
    alpha = [b"abc", b"abcx \n", b"ab", b"abc abc\n\n", b"aabbcc_x1 \t\n"][k % 5]
    ln = rnd.choice([1, 2, 3, 7, 16, 17, 33, 64, 130, 511, 513, 1000, 2050, 4100, 9000])
    inputs.append(bytes(rnd.choice(alpha) for _ in range(ln)))
PAT = "[a-c]{2}[ab]*?\\w[ab]|[ab]{1,3}a{2,}|(?:a *)xa$"


def wrong(ctx, pat, bufs, mode=G.MODE_ALL):
    p, o = G.Pattern(pat), O.Regex(pat)
    r = ctx.scan(p, bufs, mode=mode)
    got = {}
    for fid, s, l in zip(r["file_id"].tolist(), r["start"].tolist(), r["match_len"].tolist()):
        got.setdefault(fid, []).append((s, l))
    out = []
    for i, b in enumerate(bufs):
        w = o.scan_window(b, mode=mode)
        if got.get(i, []) != w:
            out.append((i, len(b), len(got.get(i, [])), len(w), [x for x in w if x not in got.get(i, [])][:3], [x for x in got.get(i, []) if x not in w][:3]))
    return out, ctx.stats()["n_candidates"], ctx.stats()["vm_limit_hit"]


print("library", os.environ.get("GSCAN_LIB", "default"))
c = G.Context(0)
print("fresh context, all inputs     ", wrong(c, PAT, inputs))
print("same context, again           ", wrong(c, PAT, inputs))
print("unit 10 alone                 ", wrong(c, PAT, [inputs[10]]))
print("unit 10 alone, literal ccab   ", wrong(c, "ccab", [inputs[10]]))
print("unit 10 alone, [a-c]{2}a[ab]  ", wrong(c, "[a-c]{2}a[ab]", [inputs[10]]))
print("unit 10 + a 1-byte unit       ", wrong(c, PAT, [inputs[10], b"a"]))
os.environ["GSCAN_TILE_SHIFT"] = "16"
print("all inputs, 64 KiB tiles      ", wrong(c, PAT, inputs))
os.environ["GSCAN_TILE_SHIFT"] = "12"
print("all inputs, 4 KiB tiles forced", wrong(c, PAT, inputs))
del os.environ["GSCAN_TILE_SHIFT"]
c.close()
c = G.Context(0)
print("fresh context, other first    ", wrong(c, "ab+a", inputs), wrong(c, PAT, inputs))
c.close()
