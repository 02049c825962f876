#!/usr/bin/env python
"""Diagnostic: resolve time and launch count of one 64 MiB unit in LINE / ALL mode with the chain path on / off / automatic."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus
import grab_b200 as G
n = 64 << 20
ctx = G.Context(0)
host = corpus.synth_file(12, 5, n)
d = ctx.device_alloc(n)
ctx.h2d(d, host)
units = G.Context.device_units(d, 1, n)
for pat, mode in (("e", G.MODE_LINE), ("ee|e", G.MODE_ALL), ("[a-f][a-z]", G.MODE_LINE), ("foo|bar|baz|quux", G.MODE_LINE)):
    for env in (None, "1", "0"):
        if env is None:
            os.environ.pop("GSCAN_CHAIN", None)
        else:
            os.environ["GSCAN_CHAIN"] = env
        p = G.Pattern(pat)
        for rep in range(2):
            r = ctx.scan_units(p, units, mode)
            st = ctx.stats()
        print("%-18s mode %d GSCAN_CHAIN=%-4s matches %8d candidates %8d resolve %9.3f ms scan %7.3f ms launches %3d total %9.3f ms" % (
            pat, mode, env, len(r), st["n_candidates"], st["resolve_ms"], st["scan_kernel_ms"], st["total_launches"], st["total_ms"]), flush=True)
