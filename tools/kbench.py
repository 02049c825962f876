#!/usr/bin/env python
"""Kernel bench (tuning): scan-kernel GB/s of the BASELINE patterns on device-resident corpora of the BASELINE shapes,
with an oracle spot check of every result, for one build of the library (GSCAN_LIB=... picks a tuning variant).

  python tools/kbench.py [--gib 16] [--only alt4,run16,...] [--label name]

One JSON line per case on stdout: {"label", "case", "kernel_gbs", "kernel_ms_best", "kernel_ms_med", "resolve_ms", "call_ms",
"matches", "parity"}.  Needs a B200."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import corpus  # noqa: E402
import grab_b200 as G  # noqa: E402
import oracle_py as O  # noqa: E402

MiB, GiB = 1 << 20, 1 << 30


def cases(gib):
    n1 = int(gib * 1024)
    return [
        # name, pattern, literal, mode, seed, file_len, n_files, needle
        ("literal", "foobardoesexist", True, G.MODE_ALL, 2, MiB, n1, b"foobardoesexist"),
        ("alt4", "foo|bar|baz|quux", False, G.MODE_ALL, 2, MiB, n1, b"foobardoesexist"),
        ("run16", "[A-Za-z0-9_]{16,}", False, G.MODE_ALL, 2, MiB, n1, b"foobardoesexist"),
        ("lits100", corpus.literals100(), False, G.MODE_ALL, 2, MiB, n1, b"foobardoesexist"),
        ("lits100_16k", corpus.literals100(), False, G.MODE_ALL, 5, 16384, int(gib * 65536), None),
        ("run16_16k", "[A-Za-z0-9_]{16,}", False, G.MODE_ALL, 5, 16384, int(gib * 65536), None),
        ("literal_16k", "foobardoesexist", True, G.MODE_ALL, 5, 16384, int(gib * 65536), None),
        ("c1_256m_first", "foobardoesnotexist", True, G.MODE_FIRST, 1, 256 * MiB, 1, None),
        ("icase", "(?i)linus", False, G.MODE_ALL, 2, MiB, n1, b"foobardoesexist"),
        ("lits8", "alpha|bravo|charlie|delta|echo|foxtrot|golf|hotel", False, G.MODE_ALL, 2, MiB, n1, b"foobardoesexist"),
        ("run4", "[0-9]{4,}", False, G.MODE_ALL, 2, MiB, n1, b"foobardoesexist"),
        ("lit2", "qz", False, G.MODE_ALL, 2, MiB, n1, b"foobardoesexist"),
    ]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=16.0)
    ap.add_argument("--only", default="")
    ap.add_argument("--label", default=os.environ.get("GSCAN_LIB", "default"))
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--check-files", type=int, default=6)
    a = ap.parse_args()
    only = set(x for x in a.only.split(",") if x)
    ctx = G.Context(0)
    d = ctx.device_alloc(int(a.gib * GiB))
    resident = None
    for name, pat, lit, mode, seed, flen, n, needle in cases(a.gib):
        if only and name not in only:
            continue
        key = (seed, flen, n)
        if resident != key:
            ctx.synth_corpus(d, seed, 0, n, flen, needle=needle, needle_every=64 if needle else 0)
            resident = key
        p = G.Pattern(pat, literal=lit)
        b = ctx.batch_create(G.Context.device_units(d, n, flen))
        ks, rs, ts = [], [], []
        for i in range(a.reps):
            r = ctx.batch_scan(p, b, mode)
            st = ctx.stats()
            if i:
                ks.append(st["scan_kernel_ms"]); rs.append(st["resolve_ms"]); ts.append(st["total_ms"])
        # oracle spot check on a few regenerated files (the first, the last, some in between)
        ok = True
        o = O.Regex(pat, literal=lit)
        omode = {G.MODE_ALL: O.MODE_ALL, G.MODE_FIRST: O.MODE_FIRST}[mode]
        ids = sorted(set([0, n - 1] + [int(x) for x in np.linspace(0, n - 1, a.check_files)])) if flen <= 4 * MiB else []
        for f in ids:
            data = corpus.synth_file(seed, f, flen, needle, 64 if needle else 0).tobytes()
            want = [s for s, _ in o.scan_window(data, mode=omode)]
            lo, hi = np.searchsorted(r["file_id"], f, "left"), np.searchsorted(r["file_id"], f, "right")
            if r["start"][lo:hi].tolist() != want:
                ok = False
        nbytes = n * flen
        print(json.dumps({"label": a.label, "case": name, "kernel_gbs": round(nbytes / (min(ks) * 1e-3) / 1e9, 1),
                          "kernel_ms_best": round(min(ks), 4), "kernel_ms_med": round(float(np.median(ks)), 4),
                          "resolve_ms": round(float(np.median(rs)), 3), "call_ms": round(float(np.median(ts)), 3),
                          "matches": int(len(r)), "engine": p.info["engine"], "tests": p.info["n_filter_tests"],
                          "parity": "ok" if ok else "MISMATCH", "files_checked": len(ids)}), flush=True)
        b.free()
    probe = min(ctx.read_probe(d, int(a.gib * GiB))[0] for _ in range(3))
    print(json.dumps({"label": a.label, "case": "read_probe", "kernel_gbs": round(a.gib * GiB / (probe * 1e-3) / 1e9, 1)}), flush=True)
    ctx.device_free(d)
    ctx.close()


if __name__ == "__main__":
    main()
