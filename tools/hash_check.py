#!/usr/bin/env python
"""Small, quick parity check of the hashed engine's sparse and dense paths (GPU; a subset of tests/test_gpu_hash_sparse.py
that prints as it goes, for a first run of a changed kernel under a short timeout)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import grab_b200 as G  # noqa: E402
import oracle_py as O  # noqa: E402
import test_gpu_hash_sparse as T  # noqa: E402


def main():
    bad = 0
    for name, mk in T.SETS:
        lits = mk()
        pat = "|".join(lits)
        units = T.make_units(lits, 99)
        o = O.Regex(pat)
        for pre in ("1", "0"):
            os.environ["GSCAN_HASH_PRE"] = pre
            p = G.Pattern(pat)
            ctx = G.Context(0)
            for mode in (G.MODE_ALL, G.MODE_LINE):
                t0 = time.time()
                res = ctx.scan(p, units, mode=mode)
                got = {}
                for fid, s, l in zip(res["file_id"], res["start"], res["match_len"]):
                    got.setdefault(int(fid), []).append((int(s), int(l)))
                wrong = [i for i, b in enumerate(units) if got.get(i, []) != o.scan_window(b, mode=mode)]
                bad += len(wrong)
                print(name, "pre", pre, "mode", mode, "records", len(res), "wrong units", wrong[:8], "%.2fs" % (time.time() - t0), flush=True)
            ctx.close()
    print("HASH_CHECK", "OK" if bad == 0 else "MISMATCH", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
