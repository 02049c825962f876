"""Literal sets beyond the replicated-table range of the hashed engine (more than 512 slots: one table copy in shared
memory) and at its edges (256 / 384 / 512 slots), CUDA path vs the oracle.  Kept in its own (last) file: added after the
round's last GPU session, so its first run on a GPU is the driver's."""
import random

import pytest

import grab_b200 as G
import oracle_py as O

pytestmark = pytest.mark.gpu


def literal_set(n, seed, lo=3, hi=6):
    r = random.Random(seed)
    lits = set()
    while len(lits) < n:
        lits.add("".join(r.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(r.randint(lo, hi))))
    return sorted(lits, key=lambda s: r.random())  # preference order is part of the semantics


@pytest.mark.parametrize("n", [40, 90, 130, 200, 280])
def test_large_literal_sets(n):
    lits = literal_set(n, 1000 + n)
    pat = "|".join(lits)
    p = G.Pattern(pat)
    assert p.info["n_filter_tests"] < 0  # hashed engine
    o = O.Regex(pat)
    r = random.Random(n)
    bufs = []
    for k in range(24):
        words = [r.choice(lits) if r.random() < 0.5 else "".join(r.choice("abcxyz \n") for _ in range(r.randint(1, 7))) for _ in range(r.choice([3, 40, 400, 3000]))]
        sep = ["", " ", "\n"][k % 3]
        bufs.append(sep.join(words).encode())
    ctx = G.Context(0)
    try:
        for mode in (G.MODE_ALL, G.MODE_FIRST, G.MODE_LINE):
            res = ctx.scan(p, bufs, mode=mode)
            got = {}
            for fid, s, l in zip(res["file_id"], res["start"], res["match_len"]):
                got.setdefault(int(fid), []).append((int(s), int(l)))
            for i, b in enumerate(bufs):
                assert got.get(i, []) == o.scan_window(b, mode=mode), (n, mode, i)
    finally:
        ctx.close()
