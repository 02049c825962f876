"""The hashed engine's two paths (class-prefiltered sparse probing / dense probing) against the oracle on text that forces
every transition: printable noise (sparse), lowercase prose (the class test passes everywhere: dense fallback inside a
slice), alternating regions, bytes >= 0x80 next to matches (the SWAR class test lets carries run: it may only ever add
candidates), matches straddling every boundary the kernel has (32-byte lane pieces, 1 KiB blocks, 4 KiB slices, tiles, unit
ends), key hits whose tail differs from the literal, and GSCAN_HASH_PRE=0 (dense only) giving the same records."""
import random

import numpy as np
import pytest

import corpus
import grab_b200 as G
import oracle_py as O

pytestmark = pytest.mark.gpu

LOWER = "abcdefghijklmnopqrstuvwxyz"


def make_units(lits, seed):
    r = random.Random(seed)
    rng = np.random.default_rng(seed)
    units = []

    def noise(n):
        return rng.integers(32, 127, n, dtype=np.uint8)

    def prose(n):
        a = rng.integers(ord("a"), ord("z") + 1, n, dtype=np.uint8)
        a[rng.random(n) < 0.15] = 32
        return a

    def plant(a, where, lit):
        b = lit.encode()
        if 0 <= where and where + len(b) <= len(a):
            a[where:where + len(b)] = np.frombuffer(b, dtype=np.uint8)

    # 1. noise with literals planted across every boundary
    for n in (40000, 65536 + 77, 3 * 65536 + 5):
        a = noise(n)
        for edge in (32, 1024, 4096, 65536, 2 * 65536):
            for d in range(-6, 3):
                if edge + d + 8 < n and r.random() < 0.7:
                    plant(a, edge + d, r.choice(lits))
        plant(a, n - 3, lits[0][:3] if len(lits[0]) >= 3 else lits[0])
        plant(a, n - len(lits[1]), lits[1])
        plant(a, 0, lits[2])
        units.append(a)
    # 2. prose (dense everywhere) and prose / noise alternating every few hundred bytes .. few KiB
    units.append(prose(50000))
    for period in (300, 1500, 5000, 20000):
        parts = []
        for i in range(8):
            m = r.randint(period // 2, period * 2)
            parts.append(prose(m) if i % 2 else noise(m))
            if r.random() < 0.8:
                plant(parts[-1], r.randint(0, max(0, m - 8)), r.choice(lits))
        units.append(np.concatenate(parts))
    # 3. high bytes right before / inside / after literals, and literal prefixes with a wrong tail
    a = noise(30000)
    for k in range(300):
        pos = r.randint(2, len(a) - 12)
        lit = r.choice(lits)
        plant(a, pos, lit)
        kind = k % 5
        if kind == 0:
            a[pos - 1] = r.choice([0x80, 0xff, 0xe0, 0x9f])
        elif kind == 1 and len(lit) > 3:
            a[pos + len(lit) - 1] = r.choice([0x80, 0xff, ord("{"), ord("`")])
        elif kind == 2:
            a[pos + len(lit)] = r.choice([0x80, 0xff, 0xc3])
        elif kind == 3 and len(lit) > 3:
            a[pos + 3] = ord(r.choice(LOWER))
    units.append(a)
    units.append(rng.integers(0, 256, 20000, dtype=np.uint8))
    # 4. tiny units and units made of back-to-back literals
    for n in (1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65, 1023, 1024, 1025):
        a = noise(n)
        if n >= 5:
            plant(a, n - 5, "".join(r.choice(LOWER) for _ in range(2)) + lits[3][:3])
        units.append(a)
    units.append(np.frombuffer(("".join(r.choice(lits) for _ in range(4000))).encode(), dtype=np.uint8).copy())
    units.append(np.frombuffer((" ".join(r.choice(lits) for _ in range(3000))).encode(), dtype=np.uint8).copy())
    return [u.tobytes() for u in units]


def literal_set(n, seed, lo, hi, alphabet=LOWER):
    r = random.Random(seed)
    s = set()
    while len(s) < n:
        s.add("".join(r.choice(alphabet) for _ in range(r.randint(lo, hi))))
    return sorted(s, key=lambda x: r.random())


SETS = [
    ("lits100", lambda: corpus.literals100().split("|")),
    ("len2to6", lambda: literal_set(60, 5, 2, 6)),
    ("digits", lambda: literal_set(40, 6, 3, 5, "0123456789")),
    ("prefixes", lambda: ["abc", "abcd", "abcde", "abd", "xyz", "xyzzy", "xy", "qrs", "qrst"] + literal_set(30, 7, 3, 5)),
    ("upperlower", lambda: literal_set(50, 8, 3, 5, "ABCDEFGHIJKLMNOPQRSTUVWXYZ" + LOWER)),
]


@pytest.mark.parametrize("name", [s[0] for s in SETS])
def test_hash_paths_vs_oracle(name, monkeypatch):
    lits = dict(SETS)[name]()
    pat = "|".join(lits)
    units = make_units(lits, 99)
    o = O.Regex(pat)
    want = {m: [o.scan_window(b, mode=m) for b in units] for m in (G.MODE_ALL, G.MODE_FIRST, G.MODE_LINE)}
    for pre in ("1", "0"):
        monkeypatch.setenv("GSCAN_HASH_PRE", pre)
        p = G.Pattern(pat)
        assert p.info["n_filter_tests"] < 0  # hashed engine
        ctx = G.Context(0)
        try:
            for mode in (G.MODE_ALL, G.MODE_FIRST, G.MODE_LINE):
                res = ctx.scan(p, units, mode=mode)
                got = {}
                for fid, s, l in zip(res["file_id"], res["start"], res["match_len"]):
                    got.setdefault(int(fid), []).append((int(s), int(l)))
                for i in range(len(units)):
                    assert got.get(i, []) == want[mode][i], (name, pre, mode, i, len(units[i]))
        finally:
            ctx.close()
