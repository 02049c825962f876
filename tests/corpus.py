"""Deterministic test/bench corpora (host side, numpy).

`synth_file` is the bit-identical host twin of the device generator in
grab_b200/csrc/corpus_gen.cu (same constants, same arithmetic): any file of the
64 GiB bench corpus can be regenerated here for parity without a D2H of the corpus
("Corpus generator").

`b3_corpus` is SURVEY.md Appendix B.3's numpy recipe (ties the engine to numbers
actually produced by the reference binary during the survey).
"""
import numpy as np

M64 = (1 << 64) - 1
K_SEED = 0x9E3779B97F4A7C15
K_FILE = 0xD1B54A32D192ED03
K_H2 = 0xA5A5A5A5A5A5A5A5
K_NEEDLE = 0x8CB92BA72F3D8DD7


def _mix64_np(x):
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def mix64(x):
    x &= M64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & M64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & M64
    x ^= x >> 31
    return x


def needle_offset(seed, file_id, file_len, needle_len):
    """Where a planted needle starts inside `file_id` (same arithmetic on the device)."""
    span = file_len - needle_len
    if span <= 0:
        return 0
    return mix64((seed * K_NEEDLE + file_id) & M64) % span


def synth_file(seed, file_id, length, needle=None, needle_every=0):
    """Bytes of synthetic file `file_id`: printable ASCII 0x20..0x7E, '\\n' with p = 3/256.

    needle/needle_every: files with file_id % needle_every == needle_every // 2 get `needle`
    written at needle_offset(...)."""
    nblk = (length + 7) // 8
    j = np.arange(nblk, dtype=np.uint64)
    base = np.uint64(((seed * K_SEED) + (file_id * K_FILE)) & M64)
    with np.errstate(over="ignore"):
        h1 = _mix64_np(base + j)
        h2 = _mix64_np(h1 ^ np.uint64(K_H2))
    b1 = h1.view(np.uint8).reshape(-1, 8).astype(np.uint32)  # little endian: byte k = bits 8k..8k+7
    b2 = h2.view(np.uint8).reshape(-1, 8)
    out = (0x20 + ((b1 * 95) >> 8)).astype(np.uint8)
    out[b2 < 3] = 10
    out = out.reshape(-1)[:length].copy()
    if needle and needle_every and file_id % needle_every == needle_every // 2 and length > len(needle):
        o = needle_offset(seed, file_id, length, len(needle))
        out[o:o + len(needle)] = np.frombuffer(needle, dtype=np.uint8)
    return out


def synth_range(seed, file_id, off, length):
    """Bytes [off, off + length) of synthetic file `file_id` without needles (off a multiple of 8): the generator is
    counter based per 8-byte block, so a big file can be regenerated piecewise (bench.py, the 256 MiB config)."""
    assert off % 8 == 0
    nblk = (length + 7) // 8
    j = np.arange(off // 8, off // 8 + nblk, dtype=np.uint64)
    base = np.uint64(((seed * K_SEED) + (file_id * K_FILE)) & M64)
    with np.errstate(over="ignore"):
        h1 = _mix64_np(base + j)
        h2 = _mix64_np(h1 ^ np.uint64(K_H2))
    b1 = h1.view(np.uint8).reshape(-1, 8).astype(np.uint32)
    b2 = h2.view(np.uint8).reshape(-1, 8)
    out = (0x20 + ((b1 * 95) >> 8)).astype(np.uint8)
    out[b2 < 3] = 10
    return out.reshape(-1)[:length].copy()


def b3_corpus(n=256 << 20, seed=12345):
    """SURVEY.md Appendix B.3 (numpy 2.x default_rng bit stream)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0x20, 0x7F, size=n, dtype=np.uint8)
    nl = rng.random(n) < (1 / 80)
    a[nl] = 10
    return a


def literals100(seed=7):
    """the 100-literal alternation (python `random`, seed 7)."""
    import random
    r = random.Random(seed)
    lits = ["".join(r.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(r.randint(3, 5))) for _ in range(100)]
    return "|".join(lits)


def overlap_file(chunk=1 << 25, size=40 << 20):
    """'.' with '\\n' every 64 B, NEEDLE at 1000, C-4096+100, C-3, EOF-6."""
    a = np.full(size, ord("."), dtype=np.uint8)
    a[63::64] = 10
    nd = np.frombuffer(b"NEEDLE", dtype=np.uint8)
    for o in (1000, chunk - 4096 + 100, chunk - 3, size - 6):
        a[o:o + 6] = nd
    return a


def overlap_run_file(chunk=1 << 25, size=40 << 20):
    """Appendix B.2 second case: a 60-byte run of 'R' starting at C-4096-10."""
    a = np.full(size, ord("."), dtype=np.uint8)
    a[63::64] = 10
    o = chunk - 4096 - 10
    a[o:o + 60] = ord("R")
    return a
