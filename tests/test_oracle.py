"""The oracle port (oracle/grab_oracle.c) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import base64
import hashlib
import json
import os

import pytest

import corpus
import oracle_py as O

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))
MINLEN = json.load(open(os.path.join(HERE, "golden", "minlen.json")))
BIG = json.load(open(os.path.join(HERE, "golden", "big.json")))


def flags_to_kwargs(flags):
    return dict(offsets="-O" in flags, line="-l" not in flags, single="-s" in flags)


@pytest.mark.parametrize("case", KAT["cases"], ids=lambda c: c["name"])
def test_kat_stdout(case):
    data = base64.b64decode(case["input"])
    want = base64.b64decode(case["stdout"])
    re = O.Regex(case["pattern"])
    got = re.grab(data, **flags_to_kwargs(case["flags"]))
    assert got == want


@pytest.mark.parametrize("case", KAT["multi"], ids=lambda c: c["name"])
def test_multi_path_prefix(case):
    re = O.Regex(case["pattern"])
    files = {fn: base64.b64decode(d) for fn, d in case["files"]}
    got = b"".join(re.grab(files[p], path=p, **flags_to_kwargs(case["flags"])) for p in case["paths"])
    assert got == base64.b64decode(case["stdout"])


@pytest.mark.parametrize("case", KAT["recursive"], ids=lambda c: c["name"])
def test_recursive_sorted(case):
    re = O.Regex(case["pattern"])
    lines = []
    for fn, d in case["tree"].items():
        out = re.grab(base64.b64decode(d), path=fn, **flags_to_kwargs(case["flags"]))
        lines += [l for l in out.split(b"\n") if l]
    assert sorted(lines) == [base64.b64decode(l) for l in case["sorted_lines"]]


@pytest.mark.parametrize("ent", MINLEN, ids=lambda e: e["pattern"][:24])
def test_minlen_matches_pcre2(ent):
    if not ent["compiles"]:
        with pytest.raises(O.OracleError):
            O.Regex(ent["pattern"])
        return
    try:
        re = O.Regex(ent["pattern"])
    except O.OracleError:
        pytest.skip("construct not modelled by the oracle port")
    assert re.minlen == ent["minlen"]
    assert re.captures == ent["captures"]


def test_scan_window_modes():
    t1 = b"xxfoobarxx\nbazquux foo\nnothing\nfoo"
    re = O.Regex("foo|bar|baz|quux")
    assert [s for s, _ in re.scan_window(t1)] == [2, 5, 11, 14, 19, 31]
    assert [s for s, _ in re.scan_window(t1, mode=O.MODE_LINE)] == [2, 11, 31]
    assert [s for s, _ in re.scan_window(t1, mode=O.MODE_FIRST)] == [2]
    assert re.scan_window(t1, base_off=1000)[0] == (1002, 3)
    assert O.Regex("(foo)").scan_window(t1) == []
    assert [s for s, _ in O.Regex("(foo)").scan_window(t1, strict_q2=False)] == [2, 19, 31]
    with pytest.raises(O.OracleError):
        O.Regex("x*").scan_window(b"abc")


@pytest.mark.parametrize("ent", BIG["overlap"], ids=lambda e: e["gen"] + str(len(e["flags"])))
def test_chunk_overlap(ent):
    data = getattr(corpus, ent["gen"])().tobytes()
    chunk = 1 << 30
    for f in ent["flags"]:
        if f == "-L":
            chunk = max(chunk >> 1, 1 << 25)
    re = O.Regex(ent["pattern"])
    assert [s for s, _ in re.scan_file(data, chunk_size=chunk)] == ent["offsets"]


@pytest.mark.parametrize("ent", BIG["synth"], ids=lambda e: e["pattern"][:16])
def test_synth_files(ent):
    re = O.Regex(ent["pattern"])
    for fid, want in ent["offsets"].items():
        data = corpus.synth_file(ent["seed"], int(fid), ent["file_len"], ent["needle"].encode(), ent["needle_every"])
        assert [s for s, _ in re.scan_window(data.tobytes())] == want


def test_b3_32mib_prefix():
    a = corpus.b3_corpus(32 << 20)
    # NB: default_rng(12345).integers(size=N) is a prefix-stable stream only for the same N, so
    # regenerate what make_golden did: the 32 MiB file is the first 32 MiB of the 256 MiB one.
    full_needed = hashlib.md5(a.tobytes()).hexdigest()
    data = None
    for ent in BIG["b3_32"]:
        if data is None:
            data = corpus.b3_corpus()[: 32 << 20].tobytes()
        re = O.Regex(ent["pattern"])
        offs = [s for s, _ in re.scan_window(data)]
        txt = "".join("%d\n" % o for o in offs).encode()
        assert len(offs) == ent["n"], ent["pattern"][:20]
        assert hashlib.md5(txt).hexdigest() == ent["md5"], ent["pattern"][:20]
    del full_needed
