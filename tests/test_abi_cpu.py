"""CPU-only checks of the product library: it loads, exports every symbol include/gscan.h declares,
compiles patterns to the same minimum length as PCRE2 (fixtures from the reference's library),
rejects what the device engines do not serve, and FAILS LOUDLY without a GPU (no CPU fallback)."""
import ctypes
import json
import os
import re

import pytest

import grab_b200 as G

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MINLEN = json.load(open(os.path.join(HERE, "golden", "minlen.json")))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gscan.h")).read()
    declared = set(re.findall(r"\b(gscan_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(G.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "libgscan.so does not export %s" % name
    assert declared == set(G._PROTOS), "python prototypes out of sync with include/gscan.h"
    assert G.lib().gscan_abi_version() == 2


def test_flag_values_match_the_header():
    """The constants the ctypes face passes are the header's (unit flags, modes, compile flags)."""
    hdr = open(os.path.join(ROOT, "include", "gscan.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(GSCAN_[A-Z_0-9]+)\s+(\d+)u?\b", hdr)}
    assert defs["GSCAN_UNIT_DEVICE"] == G.UNIT_DEVICE and defs["GSCAN_UNIT_FD"] == G.UNIT_FD
    assert (defs["GSCAN_MODE_ALL"], defs["GSCAN_MODE_FIRST"], defs["GSCAN_MODE_LINE"]) == (G.MODE_ALL, G.MODE_FIRST, G.MODE_LINE)
    assert defs["GSCAN_LITERAL"] == G.LITERAL and defs["GSCAN_STRICT_REF"] == G.STRICT_REF
    u = G.Context.fd_units([(5, 4096, 100)], file_ids=[9])
    assert (int(u[0]["ptr"]), int(u[0]["len"]), int(u[0]["base_off"]), int(u[0]["file_id"]), int(u[0]["flags"])) == (5, 100, 4096, 9, G.UNIT_FD)


def test_struct_layouts():
    assert ctypes.sizeof(G.Unit) == 32 and G.UNIT_DTYPE.itemsize == 32
    assert ctypes.sizeof(G.Match) == 16 and G.MATCH_DTYPE.itemsize == 16


UNSUPPORTED_OK = ("not supported", "only supported", "empty string", "too many", "too common", "any byte", "needs more than")


@pytest.mark.parametrize("ent", MINLEN, ids=lambda e: e["pattern"][:24])
def test_minlen_equals_pcre2(ent):
    if not ent["compiles"]:
        with pytest.raises(G.GscanError):
            G.Pattern(ent["pattern"])
        return
    try:
        p = G.Pattern(ent["pattern"])
    except G.GscanError as e:
        # loud, explained rejection is the contract for constructs outside the device engines
        assert any(s in str(e) for s in UNSUPPORTED_OK), str(e)
        if ent["minlen"] == 0:
            assert "empty string" in str(e)
        return
    assert p.minlen == ent["minlen"]
    assert p.info["captures"] == ent["captures"]


def test_engine_selection_and_filter():
    i = G.Pattern("foobardoesnotexist").info
    assert i["engine"] == G.ENGINE_FIXED and i["n_sequences"] == 1 and i["n_filter_tests"] == 1
    assert i["minlen"] == 18 and i["maxlen"] == 18 and i["filter_delta"] in (1, 2, 3, 4)
    i = G.Pattern("foo|bar|baz|quux").info
    assert i["engine"] == G.ENGINE_FIXED and i["n_sequences"] == 3 and i["minlen"] == 3 and i["maxlen"] == 4  # bar|baz merge into ba[rz]
    # three exact byte pairs (fo, ba, qu) => the balanced pair kernel, third bytes (o, [rz], u) in its inline stage 2
    assert i["scan_kernel"] == G.KERNEL_BALANCED and i["n_filter_tests"] == 3 and i["filter_delta"] == 1
    assert G.Pattern("foobardoesexist", literal=True).info["scan_kernel"] == G.KERNEL_PAIR
    assert G.Pattern("[A-Za-z0-9_]{16,}").info["scan_kernel"] == G.KERNEL_RUN
    assert G.Pattern("(?i)linus|torvalds").info["scan_kernel"] in (G.KERNEL_TRIPLE, G.KERNEL_PAIR)
    import corpus
    assert G.Pattern(corpus.literals100()).info["scan_kernel"] == G.KERNEL_HASH
    assert 1 <= i["n_filter_tests"] <= 4
    i = G.Pattern("[A-Za-z0-9_]{16,}").info
    assert i["engine"] == G.ENGINE_RUN and i["minlen"] == 16 and i["maxlen"] == -1
    assert G.Pattern("(?i)linus").info["minlen"] == 5
    assert G.Pattern("a.c", literal=True).info["n_sequences"] == 1
    # Q2: strict mode + capturing group => nothing can ever be printed
    assert G.Pattern("(foo|bar)", strict_ref=True).info["engine"] == G.ENGINE_NONE
    assert G.Pattern("(foo|bar)").info["engine"] == G.ENGINE_FIXED
    # bounded repeats expand in backtracking order
    assert G.Pattern("a{2,4}").info["n_sequences"] == 3  # aaaa, aaa, aa (greedy order)
    assert G.Pattern("fo|foo|foobar").info["n_sequences"] == 1  # later alternatives are shadowed by fo
    assert G.Pattern("colou?r").info["n_sequences"] == 2
    # general patterns: VM engine, candidates from the leading bytes
    i = G.Pattern(r"foo\d+bar").info
    assert i["engine"] == G.ENGINE_VM and i["minlen"] == 7 and i["maxlen"] == -1
    assert G.Pattern("^foo").info["engine"] == G.ENGINE_VM and G.Pattern(r"\bfoo\b").info["minlen"] == 3
    assert G.Pattern(r"\w+@\w+\.com").info["engine"] == G.ENGINE_VM  # begins with a class run: run starts are the candidates
    # a match that can begin with (almost) any byte has no candidate filter: the VM walk tries the positions itself
    for pat in (".*foo", r"(?:\w|-)+@\w+", r"x*ab|\s?c"):
        i = G.Pattern(pat).info
        assert i["engine"] == G.ENGINE_VM and i["scan_kernel"] == G.KERNEL_NONE, pat
    # many alternatives: hashed engine (negative n_filter_tests = -(table slots))
    import corpus
    i = G.Pattern(corpus.literals100()).info
    assert i["engine"] == G.ENGINE_FIXED and i["n_filter_tests"] < 0 and i["minlen"] == 3 and i["maxlen"] == 5


@pytest.mark.parametrize("pat,frag", [
    ("x*", "empty string"), ("", "empty string"), ("a|", "empty string"),
    ("a*", "empty string"), ("(?:a*)+b", "not supported"),
    ("(", "missing )"), ("[a-", "missing terminating ]"), (r"\1", "back references"), (r"(?<=\d+)x", "not fixed length"), ("(?(1)a|b)", "not supported"),
    ("a{3,2}", "quantifier"),
])
def test_rejections_are_loud(pat, frag):
    with pytest.raises(G.GscanError) as ei:
        G.Pattern(pat)
    assert frag in str(ei.value)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(G.GscanError) as ei:
        G.Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_shipped_sass_uses_bulk_tma_and_mbarriers():
    """After build(): every scan-kernel family moves its slices with the 1-D bulk TMA copy (SASS UBLKCP) completing on an
    mbarrier (SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK...TRYWAIT), compiled for sm_100a only, and -- this path has no dense
    contraction -- without tensor-core instructions (tools/sass_check.py, profiles/r02_sass_check.txt)."""
    import shutil
    import sys
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not installed")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_check
    arch, per = sass_check.sass_counts(G.LIB_PATH)
    assert arch == {"sm_100a"}
    fams = [f for f in per if f.startswith("scan_kernel")]
    assert len(fams) >= 5
    for f in fams:
        c = per[f]
        assert c["UBLKCP (cp.async.bulk, 1-D TMA)"] >= c["kernels"], f
        assert c["SYNCS.ARRIVE.TRANS64 (mbarrier arrive / expect_tx)"] >= c["kernels"], f
        assert c["SYNCS.PHASECHK.TRANS64.TRYWAIT (mbarrier try_wait)"] >= c["kernels"], f
        assert c["tensor-core / TMEM instructions"] == 0, f
