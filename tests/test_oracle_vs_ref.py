"""Pins the oracle port to the UNMODIFIED reference binary (oracle/_ref/grab_ref: grab master built against the PCRE2
shim) on seeded random patterns x random inputs, and to GNU grep -P (same libpcre2) as an independent second opinion.
CPU only; skipped where the reference binary is not built (it needs /root/reference at build time)."""
import os
import random
import shutil
import subprocess

import pytest

import oracle_py as O
from test_gpu_random_patterns import gen_pattern

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "grab_ref")

needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/grab_ref not built")


def ref_offsets(pattern, path, flags=("-O", "-l")):
    p = subprocess.run([REF] + list(flags) + [pattern, path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=20)
    assert p.returncode == 0, p.stderr
    return p.stdout


def random_inputs(rnd, n):
    out = []
    for k in range(n):
        alpha = [b"abc", b"abcx \n", b"ab", b"abc abc\n\n", b"aabbcc_x1 \t\n"][k % 5]
        ln = rnd.choice([0, 1, 2, 3, 7, 16, 33, 64, 130, 300, 513])
        out.append(bytes(rnd.choice(alpha) for _ in range(ln)))
    return out


@needs_ref
def test_random_patterns_stdout_identical(tmp_path):
    rnd = random.Random(20260924)
    inputs = random_inputs(rnd, 10)
    paths = []
    for i, b in enumerate(inputs):
        p = tmp_path / ("in%d" % i)
        p.write_bytes(b)
        paths.append(str(p))
    checked = 0
    for _ in range(150):
        pat = gen_pattern(rnd)
        try:
            o = O.Regex(pat)
        except O.OracleError:
            continue
        if o.nullable:  # the reference never terminates on these (Q4)
            continue
        try:
            for flags, kw in ((("-O", "-l"), dict(offsets=True, line=False)), ((), dict()), (("-s", "-O"), dict(offsets=True, single=True))):
                for path, data in zip(paths, inputs):
                    assert o.grab(data, **kw) == ref_offsets(pat, path, flags), (pat, flags, len(data))
        except (O.OracleError, subprocess.TimeoutExpired):
            continue  # pathological backtracking: either side hit its match limit / time budget
        checked += 1
    assert checked >= 60


OPTION_PATTERNS = [r"(?U)a+b", r"(?U)a+?b", r"(?U)a{2,}", r"(?U)\w+ ", r"(?U:a+)b", r"(?U)a*b", r"(?U)(?:ab)+c", r"a(?U)b+c?", r"(?U)a++b",
                   r"(?x) a b c", "(?x)a b # comment\n c", r"(?x)a\ b", r"(?x)a[ ]b", r"(?x)a +b", r"(?xi)A B", r"(?x)a b | b c", r"a(?x) b c",
                   r"(?x)a{2} b", r"(?x: a b ) c", r"(?x) [ab] {2} c", "(?x)# only a comment\nab",
                   r"a(?#hello)b", r"(?#c)a|b(?#d)c", r"a(?#x)+b"]


@needs_ref
def test_inline_options_ungreedy_extended_comment(tmp_path):
    """(?U), (?x) and (?#...) as the reference's PCRE build treats them: identical stdout in all three output modes."""
    rnd = random.Random(5)
    inputs = random_inputs(rnd, 15) + [b"aaab ab\nfoo  bar # x\nabcabc aXb a\nb\n", b"a b c abc a  b\n", b"aab aaab abbc abc\n"]
    checked = 0
    for pat in OPTION_PATTERNS:
        o = O.Regex(pat)
        if o.nullable:
            continue
        for i, data in enumerate(inputs):
            path = tmp_path / ("in%d" % i)
            path.write_bytes(data)
            for flags, kw in ((("-O", "-l"), dict(offsets=True, line=False)), ((), dict()), (("-s", "-O"), dict(offsets=True, single=True))):
                assert o.grab(data, **kw) == ref_offsets(pat, str(path), flags), (pat, flags, data)
        checked += 1
    assert checked >= 20


@needs_ref
def test_minlen_quirk_q1_against_reference(tmp_path):
    # the strict '<' of grab.cc:175 for every length around minlen
    for pat, unit in (("abc", b"abc"), ("[ab]{4,}", b"abab"), ("ab|abcd", b"ab")):
        o = O.Regex(pat)
        for reps in range(0, 4):
            for tail in (b"", b"x", b"\n"):
                data = unit * reps + tail
                p = tmp_path / "f"
                p.write_bytes(data)
                assert o.grab(data, offsets=True, line=False) == ref_offsets(pat, str(p)), (pat, data)


@pytest.mark.skipif(shutil.which("grep") is None, reason="no grep")
def test_second_opinion_gnu_grep_P(tmp_path):
    """grep -P -a -b -o prints every non-overlapping match start like the reference's -O -l, except for the reference's
    quirks (Q1 tail, Q2 captures) and for matches that span '\\n' -- so: inputs end in '\\n', patterns cannot match '\\n'."""
    probe = subprocess.run(["grep", "-P", "-a", "-b", "-o", "a", "/dev/null"], stderr=subprocess.PIPE)
    if probe.returncode not in (0, 1):
        pytest.skip("grep -P unavailable")
    rnd = random.Random(7)
    pats = ["foo|bar|baz|quux", "[A-Za-z0-9_]{5,}", "qz", r"\d{2}-\d{2}", "(?i)ab+c", r"a[^b\n]*b", r"\bfoo\b", "x+y+?"]
    for pat in pats:
        o = O.Regex(pat)
        for _ in range(6):
            lines = []
            for _ in range(rnd.randint(1, 40)):
                lines.append(bytes(rnd.choice(b"abcfoqzxy019-_ Bbar") for _ in range(rnd.randint(0, 60))))
            data = b"\n".join(lines) + b"\n"
            p = tmp_path / "g"
            p.write_bytes(data)
            g = subprocess.run(["grep", "-P", "-a", "-b", "-o", pat, str(p)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            assert g.returncode in (0, 1), g.stderr
            want = [int(l.split(b":", 1)[0]) for l in g.stdout.split(b"\n") if l]
            got = [s for s, _ in o.scan_window(data)]
            assert got == want, (pat, data)


@needs_ref
def test_long_lines_line_mode_against_reference(tmp_path):
    """Line output on lines longer than the 511 bytes the reference prints behind a match (grab.cc:194-196): the next search
    resumes in the middle of the line -- for class runs possibly in the middle of a run, where PCRE then reports a match at
    the resume point itself.  The resolve pass's chain path for class runs (k_chain_next / k_chain_entry) is built on
    exactly this behaviour of the oracle: pinned here against the unmodified reference's stdout, line mode and -O."""
    rnd = random.Random(511)
    inputs = []
    for alpha, n, newline_every in ((b"ab", 700, 0), (b"ab", 3000, 0), (b"aab_ 1", 2500, 0), (b"abc ", 4000, 1300), (b"aab_1 \n", 3000, 0),
                                    (b"a", 1200, 0), (b"ab", 511, 0), (b"ab", 512, 0), (b"ab", 1023, 0), (b"ab", 1024, 0)):
        b = bytearray(rnd.choice(alpha) for _ in range(n))
        if newline_every:
            for i in range(newline_every, n, newline_every):
                b[i] = 10
        inputs.append(bytes(b))
    pats = ["[ab]{2,}", "[ab]{5,}", "a{3,}", "\\w{3,}", "[^b]{2,}", "[a_1]{2,}", "[ab\\n]{4,}", "ab", "aa|ab", "ab+a", "a[ab]*b", "\\w+ ", "a+"]
    checked = 0
    for pat in pats:
        o = O.Regex(pat)
        for i, data in enumerate(inputs):
            path = tmp_path / ("long%d" % i)
            path.write_bytes(data)
            for flags, kw in (((), dict()), (("-O",), dict(offsets=True)), (("-O", "-l"), dict(offsets=True, line=False))):
                assert o.grab(data, **kw) == ref_offsets(pat, str(path), flags), (pat, flags, i, len(data))
                checked += 1
    assert checked == len(pats) * len(inputs) * 3
