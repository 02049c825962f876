#!/usr/bin/env python
"""What the shipped library's SASS says about the primitives it uses (B200_PROFILING.md, "What proves a Blackwell-native
kernel"): counts of the bulk-TMA copy (UBLKCP), the mbarrier operations (SYNCS.*) and -- as a negative check, this path has
no dense contraction -- tensor-core instructions, per scan-kernel family.  Writes profiles/r02_sass_check.txt.
Usage: python tools/sass_check.py [path/to/libgscan.so]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_counts(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode()
    per = collections.OrderedDict()
    cur = None
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            fam = "other"
            for key, label in (("FixedBEngine", "scan_kernel<FixedBEngine> (balanced pair filter)"), ("Fixed3Engine", "scan_kernel<Fixed3Engine> (triple filter)"),
                               ("FixedEngine", "scan_kernel<FixedEngine> (pair filter)"), ("HashEngine", "scan_kernel<HashEngine>"),
                               ("RunEngine", "scan_kernel<RunEngine>"), ("NullEngine", "scan_kernel<NullEngine> (TMA probe)")):
                if key in name and "scan_kernel" in name:
                    fam = label
                    break
            cur = per.setdefault(fam, collections.Counter())
            cur["kernels"] += 1
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["instructions"] += 1
        if op.startswith("UBLKCP"):
            cur["UBLKCP (cp.async.bulk, 1-D TMA)"] += 1
        elif op.startswith("SYNCS.ARRIVE"):
            cur["SYNCS.ARRIVE.TRANS64 (mbarrier arrive / expect_tx)"] += 1
        elif op.startswith("SYNCS.PHASECHK"):
            cur["SYNCS.PHASECHK.TRANS64.TRYWAIT (mbarrier try_wait)"] += 1
        elif op.startswith(("UTC", "HMMA", "IMMA", "HGMMA", "QGMMA", "LDTM", "STTM")):
            cur["tensor-core / TMEM instructions"] += 1
    return arch, per


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "grab_b200", "libgscan.so")
    arch, per = sass_counts(lib)
    lines = ["SASS check of %s (cuobjdump -sass), architectures: %s" % (os.path.relpath(lib, ROOT), ", ".join(sorted(arch))), ""]
    for fam, c in per.items():
        lines.append("%s: %d kernel(s), %d instructions" % (fam, c["kernels"], c["instructions"]))
        for k, v in c.items():
            if k not in ("kernels", "instructions"):
                lines.append("    %-58s %d" % (k, v))
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    lines += ["", "total: UBLKCP %d, SYNCS.ARRIVE %d, SYNCS.PHASECHK %d, tensor-core %d" % (
        tot["UBLKCP (cp.async.bulk, 1-D TMA)"], tot["SYNCS.ARRIVE.TRANS64 (mbarrier arrive / expect_tx)"],
        tot["SYNCS.PHASECHK.TRANS64.TRYWAIT (mbarrier try_wait)"], tot["tensor-core / TMEM instructions"])]
    text = "\n".join(lines) + "\n"
    if len(sys.argv) <= 2:
        open(os.path.join(ROOT, "profiles", "r02_sass_check.txt"), "w").write(text)
    print(text)
    return tot


if __name__ == "__main__":
    main()
