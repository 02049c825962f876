#!/usr/bin/env python
"""Times the two command lines on the same tmpfs tree: grab-b200 (GPU) and the unmodified reference (host cores).
Usage: python tools/cli_bench.py [n_files]"""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
d = bench.materialise_sample(n)
try:
    nbytes = n * bench.FILE_LEN
    ours = os.path.join(ROOT, "grab_b200", "bin", "grab-b200")
    for name, cmd in (("grab-b200 (1 thread)", [ours, "-r", "-O", "-l", bench.PATTERN, d]),
                      ("grab-b200 -n 8", [ours, "-n", "8", "-r", "-O", "-l", bench.PATTERN, d]),
                      ("grab-b200 -n 32", [ours, "-n", "32", "-r", "-O", "-l", bench.PATTERN, d]),
                      ("grab_ref (1 thread)", [bench.REF_BIN, "-r", "-O", "-l", bench.PATTERN, d]),
                      ("grab_ref -n 32", [bench.REF_BIN, "-n", "32", "-r", "-O", "-l", bench.PATTERN, d]),
                      ("grab_ref -n 128", [bench.REF_BIN, "-n", "128", "-r", "-O", "-l", bench.PATTERN, d])):
        best, out = None, None
        for _ in range(3):
            t0 = time.perf_counter()
            p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            out = p.stdout
        print("%-24s %7.3f s  %6.2f GB/s  rc=%d  lines=%d  sorted-md5=%s" % (
            name, best, nbytes / best / 1e9, p.returncode, out.count(b"\n"),
            __import__("hashlib").md5(b"\n".join(sorted(out.split(b"\n")))).hexdigest()[:12]), flush=True)
finally:
    shutil.rmtree(d, ignore_errors=True)
