#!/usr/bin/env python
"""Times the two command lines on the same tmpfs tree: grab-b200 (GPU) and the unmodified reference (host cores),
then on ONE file of the same bytes (the reference has a single thread there; grab-b200 spreads the file's windows
over lanes / GPUs), then -- `patterns` given -- the other BASELINE patterns over the tree (the literal is the reference's
best case: its PCRE2-JIT search is a memchr; alternations, class runs and literal sets are where its cores are busy).
Usage: python tools/cli_bench.py [n_files] [n_gpus] [patterns]"""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ngpu = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = bench.materialise(bench.baseline_configs()[1], n)


def timed(name, cmd, nbytes, env=None, reps=3):
    e = dict(os.environ)
    e.update(env or {})
    best, out = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        out = p.stdout
    print("%-58s %7.3f s  %6.2f GB/s  rc=%d  lines=%d  sorted-md5=%s" % (
        name, best, nbytes / best / 1e9, p.returncode, out.count(b"\n"),
        __import__("hashlib").md5(b"\n".join(sorted(out.split(b"\n")))).hexdigest()[:12]), flush=True)


try:
    nbytes = n * bench.FILE_LEN
    ours = os.path.join(ROOT, "grab_b200", "bin", "grab-b200")
    tree = ["-r", "-O", "-l", bench.PATTERN, d]
    timed("grab-b200 (1 lane)", [ours] + tree, nbytes)
    timed("grab-b200 (2 lanes)", [ours] + tree, nbytes, dict(GRAB_B200_LANES="2"))
    if ngpu > 1:
        timed("grab-b200 (%d gpus)" % ngpu, [ours] + tree, nbytes, dict(GRAB_B200_NDEV=str(ngpu)))
        timed("grab-b200 (%d gpus x 2 lanes)" % ngpu, [ours] + tree, nbytes, dict(GRAB_B200_NDEV=str(ngpu), GRAB_B200_LANES="2"))
    timed("grab-b200 -n 8", [ours, "-n", "8"] + tree, nbytes, dict(GRAB_B200_NDEV=str(ngpu)))
    timed("grab-b200 -n 32", [ours, "-n", "32"] + tree, nbytes, dict(GRAB_B200_NDEV=str(ngpu)))
    timed("grab_ref (1 thread)", [bench.REF_BIN] + tree, nbytes, reps=1)
    timed("grab_ref -n 32", [bench.REF_BIN, "-n", "32"] + tree, nbytes)
    timed("grab_ref -n 128", [bench.REF_BIN, "-n", "128"] + tree, nbytes)
    # one file: same bytes concatenated (tmpfs); 1 GiB windows with 4 KiB overlap, each its own batch
    big = os.path.join(d, "big.bin")
    with open(big, "wb") as f:
        for root, _, names in sorted(os.walk(d)):
            for nm in sorted(names):
                if nm != "big.bin":
                    with open(os.path.join(root, nm), "rb") as g:
                        shutil.copyfileobj(g, f, 1 << 24)
    one = ["-O", "-l", bench.PATTERN, big]
    timed("one file: grab-b200 (1 lane)", [ours] + one, nbytes)
    timed("one file: grab-b200 (2 lanes)", [ours] + one, nbytes, dict(GRAB_B200_LANES="2"))
    if ngpu > 1:
        timed("one file: grab-b200 (%d gpus)" % ngpu, [ours] + one, nbytes, dict(GRAB_B200_NDEV=str(ngpu)))
    timed("one file: grab_ref", [bench.REF_BIN] + one, nbytes, reps=1)
    os.unlink(big)
    if len(sys.argv) > 3:
        for cfg in bench.baseline_configs()[2:]:
            pat = cfg["pattern"]
            args = ["-r", "-O", "-l", pat, d]
            tag = cfg["key"] + " " + (pat if len(pat) < 24 else pat[:16] + "...")
            timed(tag + ": grab-b200 (2 lanes)", [ours] + args, nbytes, dict(GRAB_B200_LANES="2"), reps=2)
            timed(tag + ": grab_ref -n 32", [bench.REF_BIN, "-n", "32"] + args, nbytes, reps=1)
            timed(tag + ": grab_ref -n 128", [bench.REF_BIN, "-n", "128"] + args, nbytes, reps=1)
finally:
    shutil.rmtree(d, ignore_errors=True)
