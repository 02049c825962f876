#!/usr/bin/env python
"""bench.py -- GB/s scanned on BASELINE.json configs[1]: a literal (-S semantics) over a synthetic
corpus of 1 MiB files, 64 GiB per GPU, device resident; all offsets (-O -l).

  python bench.py [--gpus N --steps K --warmup W]                 our engine (1 process per GPU)
  python bench.py --impl reference [...]                          grab master (PCRE2-JIT shim) on the host cores

One JSON line on stdout (rank 0).  Contract notes:
  value    whole-job GB/s with the corpus already in HBM (inputs 64 GiB >> 126 MB L2, so no flush needed)
  e2e      the same metric through gscan_scan_batch() with HOST buffers: H2D + scan + D2H inside the timed region
  roofline scan kernel only: algorithmic bytes (1 B read per corpus byte) / CUDA-event kernel time
  cpu_baseline  the unmodified reference (oracle/_ref/grab_ref) timed on this box's host cores on a bounded sample
"""
import argparse
import ctypes
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GiB = 1 << 30
PATTERN = "foobardoesexist"      # literal; planted once per 64 files (= per 64 MiB) + natural hits (none expected)
NEEDLE_EVERY = 64
SEED = 2
FILE_LEN = 1 << 20
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "grab_ref")


def ctypes_memmove(dst, src, n):
    ctypes.memmove(ctypes.c_void_p(dst), ctypes.c_void_p(src), n)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe):
    one nvidia-smi process looping every 20 ms, started just before and stopped just after."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            time.sleep(0.15)  # first sample lands before the timed region starts
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return
        time.sleep(0.03)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0].decode()
        except Exception:
            out = ""
        for line in out.splitlines():
            r = [x.strip() for x in line.split(",")]
            if len(r) >= 6 and r[0].isdigit():
                self.rows.append(r)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[6]) for r in self.rows if len(r) > 6 and r[6].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows), "power_w_max": max(pw) if pw else None}


def _gen_file(args):
    import corpus
    d, fid = args
    corpus.synth_file(SEED, fid, FILE_LEN, PATTERN.encode(), NEEDLE_EVERY).tofile(os.path.join(d, "f%06d" % fid))
    return fid


def materialise_sample(n_files, from_device=None):
    """Writes the first n_files of the corpus under a tmp dir (tmpfs if present) and returns the path."""
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="gscan_bench_", dir=base)
    if from_device is not None:
        ctx, dptr = from_device
        for fid in range(n_files):
            ctx.d2h(dptr + fid * FILE_LEN, FILE_LEN).tofile(os.path.join(d, "f%06d" % fid))
    else:
        import multiprocessing as mp
        with mp.Pool(min(32, os.cpu_count() or 1)) as pool:
            list(pool.imap_unordered(_gen_file, [(d, fid) for fid in range(n_files)], chunksize=8))
    return d


def time_reference(sample_dir, n_files, cores, repeats=3):
    """grab_ref -n cores -r -O -l PATTERN dir, stdout -> /dev/null; returns best seconds."""
    args = [REF_BIN] + (["-n", str(cores)] if cores > 1 else []) + ["-r", "-O", "-l", PATTERN, sample_dir]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        p = subprocess.run(args, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            raise RuntimeError("grab_ref failed: %s" % p.stderr.decode()[:200])
        best = dt if best is None else min(best, dt)
    return best


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_cores():
    # grab pins thread i to CPU i (main.cc:200-215): every CPU 0..n-1 must be in our affinity mask
    try:
        aff = sorted(os.sched_getaffinity(0))
        n = 0
        while n < len(aff) and aff[n] == n:
            n += 1
        return max(1, n)
    except Exception:
        return usable_cores()


def run_reference(a, rank, world):
    if rank != 0:
        return
    line = {"impl": "reference", "metric": "GB/s scanned", "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic"}
    if not os.path.exists(REF_BIN):
        line["unavailable"] = "oracle/_ref/grab_ref is not built (needs /root/reference at build time)"
        print(json.dumps(line))
        return
    n_files = a.ref_files
    d = materialise_sample(n_files)
    try:
        cores = reference_cores()
        for _ in range(a.warmup):
            time_reference(d, n_files, cores, 1)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            time_reference(d, n_files, cores, 1)
        dt = (time.perf_counter() - t0) / a.steps
        one = time_reference(d, n_files, 1, 1) if not a.quick else None
    finally:
        shutil.rmtree(d, ignore_errors=True)
    nbytes = n_files * FILE_LEN
    gbs = nbytes / dt / 1e9
    line.update({
        "value": gbs, "ms_per_step": dt * 1e3,
        "config": {"workload": "BASELINE configs[1]: literal over a corpus of 1 MiB files; reference arm scans a bounded sample",
                   "pattern": PATTERN, "files": n_files, "file_bytes": FILE_LEN, "mode": "-n %d -r -O -l" % cores,
                   "engine": "grab master + PCRE2 10.42 JIT via oracle/shim (NOT hyperscan: no -H source or library available)"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "reference",
                         "sample": "%d files x 1 MiB on tmpfs, page cache warm; 1-core: %s GB/s" %
                                   (n_files, ("%.2f" % (nbytes / one / 1e9)) if one else "n/a")},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    print(json.dumps(line))


def e2e_all_ranks(scan, want_bytes, steps, world, rank, torch, dist, device, sync_all):
    """The timed end-to-end region at N > 1: every rank runs `scan()` (one gscan_scan_batch over its host sample) `steps`
    times between barriers.  Returns (seconds per step of the slowest rank, records per step on this rank, results equal
    `want_bytes` on every rank), or None when any rank could not run it (scan is None there, or it raised): every rank
    reaches every collective whatever happens on the others, so a local failure can drop the number but never hang the job."""
    failed, bad, n_rec = (1.0 if scan is None else 0.0), 0.0, 0
    sync_all()
    t0 = time.perf_counter()
    try:
        if not failed:
            out = None
            for _ in range(steps):
                out = scan()
            n_rec = len(out)
            bad = 0.0 if out.tobytes() == want_bytes else 1.0
    except Exception as ex:  # noqa: BLE001
        failed = 1.0
        sys.stderr.write("bench: e2e leg failed on rank %d: %r\n" % (rank, ex))
    if device == "cuda":
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    sync_all()
    t = torch.tensor([dt, bad, failed], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if float(t[2].item()) != 0.0:
        return None
    return float(t[0].item()), n_rec, float(t[1].item()) == 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--corpus-gib", type=float, default=64.0, help="device-resident corpus per GPU")
    ap.add_argument("--e2e-gib", type=float, default=4.0, help="host-resident sample for the end-to-end leg")
    ap.add_argument("--ref-files", type=int, default=4096, help="files of the corpus the CPU reference scans")
    ap.add_argument("--pattern", default=PATTERN)
    ap.add_argument("--quick", action="store_true", help="skip the cpu baseline and the e2e leg")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        return run_reference(a, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import grab_b200 as G

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's own log (version banner, NCCL_DEBUG output) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = G.Context(local_rank)

    free_b, total_b = torch.cuda.mem_get_info()
    corpus_bytes = int(a.corpus_gib * GiB)
    if corpus_bytes > free_b - 12 * GiB:
        corpus_bytes = max(GiB, (free_b - 12 * GiB))
    n_files = corpus_bytes // FILE_LEN
    corpus_bytes = n_files * FILE_LEN
    first_id = rank * n_files  # weak scaling: every rank owns its own shard of files (reference main.cc:94 strides by thread)
    dptr = ctx.device_alloc(corpus_bytes)
    ctx.synth_corpus(dptr, SEED, first_id, n_files, FILE_LEN, needle=PATTERN.encode(), needle_every=NEEDLE_EVERY)
    pat = G.Pattern(a.pattern, literal=True)
    batch = ctx.batch_create(G.Context.device_units(dptr, n_files, FILE_LEN, first_file_id=first_id))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    from grab_b200 import shard
    state = {"counts": None, "pending": None}

    def step():
        # the one collective of the path: an NCCL all-gather of the per-rank match counts, once per step.  It is
        # enqueued right after the scan and collected one step later (the last one before the timed region ends), so
        # its launch/completion latency and the rank skew it exposes overlap the next scan instead of adding to it
        r = ctx.batch_scan(pat, batch, G.MODE_ALL)
        h = shard.gather_counts_start(len(r))
        if state["pending"] is not None:
            state["counts"] = state["pending"].finish()
        state["pending"] = h
        return r

    def drain():
        if state["pending"] is not None:
            state["counts"] = state["pending"].finish()
            state["pending"] = None

    for _ in range(a.warmup):
        r = step()
    drain()
    # ---- parity gate (outside the timed region): planted needles + oracle on regenerated files ----
    import corpus
    import oracle_py as O
    ids = np.arange(first_id, first_id + n_files)
    planted = ids[ids % NEEDLE_EVERY == NEEDLE_EVERY // 2]
    want = {int(f): corpus.needle_offset(SEED, int(f), FILE_LEN, len(PATTERN)) for f in planted}
    got = {}
    for f, s in zip(r["file_id"].tolist(), r["start"].tolist()):
        got.setdefault(f, []).append(s)
    parity = all(got.get(f, [None])[0] == o or o in got.get(f, []) for f, o in want.items()) and a.pattern == PATTERN
    oreg = O.Regex(a.pattern, literal=True)
    sample_ids = sorted(set(list(planted[:4]) + [first_id, first_id + 1, first_id + n_files - 1]))
    for f in sample_ids:
        host = corpus.synth_file(SEED, int(f), FILE_LEN, PATTERN.encode(), NEEDLE_EVERY).tobytes()
        parity = parity and [s for s, _ in oreg.scan_window(host)] == got.get(int(f), [])
    extra = sum(len(v) for v in got.values()) - len(want)

    # ---- timed region: K resident steps ----
    sampler = ClockSampler(local_rank)
    kernel_ms, launches = [], 0
    sampler.start()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
        st = ctx.stats()
        kernel_ms.append(st["scan_kernel_ms"])
        launches += st["total_launches"]
    drain()  # every step's all-gather has completed inside the timed region
    sync_all()
    dt = time.perf_counter() - t0
    sampler.stop()
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    ms_per_step = dt / a.steps * 1e3
    value = world * corpus_bytes / (dt / a.steps) / 1e9

    # ---- roofline of the scan kernel ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k_ms = float(np.mean(kernel_ms))
    achieved = corpus_bytes / (k_ms * 1e-3) / 1e9
    probe_ms = min(ctx.read_probe(dptr, corpus_bytes)[0] for _ in range(3))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_corpus_byte")
        if traffic is not None:
            traffic = traffic * corpus_bytes
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, burst)" if peaks else "fallback 6650 GB/s",
                "kernel": "scan_kernel<FixedEngine<D,1>>", "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": corpus_bytes,
                "read_probe_gbs": corpus_bytes / (probe_ms * 1e-3) / 1e9}

    line = {"metric": "GB/s scanned", "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: literal (-S) over %.1f GiB/GPU synthetic corpus of 1 MiB files, all offsets (-O -l)" % (corpus_bytes / GiB),
                       "pattern": a.pattern, "files_per_gpu": int(n_files), "file_bytes": FILE_LEN, "mode": "ALL",
                       "sharding": "files by rank, no data-path collective; one all-gather of match counts per step",
                       "l2": "inputs larger than L2 (%.0f GiB vs 126 MB), no flush" % (corpus_bytes / GiB)},
            "roofline": roofline, "clocks": sampler.summary(), "gpu_launches": launches,
            "parity": "ok" if parity else "MISMATCH", "matches_per_step": int(state["counts"].sum()), "natural_hits": int(extra)}

    if rank == 0 and world == 1 and not a.quick:
        # ---- e2e: host buffers through gscan_scan_batch (pinned host memory, H2D + scan + D2H timed) ----
        e_files = int(min(a.e2e_gib * GiB, corpus_bytes) // FILE_LEN)
        hptr = G.lib().gscan_host_alloc(e_files * FILE_LEN)
        G.lib().gscan_memcpy_d2h(ctx._h, hptr, dptr, e_files * FILE_LEN)
        hunits = np.zeros(e_files, dtype=G.UNIT_DTYPE)
        hunits["ptr"] = hptr + np.arange(e_files, dtype=np.uint64) * np.uint64(FILE_LEN)
        hunits["len"] = FILE_LEN
        hunits["file_id"] = first_id + np.arange(e_files, dtype=np.uint32)
        for _ in range(2):
            re2e = ctx.scan_units(pat, hunits)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            re2e = ctx.scan_units(pat, hunits)
        torch.cuda.synchronize()
        e_dt = (time.perf_counter() - t0) / a.steps
        ref_first = r[r["file_id"] < first_id + e_files]
        if re2e.tobytes() != ref_first.tobytes():
            line["parity"] = "MISMATCH(e2e)"
        line["e2e"] = {"value": e_files * FILE_LEN / e_dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(e_files * FILE_LEN + e_files * 32),
                       "d2h_bytes_per_step": int(len(re2e) * 16 + 24), "sample": "%d files x 1 MiB in pinned host memory" % e_files}
        # the same call on PAGEABLE host memory (what the CLI's mmap windows are): staged by the engine's helper lanes
        pag = np.empty(e_files * FILE_LEN, dtype=np.uint8)
        ctypes_memmove(pag.ctypes.data, hptr, e_files * FILE_LEN)
        punits = hunits.copy()
        punits["ptr"] = pag.ctypes.data + np.arange(e_files, dtype=np.uint64) * np.uint64(FILE_LEN)
        ctx.scan_units(pat, punits)
        t0 = time.perf_counter()
        for _ in range(5):
            rp = ctx.scan_units(pat, punits)
        p_dt = (time.perf_counter() - t0) / 5
        if rp.tobytes() != re2e.tobytes():
            line["parity"] = "MISMATCH(e2e pageable)"
        line["e2e"]["pageable_host_value"] = e_files * FILE_LEN / p_dt / 1e9
        del pag
        G.lib().gscan_host_free(hptr)

        # ---- cpu baseline: the unmodified reference on this box's host cores ----
        if os.path.exists(REF_BIN):
            nf = min(a.ref_files, int(n_files))
            d = materialise_sample(nf, from_device=(ctx, dptr))
            try:
                cores = reference_cores()
                best = time_reference(d, nf, cores, 3)
                best1 = time_reference(d, min(nf, 1024), 1, 1) if False else None
                line["cpu_baseline"] = {"value": nf * FILE_LEN / best / 1e9, "unit": "GB/s", "cores": cores, "kind": "reference",
                                        "sample": "grab master + PCRE2-JIT shim, -n %d -r -O -l over the first %d files (tmpfs, warm), best of 3" % (cores, nf)}
                del best1
            finally:
                shutil.rmtree(d, ignore_errors=True)
        else:
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/grab_ref missing"}

    if world > 1 and not a.quick:
        # ---- e2e at N GPUs: every rank feeds its own GPU from its own pinned host sample over its own PCIe link,
        # same call and same timed region as at N = 1; aggregate = N x bytes / slowest rank
        e_files = int(min(a.e2e_gib * GiB, corpus_bytes) // FILE_LEN)
        hptr, scan = None, None
        try:
            hptr = G.lib().gscan_host_alloc(e_files * FILE_LEN)
            if not hptr:
                raise RuntimeError("pinned allocation failed")
            G.lib().gscan_memcpy_d2h(ctx._h, hptr, dptr, e_files * FILE_LEN)
            hunits = np.zeros(e_files, dtype=G.UNIT_DTYPE)
            hunits["ptr"] = hptr + np.arange(e_files, dtype=np.uint64) * np.uint64(FILE_LEN)
            hunits["len"] = FILE_LEN
            hunits["file_id"] = first_id + np.arange(e_files, dtype=np.uint32)
            for _ in range(2):
                ctx.scan_units(pat, hunits)
            scan = lambda: ctx.scan_units(pat, hunits)  # noqa: E731
        except Exception as ex:  # noqa: BLE001
            sys.stderr.write("bench: e2e leg at %d GPUs failed on rank %d: %r\n" % (world, rank, ex))
        want = r[r["file_id"] < first_id + e_files].tobytes()
        res = e2e_all_ranks(scan, want, a.steps, world, rank, torch, dist, "cuda", sync_all)
        if res is not None:
            e_dt, n_rec, ok = res
            if not ok:
                line["parity"] = "MISMATCH(e2e)"
            line["e2e"] = {"value": world * e_files * FILE_LEN / e_dt / 1e9, "unit": "GB/s",
                           "h2d_bytes_per_step": int(world * (e_files * FILE_LEN + e_files * 32)),
                           "d2h_bytes_per_step": int(world * (n_rec * 16 + 24)),
                           "sample": "%d files x 1 MiB in pinned host memory per rank, %d ranks, slowest rank" % (e_files, world)}
        if hptr:
            G.lib().gscan_host_free(hptr)

    batch.free()
    ctx.device_free(dptr)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
