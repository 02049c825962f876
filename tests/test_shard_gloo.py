"""N>1 path on CPU: world_size-2 gloo.  The scan results per rank come from the oracle (the product
has no CPU scan); what is under test is the sharding rule and the count / record exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import corpus
import oracle_py as O
from grab_b200 import shard

N_FILES, FILE_LEN, PAT = 12, 1 << 14, "foo|bar|baz|quux"


def _scan_files(ids):
    re = O.Regex(PAT)
    recs = []
    for f in ids:
        data = corpus.synth_file(2, int(f), FILE_LEN).tobytes()
        for s, l in re.scan_window(data):
            recs.append((s, int(f), l))
    return np.array(recs, dtype=shard.MATCH_DTYPE) if recs else np.zeros(0, dtype=shard.MATCH_DTYPE)


def _worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = shard.files_of_rank(N_FILES, rank, world, mode)
    local = _scan_files(ids)
    counts = shard.gather_counts(len(local))
    pend = shard.gather_counts_start(len(local))  # asynchronous flavour: same answer
    assert pend.finish().tolist() == counts.tolist()
    merged = shard.gather_matches(local, dst=0)
    q.put((rank, ids.tolist(), counts.tolist(), None if merged is None else merged.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["stride", "block"])
def test_two_rank_shard_and_gather(mode):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids0, ids1 = res[0][1], res[1][1]
    assert sorted(ids0 + ids1) == list(range(N_FILES)) and not set(ids0) & set(ids1)
    if mode == "stride":
        assert ids0 == list(range(0, N_FILES, 2))  # reference main.cc:94
    single = np.sort(_scan_files(range(N_FILES)), order=["file_id", "start"], kind="stable")
    assert res[0][2] == res[1][2] and sum(res[0][2]) == len(single)
    assert res[0][3] == single.tobytes() and res[1][3] is None


def test_single_rank_passthrough():
    local = _scan_files(range(3))
    assert shard.gather_counts(len(local)).tolist() == [len(local)]
    assert shard.gather_counts_start(len(local)).finish().tolist() == [len(local)]
    assert shard.gather_matches(local).tobytes() == np.sort(local, order=["file_id", "start"], kind="stable").tobytes()


# ---- f4: one file, its reference windows spread over the ranks ---------------------------------------------
BIG_CHUNK, BIG_SIZE = 1 << 16, 5 * (1 << 16) + 777


def _big_image():
    a = corpus.synth_file(7, 1, BIG_SIZE).copy()
    C = BIG_CHUNK
    for off in (C - 4000, C - 100, C - 2, 2 * (C - 4096) + 5, 3 * (C - 4096) - 1, BIG_SIZE - 5):  # in overlaps, straddling, at EOF
        a[off:off + 3] = np.frombuffer(b"foo", dtype=np.uint8)
    return a.tobytes()


def _scan_windows(img, idx):
    re = O.Regex(PAT)
    wins = shard.file_windows(len(img), BIG_CHUNK)
    recs = []
    for w in idx:
        off, clen = wins[int(w)]
        for s, l in re.scan_window(img[off:off + clen], base_off=off):
            recs.append((s, int(w), l))
    return np.array(recs, dtype=shard.MATCH_DTYPE) if recs else np.zeros(0, dtype=shard.MATCH_DTYPE)


def _worker_f4(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    img = _big_image()
    idx = shard.windows_of_rank(len(shard.file_windows(len(img), BIG_CHUNK)), rank, world)
    merged = shard.gather_matches(_scan_windows(img, idx), dst=0)
    q.put((rank, None if merged is None else merged["start"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_file_windows_match_reference_loop():
    import chunker
    for size in (0, 1, 4095, 4096, 4097, BIG_CHUNK - 1, BIG_CHUNK, BIG_CHUNK + 1, 2 * BIG_CHUNK - 4096, BIG_SIZE, 10 * BIG_CHUNK + 3):
        assert shard.file_windows(size, BIG_CHUNK) == chunker.windows(size, BIG_CHUNK)
    w = shard.file_windows(BIG_SIZE, BIG_CHUNK)
    assert all(w[i + 1][0] == w[i][0] + BIG_CHUNK - 4096 for i in range(len(w) - 1)) and w[-1][0] + w[-1][1] == BIG_SIZE
    idx = [shard.windows_of_rank(len(w), r, 3).tolist() for r in range(3)]
    assert sorted(sum(idx, [])) == list(range(len(w)))


def test_two_rank_one_file():
    """The merged records of 2 ranks == the reference's sequence for the whole file (oracle restatement of the chunk
    loop), duplicates of the overlaps in the reference's order."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_f4, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    img = _big_image()
    want = [s for s, l in O.Regex(PAT).scan_file(img, chunk_size=BIG_CHUNK)]
    assert res[1][1] is None and res[0][1] == want
    assert len(want) > len(set(want))  # the overlaps did produce duplicates
