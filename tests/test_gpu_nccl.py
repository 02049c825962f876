"""The N > 1 exchange on real GPUs: two ranks, NCCL, the asynchronous record gather of BASELINE config 5
(shard.gather_matches_start reading the records from device memory).  Skipped below 2 GPUs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import corpus
    import grab_b200 as G
    from grab_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = G.Context(rank)
    n, flen, seed = 4096, 16384, 5
    d = ctx.device_alloc(n * flen)
    ctx.synth_corpus(d, seed, rank * n, n, flen)
    batch = ctx.batch_create(G.Context.device_units(d, n, flen, first_file_id=rank * n))
    p = G.Pattern(corpus.literals100())
    out = []
    pend = None
    for _ in range(3):  # pipelined like bench.py: start after the scan, collect one step later
        r = ctx.batch_scan(p, batch, copy=False)
        h = shard.gather_matches_start(r, dst=0, device_records=ctx.last_device_matches())
        if pend is not None:
            out.append(pend.finish())
        pend = h
    out.append(pend.finish())
    host_path = shard.gather_matches(r.copy(), dst=0)  # the blocking flavour through host memory: same answer
    counts = shard.gather_counts(len(r))
    if rank == 0:
        q.put((r.copy().tobytes(), [o.tobytes() for o in out], host_path.tobytes(), counts.tolist()))
    else:
        assert all(o is None for o in out) and host_path is None
        q.put((r.copy().tobytes(), None, None, counts.tolist()))
    dist.barrier()
    batch.free()
    ctx.device_free(d)
    ctx.close()
    dist.destroy_process_group()


def test_two_rank_record_gather_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from grab_b200 import shard
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    root = [x for x in res if x[1] is not None][0]
    other = [x for x in res if x[1] is None][0]
    mine = np.frombuffer(root[0], dtype=shard.MATCH_DTYPE)
    theirs = np.frombuffer(other[0], dtype=shard.MATCH_DTYPE)
    want = np.concatenate([mine, theirs]).tobytes()  # rank 0 owns the lower file ids
    assert len(mine) > 100 and len(theirs) > 100
    assert all(o == want for o in root[1]) and root[2] == want
    assert root[3] == [len(mine), len(theirs)] == other[3]
