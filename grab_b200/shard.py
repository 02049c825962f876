"""Multi-GPU plumbing for the scan path: one process per GPU, files sharded by rank, no collective on
the data path.

The reference's only parallel axis is a static round-robin of files over N threads with nothing
shared but stdout (/root/reference/src/main.cc:86-100, stride at :94); ranks take the place of
threads.  The single exchange is at the end: an all-gather of per-rank match counts and, when the
caller wants the records in one place (BASELINE config 5), a gather of (file_id, start, len)
records to one rank followed by a merge by file id.  Payload is bytes to megabytes, so the backend
only matters for latency: NCCL on GPUs, gloo in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist

MATCH_DTYPE = np.dtype([("start", "<u8"), ("file_id", "<u4"), ("match_len", "<u4")])


def files_of_rank(n_files, rank, world, mode="stride"):
    """File ids scanned by `rank`.  'stride' is the reference's rule (main.cc:94: i, i+N, i+2N, ...);
    'block' gives contiguous, size-balanced ranges (better locality for a device-resident corpus)."""
    if mode == "stride":
        return np.arange(rank, n_files, world, dtype=np.int64)
    per, extra = divmod(n_files, world)
    lo = rank * per + min(rank, extra)
    return np.arange(lo, lo + per + (1 if rank < extra else 0), dtype=np.int64)


def file_windows(size, chunk_size=1 << 30, overlap=0x1000):
    """(off, clen) of the windows the reference cuts one file into (grab.cc:151-159): clen = min(chunk, size - off),
    off += chunk - 4096.  They are independent scan units (every window is searched statelessly, quirk Q3), which
    makes them the partition for scanning ONE huge file on several GPUs."""
    out, off = [], 0
    while off < size:
        out.append((off, min(chunk_size, size - off)))
        off += chunk_size - overlap
    return out


def windows_of_rank(n_windows, rank, world):
    """Window indices of one file scanned by `rank`: round-robin, like the files (main.cc:94).  Scan them as units
    with file_id = window index and base_off = off; gather_matches() then restores the reference's output order
    (window by window, ascending inside a window -- duplicates in the overlaps included)."""
    return np.arange(rank, n_windows, world, dtype=np.int64)


def _device(group=None):
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def gather_counts(n_local, group=None):
    """All-gather of one int64 per rank: every rank learns every rank's match count."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.array([n_local], dtype=np.int64)
    dev = _device(group)
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=dev)
    out = torch.zeros(dist.get_world_size(group), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.cpu().numpy()


class PendingCounts:
    """An all-gather of match counts in flight (gather_counts_start): finish() returns the per-rank counts.  Lets the
    next scan overlap the collective's launch and completion latency -- the counts are only needed at the end."""

    def __init__(self, out, work, n_local):
        self._out, self._work, self._n = out, work, n_local

    def finish(self):
        if self._work is None:
            return np.array([self._n], dtype=np.int64) if self._out is None else self._out.cpu().numpy()
        self._work.wait()
        return self._out.cpu().numpy()


def gather_counts_start(n_local, group=None):
    """Asynchronous gather_counts(): enqueues the all-gather and returns a PendingCounts."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return PendingCounts(None, None, int(n_local))
    dev = _device(group)
    mine = torch.tensor([int(n_local)], dtype=torch.int64).to(dev, non_blocking=True)
    out = torch.zeros(dist.get_world_size(group), dtype=torch.int64, device=dev)
    work = dist.all_gather_into_tensor(out, mine, group=group, async_op=True)
    return PendingCounts(out, work, int(n_local))


def merge_parts(parts):
    """Records of several ranks -> one array ordered by (file_id, start).  Every part is already ordered (the engine
    returns records by unit, then offset); when the ranks own increasing, disjoint file-id ranges (block sharding) the
    concatenation is the answer, otherwise (stride sharding) a lexsort by (file_id, start)."""
    parts = [p for p in parts if len(p)]
    if not parts:
        return np.zeros(0, dtype=MATCH_DTYPE)
    allm = parts[0] if len(parts) == 1 else np.concatenate(parts)
    if all(int(parts[i]["file_id"][-1]) < int(parts[i + 1]["file_id"][0]) for i in range(len(parts) - 1)):
        return allm
    return allm[np.lexsort((allm["start"], allm["file_id"]))]


class _DevRecords:
    """__cuda_array_interface__ over the engine's device records, so torch can wrap them without a copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "strides": None, "version": 2}


class PendingMatches:
    """A gather of match records in flight (gather_matches_start): finish() returns the merged records on `dst` (a view of a
    pinned buffer that stays valid until the next finish()), None elsewhere."""

    _pinned = None  # grow-only pinned landing buffer on the destination rank

    def __init__(self, result=None, event=None, counts=None, host=None, is_dst=False, keep=None):
        self._result, self._event, self._counts, self._host, self._is_dst, self._keep = result, event, counts, host, is_dst, keep

    def finish(self):
        if self._event is None:
            return self._result
        self._event.synchronize()
        self._keep = None
        if not self._is_dst:
            return None
        offs = np.concatenate([[0], np.cumsum(self._counts)])
        rec = self._host.numpy()[: int(offs[-1]) * MATCH_DTYPE.itemsize].view(MATCH_DTYPE)
        parts = [rec[int(offs[r]):int(offs[r + 1])] for r in range(len(self._counts))]
        live = [p for p in parts if len(p)]
        # the parts already sit back to back in the landing buffer: with increasing, disjoint file-id ranges per rank (block
        # sharding) that IS the merged result -- no copy (concatenating 50 MB of records cost more than the scan of a step)
        if all(int(live[i]["file_id"][-1]) < int(live[i + 1]["file_id"][0]) for i in range(len(live) - 1)):
            return rec
        return merge_parts(parts)


def gather_matches_start(local, dst=0, group=None, device_records=None):
    """Asynchronous gather_matches(): counts first (small blocking all-gather: they size the exchange), then ONE padded
    all-gather of the raw records on a side stream, followed on the destination rank by device-to-host copies of every
    rank's live part straight into a contiguous pinned buffer -- all enqueued, nothing waited for: the caller scans the
    next batch meanwhile and collects with finish().  device_records = (device pointer, count) of the records in HBM
    (Context.last_device_matches()) avoids bouncing this rank's records through the host."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1 or dist.get_backend(group) != "nccl":
        return PendingMatches(result=gather_matches(local, dst, group))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local = len(local) if device_records is None else int(device_records[1])
    counts = gather_counts(n_local, group)
    cap = max(int(counts.max()), 1) * MATCH_DTYPE.itemsize
    dev = _device(group)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        buf = torch.empty(cap, dtype=torch.uint8, device=dev)
        nb = n_local * MATCH_DTYPE.itemsize
        if nb:
            src = None
            if device_records is not None:
                try:
                    src = torch.as_tensor(_DevRecords(int(device_records[0]), nb), device=dev)
                except Exception:  # noqa: BLE001 -- a torch build without __cuda_array_interface__ import: go through the host
                    src = None
            if src is None:
                local = np.ascontiguousarray(local, dtype=MATCH_DTYPE)
                src = torch.from_numpy(local.view(np.uint8).reshape(-1).copy())
            buf[:nb].copy_(src)
        out = torch.empty(world * cap, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, buf, group=group)  # enqueued on `side` (NCCL orders it after the copies above)
        host = None
        if rank == dst:
            total = int(counts.sum()) * MATCH_DTYPE.itemsize
            if PendingMatches._pinned is None or PendingMatches._pinned.numel() < total:
                PendingMatches._pinned = torch.empty(max(total * 5 // 4, 1 << 20), dtype=torch.uint8, pin_memory=True)
            host = PendingMatches._pinned
            o = 0
            for r in range(world):
                nbr = int(counts[r]) * MATCH_DTYPE.itemsize
                if nbr:
                    host[o:o + nbr].copy_(out[r * cap:r * cap + nbr], non_blocking=True)
                o += nbr
        ev = torch.cuda.Event()
        ev.record(side)
    return PendingMatches(event=ev, counts=counts, host=host, is_dst=rank == dst, keep=(buf, out, side))


def gather_matches(local, dst=0, group=None):
    """Gathers MATCH_DTYPE records of all ranks on `dst`, merged by (file_id, start); others get None.
    Counts first (all-gather), then one padded all-gather of the raw record bytes."""
    local = np.ascontiguousarray(local, dtype=MATCH_DTYPE)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return merge_parts([local])
    counts = gather_counts(len(local), group)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cap = int(counts.max())
    dev = _device(group)
    buf = torch.zeros(max(cap, 1) * MATCH_DTYPE.itemsize, dtype=torch.uint8)
    if len(local):
        buf[: local.nbytes] = torch.from_numpy(local.view(np.uint8).reshape(-1).copy())
    buf = buf.to(dev)
    out = torch.zeros(world * buf.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    if rank != dst:
        return None
    raw = out.cpu().numpy().reshape(world, -1)
    return merge_parts([raw[r, : int(counts[r]) * MATCH_DTYPE.itemsize].view(MATCH_DTYPE) for r in range(world)])
