"""Multi-GPU sharding of the detection path: one process per GPU, no collective on the data path.

Every (recording, channel) is an independent stream (upstream builds one SyllableDetector per channel/track:
Processor.swift:57-59, main.swift:86-89) and evaluation j is a pure function of samples
[j*hop, j*hop + gap + W + (T-1)*hop), so work shards two ways:

  level 1  recordings -> ranks, contiguous blocks balanced by sample count (`partition`)
  level 2  one long recording -> time slices cut on the hop lattice with a halo of gap + W + (T-2)*hop samples
           (`time_slices`), so evaluation indices and sample numbers are identical to the sequential run.

Only the sparse detection events are gathered (`gather_events`, torch.distributed gather of small arrays; NCCL or gloo).
Debounce is the one order-dependent step (TrackDetector.swift:80,99): after a level-2 split it must run on the gathered,
sorted events (`debounce_rows`), never inside a shard.
"""
import numpy as np


def partition(sizes, world):
    """Contiguous blocks of items (recordings) per rank, balanced by total size. -> [(start, end)] * world"""
    sizes = np.asarray(sizes, dtype=np.float64)
    n = sizes.size
    cum = np.concatenate([[0.0], np.cumsum(sizes)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(cum, target, side="left"))
        # pick the boundary closest to the target, keep it monotone
        if i > 0 and abs(cum[i - 1] - target) <= abs(cum[min(i, n)] - target):
            i -= 1
        bounds.append(min(max(i, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def time_slices(n_samples, hop, gap, win_len, time_range, world):
    """Split evaluations [0, E) of one recording into `world` contiguous ranges.
    -> [(eval_start, eval_end, sample_start, sample_end)]; the slice covers every sample its evaluations need."""
    need = gap + win_len
    cols = 0 if n_samples < need else (n_samples - need) // hop + 1
    E = max(0, cols - time_range + 1)
    out = []
    for r in range(world):
        e0, e1 = E * r // world, E * (r + 1) // world
        if e1 > e0:
            s0 = e0 * hop
            s1 = (e1 - 1) * hop + gap + win_len + (time_range - 1) * hop
            out.append((e0, e1, s0, min(s1, n_samples)))
        else:
            out.append((e0, e0, 0, 0))
    return out


def debounce_rows(samples, debounce_frames):
    """Greedy debounce over ascending sample numbers of ONE channel. -> boolean keep mask"""
    keep = np.zeros(len(samples), dtype=bool)
    until = -1
    for i, s in enumerate(samples):
        if until < s:
            keep[i] = True
            until = int(s) + int(debounce_frames)
    return keep


def _as_rows(outputs, n, dtype):
    """outputs as [n, O]; an empty [0, O] array keeps its O (reshape(0, -1) cannot infer it)"""
    outputs = np.asarray(outputs, dtype=dtype)
    if n == 0:
        return outputs.reshape(0, outputs.shape[-1] if outputs.ndim > 1 else 1)
    return outputs.reshape(n, -1)


def pack_events(recording, channel, sample, outputs):
    """-> float64 [n, 3 + O] rows (recording, channel, sample, out...) ready for gather; exact for |sample| < 2^53."""
    sample = np.asarray(sample, dtype=np.int64)
    outputs = _as_rows(outputs, sample.size, np.float64)
    rows = np.empty((sample.size, 3 + outputs.shape[1]), dtype=np.float64)
    rows[:, 0] = recording
    rows[:, 1] = channel
    rows[:, 2] = sample
    rows[:, 3:] = outputs
    return rows


def event_dtype(n_outputs):
    """Compact event row for the gather: key = recording << 16 | channel, sample number, network outputs: 12 + 4 O bytes where the
    float64 table spends 24 + 8 O. A corpus run gathers tens of millions of rows, so the row size IS the gather time."""
    return np.dtype([("key", "<u4"), ("sample", "<i8"), ("out", "<f4", (int(n_outputs),))])


def pack_events_compact(recording, channel, sample, outputs):
    """-> structured array (event_dtype) of the rows of one recording; recording and channels below 65 536."""
    sample = np.asarray(sample, dtype=np.int64)
    outputs = _as_rows(outputs, sample.size, np.float32)
    channel = np.asarray(channel)
    if not (0 <= int(recording) < 65536) or (channel.size and (int(channel.min()) < 0 or int(channel.max()) >= 65536)):
        raise ValueError("compact event rows hold recordings and channels below 65 536")
    rows = np.empty(sample.size, dtype=event_dtype(outputs.shape[1]))
    rows["key"] = (np.uint32(int(recording)) << np.uint32(16)) | channel.astype(np.uint32)
    rows["sample"] = sample
    rows["out"] = outputs
    return rows


def unpack_events(rows):
    """Structured rows (pack_events_compact / gather_events) -> recording, channel, sample, outputs arrays."""
    key = rows["key"]
    return (key >> np.uint32(16)).astype(np.int32), (key & np.uint32(0xFFFF)).astype(np.int32), rows["sample"], rows["out"]


def rows_in_order(rows, block=1 << 18):
    """True when rows [n, 3 + O] are ordered by (recording, channel, sample). One pass in cache-sized blocks: (recording, channel)
    folds into one exact float64 key (channel < 2^16), so two differences per block decide it and nothing the size of the table
    (tens of millions of rows after a corpus run) is ever allocated."""
    n = rows.shape[0]
    if rows.dtype.names:                     # compact rows: the key already is (recording, channel)
        key, smp = rows["key"], rows["sample"]
        for a in range(0, n - 1, block):
            b = min(n, a + block + 1)
            k0, k1 = key[a:b - 1], key[a + 1:b]
            if not bool(np.all((k1 > k0) | ((k1 == k0) & (smp[a + 1:b] >= smp[a:b - 1])))):
                return False
        return True
    for a in range(0, n - 1, block):
        b = min(n, a + block + 1)            # one row of overlap: the pair across the block boundary is checked too
        rc = rows[a:b, 0] * 65536.0 + rows[a:b, 1]
        d_rc = rc[1:] - rc[:-1]
        d_t = rows[a + 1:b, 2] - rows[a:b - 1, 2]
        if not bool(np.all((d_rc > 0) | ((d_rc == 0) & (d_t >= 0)))):
            return False
    return True


def _pinned_bytes(nbytes):
    """uint8 host buffer of nbytes: page-locked when a CUDA device is there (copies to and from the device then run at PCIe speed and
    nothing is staged), plain otherwise. Returns (numpy view, torch tensor or None)."""
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(int(nbytes), dtype=torch.uint8, pin_memory=True)
            return t.numpy(), t
    except Exception:  # noqa: BLE001 - no torch / no device: a plain buffer does
        pass
    return np.empty(int(nbytes), dtype=np.uint8), None


class EventTable:
    """A rank's detections as compact rows (event_dtype) in ONE reusable page-locked buffer: recordings are appended in place, so a
    job never concatenates per-recording tables or touches fresh pages per gather, and whether the rows are in (recording, channel,
    sample) order is tracked as they arrive (each block is checked while it is still in cache). gather_events takes the table as is."""

    def __init__(self, n_outputs, capacity=1 << 20):
        self.dtype = event_dtype(n_outputs)
        self._cap = 0
        self._raw = self._pin = None
        self.n = 0
        self.in_order = True
        self._reserve(int(capacity))

    def _reserve(self, cap):
        if cap <= self._cap:
            return
        cap = max(cap, 2 * self._cap)
        raw, pin = _pinned_bytes(cap * self.dtype.itemsize)
        if self.n:
            raw[:self.n * self.dtype.itemsize] = self._raw[:self.n * self.dtype.itemsize]
        self._raw, self._pin, self._cap = raw, pin, cap

    def clear(self):
        self.n = 0
        self.in_order = True

    @property
    def rows(self):
        return self._raw[:self.n * self.dtype.itemsize].view(self.dtype)

    def bytes_tensor(self):
        """the rows as a (pinned) torch uint8 tensor [n, itemsize], or None without torch"""
        if self._pin is None:
            return None
        return self._pin[:self.n * self.dtype.itemsize].view(self.n, self.dtype.itemsize)

    def claim(self, m):
        """room for m more rows -> (address of the first, view of the m rows); commit(m, ordered) makes them part of the table. For
        producers that write rows themselves (BatchDetector.run_into: the library fills them straight from its event list)."""
        self._reserve(self.n + int(m))
        it = self.dtype.itemsize
        blk = self._raw[self.n * it:(self.n + int(m)) * it]
        return blk.ctypes.data, blk.view(self.dtype)

    def commit(self, m, ordered):
        m = int(m)
        if m:
            it = self.dtype.itemsize
            if self.in_order:
                ok = bool(ordered)
                if ok and self.n:   # the pair across the boundary
                    ok = rows_in_order(self._raw[(self.n - 1) * it:(self.n + 1) * it].view(self.dtype))
                self.in_order = ok
            self.n += m

    def append(self, recording, channel, sample, outputs):
        """one recording's detections (as Events gives them: ordered by channel, sample)"""
        sample = np.asarray(sample)
        m = int(sample.size)
        channel = np.asarray(channel)
        if not (0 <= int(recording) < 65536) or (m and (int(channel.min()) < 0 or int(channel.max()) >= 65536)):
            raise ValueError("compact event rows hold recordings and channels below 65 536")
        self._reserve(self.n + m)
        it = self.dtype.itemsize
        blk = self._raw[self.n * it:(self.n + m) * it].view(self.dtype)
        blk["key"] = (np.uint32(int(recording)) << np.uint32(16)) | channel.astype(np.uint32)
        blk["sample"] = sample
        blk["out"] = _as_rows(outputs, m, np.float32)
        if self.in_order and m:
            lo = max(self.n - 1, 0)   # with the last row of what was there: the boundary pair is checked too
            self.in_order = rows_in_order(self._raw[lo * it:(self.n + m) * it].view(self.dtype))
        self.n += m
        return blk


_recv = {}


def _recv_table(dtype, total):
    """the destination rank's table of all ranks' rows: one page-locked buffer per row layout, kept between gathers (valid until the
    next gather_events of that layout)"""
    it = dtype.itemsize
    ent = _recv.get(dtype.str + str(dtype.itemsize))
    if ent is None or ent[0].size < total * it:
        ent = _pinned_bytes(max(total * it, 1))
        _recv[dtype.str + str(dtype.itemsize)] = ent
    return ent[0][:total * it].view(dtype), (ent[1][:total * it].view(total, it) if ent[1] is not None and total else None)


def reserve_gather(n_outputs, total_rows):
    """Page-lock the destination table of a coming gather of about `total_rows` compact rows ahead of time (a job that knows its size:
    pinning hundreds of megabytes costs more than moving them). Call on the destination rank; other ranks need nothing."""
    _recv_table(event_dtype(n_outputs), int(total_rows))


def gather_events(rows, dist=None, dst=0):
    """Gather per-rank event rows on `dst`, sorted by (recording, channel, sample). Other ranks get None.
    `dist` is torch.distributed (initialised) or None for a single process. `rows` is the float64 table of pack_events, the structured
    array of pack_events_compact (less than half the bytes per row), or an EventTable of such rows (what bench.py gathers: page-locked,
    filled in place, order tracked on the way); the result has the form of the rows. For compact rows the result on `dst` lives in a
    buffer that the next gather of that layout reuses.

    Only `dst` receives rows (point-to-point gather, not all_gather). Every rank checks the order of its own rows (in parallel);
    ranks own contiguous recording blocks, so when every block is ordered and the block boundaries are, the concatenation in rank
    order IS the sorted result and no sort of the (tens of millions of) rows is needed; otherwise `dst` sorts."""
    table = rows if isinstance(rows, EventTable) else None
    if table is not None:
        rows = table.rows
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        allrows = rows
        in_order = table.in_order if table is not None else rows_in_order(rows)
    else:
        import torch
        world, rank = dist.get_world_size(), dist.get_rank()
        nccl = dist.get_backend() == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
        if rows.dtype.names:
            return _gather_compact(rows, dist, dst, torch, world, rank, dev, table)
        n, width = rows.shape[0], (rows.shape[1] if rows.size else 0)
        first = rows[0, :3] if n else np.zeros(3)
        last = rows[-1, :3] if n else np.zeros(3)
        meta = torch.tensor([float(n), float(width), 1.0 if rows_in_order(rows) else 0.0, *first, *last], dtype=torch.float64, device=dev)
        metas = [torch.zeros_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta)
        metas = [m.cpu().numpy() for m in metas]
        counts = [int(m[0]) for m in metas]
        width = int(max(m[1] for m in metas))
        cap = max(max(counts), 1)
        buf = torch.zeros((cap, max(width, 1)), dtype=torch.float64, device=dev)
        if rows.size:
            src = torch.from_numpy(np.ascontiguousarray(rows))
            buf[:n].copy_(src.pin_memory() if nccl else src, non_blocking=nccl)
        bufs = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, bufs, dst=dst)
        if rank != dst:
            return None
        total = sum(counts)
        allrows = np.zeros((0, max(width, 3)), dtype=np.float64)
        if total:
            host = torch.empty((total, max(width, 3)), dtype=torch.float64, pin_memory=nccl)
            pos = 0
            for b, c in zip(bufs, counts):
                if c:
                    host[pos:pos + c].copy_(b[:c], non_blocking=nccl)
                    pos += c
            if nccl:
                torch.cuda.synchronize(dev)
            allrows = host.numpy()
        # ordered blocks in rank order, boundaries in order => sorted
        in_order = all(m[2] > 0.5 for m in metas)
        prev = None
        for m in metas:
            if int(m[0]) == 0:
                continue
            if prev is not None and tuple(m[3:6]) < prev:
                in_order = False
            prev = tuple(m[6:9])
    if allrows.shape[0] > 1 and not in_order:
        if allrows.dtype.names:
            allrows = allrows[np.lexsort((allrows["sample"], allrows["key"]))]
        else:
            allrows = allrows[np.lexsort((allrows[:, 2], allrows[:, 1], allrows[:, 0]))]
    return allrows


def _gather_compact(rows, dist, dst, torch, world, rank, dev, table=None):
    """gather_events for structured rows: the rows travel as raw bytes (one point-to-point gather of uint8 tensors). From an EventTable
    they leave page-locked memory and the order is already known; on `dst` they land in a page-locked table kept between gathers."""
    n, item = rows.shape[0], rows.dtype.itemsize
    first = (int(rows["key"][0]), int(rows["sample"][0])) if n else (0, 0)
    last = (int(rows["key"][-1]), int(rows["sample"][-1])) if n else (0, 0)
    ordered = table.in_order if table is not None else rows_in_order(rows)
    meta = torch.tensor([n, item, 1 if ordered else 0, *first, *last], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [m.cpu().numpy() for m in metas]
    counts = [int(m[0]) for m in metas]
    if any(int(m[1]) != item for m in metas if int(m[0])):
        raise ValueError("ranks disagree on the event row layout")
    cap = max(max(counts), 1)
    buf = torch.empty((cap, item), dtype=torch.uint8, device=dev)
    if n:
        src = table.bytes_tensor() if table is not None else None
        if src is None:
            src = torch.from_numpy(np.ascontiguousarray(rows).view(np.uint8).reshape(n, item))
        buf[:n].copy_(src, non_blocking=src.is_pinned() if dev.type == "cuda" else False)
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst)
    if rank != dst:
        return None
    total = sum(counts)
    allrows, raw = _recv_table(rows.dtype, total)
    if raw is None and total:
        raw = torch.from_numpy(allrows.view(np.uint8).reshape(total, item))
    pos = 0
    for b, c in zip(bufs, counts):
        if c:
            raw[pos:pos + c].copy_(b[:c], non_blocking=dev.type == "cuda")   # device -> the final table, no intermediate host copy
            pos += c
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    in_order = all(int(m[2]) == 1 for m in metas)
    prev = None
    for m in metas:     # ordered blocks in rank order, boundaries in order => sorted
        if int(m[0]) == 0:
            continue
        if prev is not None and (int(m[3]), int(m[4])) < prev:
            in_order = False
        prev = (int(m[5]), int(m[6]))
    if total > 1 and not in_order:
        allrows = allrows[np.lexsort((allrows["sample"], allrows["key"]))]
    return allrows
