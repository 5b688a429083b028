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


def pack_events(recording, channel, sample, outputs):
    """-> float64 [n, 3 + O] rows (recording, channel, sample, out...) ready for gather; exact for |sample| < 2^53."""
    sample = np.asarray(sample, dtype=np.int64)
    outputs = np.asarray(outputs, dtype=np.float64).reshape(sample.size, -1)
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
    outputs = np.asarray(outputs, dtype=np.float32).reshape(sample.size, -1)
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
        for a in range(0, n - 1, block):
            b = min(n, a + block + 1)
            key = rows["key"][a:b].astype(np.int64)
            d_k = key[1:] - key[:-1]
            d_t = rows["sample"][a + 1:b] - rows["sample"][a:b - 1]
            if not bool(np.all((d_k > 0) | ((d_k == 0) & (d_t >= 0)))):
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


def gather_events(rows, dist=None, dst=0):
    """Gather per-rank event rows on `dst`, sorted by (recording, channel, sample). Other ranks get None.
    `dist` is torch.distributed (initialised) or None for a single process. `rows` is either the float64 table of pack_events or
    the structured array of pack_events_compact (less than half the bytes per row; what bench.py gathers); the result has the same form.

    Only `dst` receives rows (point-to-point gather, not all_gather). Every rank checks the order of its own rows (in parallel);
    ranks own contiguous recording blocks, so when every block is ordered and the block boundaries are, the concatenation in rank
    order IS the sorted result and no sort of the (tens of millions of) rows is needed; otherwise `dst` sorts."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        allrows = rows
        in_order = rows_in_order(rows)
    else:
        import torch
        world, rank = dist.get_world_size(), dist.get_rank()
        nccl = dist.get_backend() == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
        if rows.dtype.names:
            return _gather_compact(rows, dist, dst, torch, world, rank, dev)
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


def _gather_compact(rows, dist, dst, torch, world, rank, dev):
    """gather_events for structured rows: the rows travel as raw bytes (one point-to-point gather of uint8 tensors, pageable host
    memory on both sides: pinning a buffer of this size costs more than the copy saves)."""
    n, item = rows.shape[0], rows.dtype.itemsize
    first = (int(rows["key"][0]), int(rows["sample"][0])) if n else (0, 0)
    last = (int(rows["key"][-1]), int(rows["sample"][-1])) if n else (0, 0)
    meta = torch.tensor([n, item, 1 if rows_in_order(rows) else 0, *first, *last], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [m.cpu().numpy() for m in metas]
    counts = [int(m[0]) for m in metas]
    if any(int(m[1]) != item for m in metas if int(m[0])):
        raise ValueError("ranks disagree on the event row layout")
    cap = max(max(counts), 1)
    buf = torch.zeros((cap, item), dtype=torch.uint8, device=dev)
    if n:
        buf[:n].copy_(torch.from_numpy(np.ascontiguousarray(rows).view(np.uint8).reshape(n, item)))
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst)
    if rank != dst:
        return None
    total = sum(counts)
    allrows = np.empty(total, dtype=rows.dtype)
    raw = allrows.view(np.uint8).reshape(total, item) if total else None
    pos = 0
    for b, c in zip(bufs, counts):
        if c:
            torch.from_numpy(raw[pos:pos + c]).copy_(b[:c])   # device -> the final table, no intermediate host copy
            pos += c
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
