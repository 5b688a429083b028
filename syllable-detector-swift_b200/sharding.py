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


def gather_events(rows, dist=None, dst=0):
    """Gather per-rank event rows on `dst`, sorted by (recording, channel, sample). Other ranks get None.
    `dist` is torch.distributed (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        allrows = rows
    else:
        import torch
        world, rank = dist.get_world_size(), dist.get_rank()
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        width = torch.tensor([rows.shape[1] if rows.size else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(width, op=dist.ReduceOp.MAX)
        width = int(width.item())
        count = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(counts, count)
        counts = [int(c.item()) for c in counts]
        cap = max(max(counts), 1)
        buf = torch.zeros((cap, max(width, 1)), dtype=torch.float64, device=dev)
        if rows.size:
            buf[:rows.shape[0]] = torch.from_numpy(np.ascontiguousarray(rows)).to(dev)
        bufs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf)
        if rank != dst:
            return None
        allrows = np.concatenate([b[:c].cpu().numpy() for b, c in zip(bufs, counts)], axis=0) if sum(counts) else np.zeros((0, max(width, 3)))
    if allrows.shape[0] > 1:
        # every rank's rows arrive ordered by (recording, channel, sample) and ranks own contiguous recording blocks, so the
        # concatenation is usually in order already: one vectorised check instead of a sort of tens of millions of rows
        r, c, t = allrows[:, 0], allrows[:, 1], allrows[:, 2]
        dr, dc, dt = np.diff(r), np.diff(c), np.diff(t)
        in_order = bool(np.all((dr > 0) | ((dr == 0) & ((dc > 0) | ((dc == 0) & (dt >= 0))))))
        if not in_order:
            allrows = allrows[np.lexsort((t, c, r))]
    return allrows
