"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard recordings / time slices, gather events, and the
result equals the single-process run. (The detection step itself is stood in for by the oracle here; the GPU path is
covered by the -m gpu tests.)"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, SAMPLE_TXT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("syllable-detector-swift_b200.sharding")
    synth = importlib.import_module("tools.synth")
    from oracle import Oracle
    orc = Oracle(SAMPLE_TXT)
    rows = []
    if mode == "recordings":
        lengths = [44100 * 2, 44100 * 3, 44100, 44100 * 4, 44100 * 2]
        a, b = sh.partition(lengths, world)[rank]
        for rec in range(a, b):
            x = synth.make_audio(2, lengths[rec], seed=100 + rec)
            for ch in range(2):
                s, _, outs = orc.events(x[ch], 2205)
                rows.append(sh.pack_events(rec, ch, s, outs))
    else:  # one long recording split in time; debounce only after the gather
        x = synth.make_audio(1, 44100 * 8, seed=7)[0]
        e0, e1, s0, s1 = sh.time_slices(x.size, orc.stride, orc.gap, orc.windowLength, orc.timeRange, world)[rank]
        outs, da, _ = orc.run(x[s0:s1])
        assert outs.shape[0] == e1 - e0
        j = np.nonzero(da)[0]
        rows.append(sh.pack_events(0, 0, np.array([orc.eval_sample(int(k) + e0) for k in j], dtype=np.int64), outs[j]))
    local = np.concatenate(rows, axis=0) if rows else np.zeros((0, 4))
    if mode == "recordings":   # the compact transport must deliver the same rows
        crow = [sh.pack_events_compact(int(r[0, 0]), r[:, 1].astype(np.int64), r[:, 2].astype(np.int64), r[:, 3:]) for r in rows if r.shape[0]]
        compact = np.concatenate(crow) if crow else np.zeros(0, dtype=sh.event_dtype(1))
        call = sh.gather_events(compact, dist, dst=0)
        if rank == 0:
            rec, ch, smp, outs = sh.unpack_events(call)
            np.save(out_path + ".compact.npy", np.column_stack([rec, ch, smp, outs.astype(np.float64)]))
        else:
            assert call is None
        # the same rows appended recording by recording to an EventTable (in place, order tracked on the way), gathered twice: the
        # destination's table is reused between gathers
        table = sh.EventTable(1, capacity=4)
        for rep in range(2):
            table.clear()
            for r in rows:
                if r.shape[0]:
                    table.append(int(r[0, 0]), r[:, 1].astype(np.int64), r[:, 2].astype(np.int64), r[:, 3:])
            assert table.in_order and np.array_equal(table.rows, compact)
            tall = sh.gather_events(table, dist, dst=0)
            assert (tall is None) == (rank != 0)
            if rank == 0:
                assert np.array_equal(tall, call)
    allrows = sh.gather_events(local, dist, dst=0)
    if rank == 0:
        np.save(out_path, allrows)
    else:
        assert allrows is None
    dist.barrier()
    dist.destroy_process_group()


def _run(mode, tmp_path):
    out = str(tmp_path / ("rows_%s.npy" % mode))
    mp.spawn(_worker, args=(2, _free_port(), mode, out), nprocs=2, join=True)
    return np.load(out)


def test_partition_and_slices():
    sh = importlib.import_module("syllable-detector-swift_b200.sharding")
    assert sh.partition([1, 1, 1, 1], 2) == [(0, 2), (2, 4)]
    assert sh.partition([10, 1, 1, 1, 1], 2)[0] == (0, 1)
    parts = sh.partition([3, 1, 4, 1, 5, 9, 2, 6], 4)
    assert parts[0][0] == 0 and parts[-1][1] == 8 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert sh.partition([5, 5], 4)[-1][1] == 2  # more ranks than recordings: some ranks idle
    sl = sh.time_slices(44100 * 8, 132, 0, 256, 10, 3)
    assert sl[0][0] == 0 and sl[-1][1] == (44100 * 8 - 256) // 132 + 1 - 9
    for e0, e1, s0, s1 in sl:
        assert s0 == e0 * 132 and s1 - s0 == (e1 - e0 - 1) * 132 + 1444
    assert sh.time_slices(100, 132, 0, 256, 10, 2) == [(0, 0, 0, 0), (0, 0, 0, 0)]
    assert list(sh.debounce_rows([10, 20, 30, 300], 100)) == [True, False, False, True]


def test_two_rank_recording_shards_match_single_process(tmp_path, oracle_mod, synth):
    sh = importlib.import_module("syllable-detector-swift_b200.sharding")
    got = _run("recordings", tmp_path)
    orc = oracle_mod.Oracle(SAMPLE_TXT)
    rows = []
    for rec, n in enumerate([44100 * 2, 44100 * 3, 44100, 44100 * 4, 44100 * 2]):
        x = synth.make_audio(2, n, seed=100 + rec)
        for ch in range(2):
            s, _, outs = orc.events(x[ch], 2205)
            rows.append(sh.pack_events(rec, ch, s, outs))
    want = np.concatenate(rows, axis=0)
    assert got.shape == want.shape and got.shape[0] > 10 and np.array_equal(got, want)
    compact = np.load(str(tmp_path / "rows_recordings.npy") + ".compact.npy")
    assert np.array_equal(compact, want)           # structured rows through the byte gather: the same table
    c = sh.pack_events_compact(3, want[:5, 1].astype(np.int64), want[:5, 2].astype(np.int64), want[:5, 3:])
    assert c.dtype.itemsize == 16 and sh.rows_in_order(c) and not sh.rows_in_order(c[::-1])
    assert sh.gather_events(c, None).shape[0] == 5
    # a recording without detections packs to zero rows of the right layout (reshape(0, -1) cannot infer the width)
    empty = sh.pack_events_compact(7, np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros((0, 1), np.float32))
    assert empty.shape == (0,) and empty.dtype == c.dtype and sh.pack_events(7, [], [], np.zeros((0, 2))).shape == (0, 5)
    t = sh.EventTable(want.shape[1] - 3, capacity=2)
    t.append(1, np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros((0, 1), np.float32))
    assert t.n == 0 and t.in_order
    t.append(3, want[:5, 1].astype(np.int64), want[:5, 2].astype(np.int64), want[:5, 3:])
    assert t.in_order and t.n == 5 and np.array_equal(t.rows, c) and np.array_equal(sh.gather_events(t, None), c)
    t.append(2, want[:5, 1].astype(np.int64), want[:5, 2].astype(np.int64), want[:5, 3:])     # an earlier recording after a later one
    assert not t.in_order and sh.rows_in_order(sh.gather_events(t, None)) and sh.gather_events(t, None).shape[0] == 10


def test_two_rank_time_slices_match_sequential_run(tmp_path, oracle_mod, synth):
    sh = importlib.import_module("syllable-detector-swift_b200.sharding")
    got = _run("slices", tmp_path)
    orc = oracle_mod.Oracle(SAMPLE_TXT)
    x = synth.make_audio(1, 44100 * 8, seed=7)[0]
    s, _, outs = orc.events(x, 0)
    assert np.array_equal(got[:, 2].astype(np.int64), s) and np.array_equal(got[:, 3], outs[:, 0].astype(np.float64))
    keep = sh.debounce_rows(got[:, 2], 2205)  # debounce AFTER the gather equals the sequential debounce
    s_db, _, _ = orc.events(x, 2205)
    assert np.array_equal(got[keep, 2].astype(np.int64), s_db)
