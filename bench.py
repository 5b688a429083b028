#!/usr/bin/env python
"""Benchmark of the detection hot path (STFT -> band -> sliding window -> MLP -> detect) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4] [--hours H] [--channels C]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--config 2 (default, BASELINE.json configs[1], the configuration the metric is quoted on): the sample.txt network over a 1-hour,
  8-channel, 44.1 kHz synthetic recording per GPU (noise + synthetic syllables, generated on the device). One step = one pass of
  the whole recording through the hot path. With N > 1 every rank owns its own recording (sharding by recording, no collective on
  the data path): weak scaling.
--config 3 (configs[2]): a corpus of --corpus-hours audio-hours (default 1000) as 1 h x 8 ch recordings sharded by recording over the
  ranks (strong scaling): per recording device synthesis (untimed), detection + event collection (timed, CUDA events), then one
  gather of all detection events on rank 0 (timed, `sharding.gather_events`), with per-shard parity spot checks.
--config 4 (configs[3]): FFT 1024, hop 4, 162 bins x 8 columns = 1296 inputs -> H tansig -> 2 outputs (H = --hidden, default 256):
  the wide-hidden tcgen05 path; roofline against the measured TF32 tensor peak.

One JSON line on stdout (rank 0):
  value     audio-seconds processed per second, inputs resident in HBM (kernel launches only), all ranks together
  e2e       same metric through the public host API (syldet_batch_run_host): pinned host 16-bit PCM in (what a WAV corpus holds;
            converted on the device as x / 32768), debounced events out - H2D + kernels + D2H every step. e2e_f32: float32 PCM in.
  roofline  dominant kernel vs the measured HBM peak, algorithmic bytes = 4*hop + 4*outputs per evaluation
  cpu_baseline  the CPU oracle (oracle/oracle.c, a port: the Swift/Accelerate reference cannot be built on Linux), -O3 -march=native
                build, timed on this box's host cores on a bounded sample of the same workload (all cores, and one core)
  parity    EVERY evaluation of rank 0's recording against the oracle (outside the timed regions)
`--impl reference` times that CPU port alone, on all host threads, and prints the same line with "impl": "reference"; it imports
nothing of the product.
"""
import argparse
import ctypes
import importlib
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 44100
SAMPLE_TXT = os.path.join(ROOT, "tests", "golden", "sample.txt")
# other trained-net shapes for --shape (device-resident throughput + parity of whichever kernel the configuration qualifies for)
SHAPES = {
    "fft512_hop256_h8": dict(fft_len=512, overlap=256, freq_range=(1000.0, 9000.0), time_range=5, hidden=(8,), input_funcs=("l2normalize", "mapminmax")),
    "fft256_hop128_h8_minmax": dict(fft_len=256, overlap=128, freq_range=(1500.0, 6500.0), time_range=6, hidden=(8,),
                                    input_funcs=("normalize", "mapminmax"), transfer="LogSig"),
}
METRIC = "audio-sec/sec"
UNIT = "audio-seconds/second"


def workload_config(args, n_gpus):
    if args.config == 3:
        return {"workload": "sample.txt network over a %g-hour synthetic corpus of 1-hour %d-channel 44.1 kHz recordings, sharded by "
                            "recording over %d rank(s)" % (args.corpus_hours, args.channels, n_gpus),
                "corpus_hours": args.corpus_hours, "channels_per_recording": args.channels, "sampling_rate": FS,
                "sharding": "by recording (contiguous blocks), no collective on the data path; one event gather at the end",
                "l2": "every recording (%.2f GB) is far larger than L2" % (args.channels * 3600 * FS * 4 / 1e9)}
    return {"workload": "sample.txt network (FFT 256, hop 132, 29 bins x 10 columns -> 4 tansig -> 1) over a %g-hour %d-channel "
                        "44.1 kHz synthetic recording per GPU" % (args.hours, args.channels),
            "channels_per_gpu": args.channels, "seconds_per_channel": args.hours * 3600.0, "sampling_rate": FS,
            "sharding": "by recording, %d rank(s), no collective on the data path" % n_gpus,
            "l2": "inputs (%.2f GB per GPU) are far larger than L2, no flush needed" % (args.channels * args.hours * 3600 * FS * 4 / 1e9)}


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference(args, steps, warmup, single_core=False):
    """Times the CPU port of the reference path (oracle, timing build) with all host threads. One step = `reps` passes over
    n_threads channels x 120 s of the same synthetic audio.  Returns dict(value, cores, sample, ms_per_step, frames_per_s[, single])."""
    import oracle
    from tools import synth
    orc = oracle.Oracle(SAMPLE_TXT, fast=True)
    # every host core this process may run on - not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sec = 120
    x = synth.make_audio(threads, sec * FS, seed=123)
    t0 = time.perf_counter()
    orc.run_multi(x, n_threads=threads, want_outputs=False)
    one = time.perf_counter() - t0
    target = args.cpu_step_seconds
    reps = max(1, int(round(target / max(one, 1e-3))))
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        for _ in range(reps):
            orc.run_multi(x, n_threads=threads, want_outputs=False)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    audio_s = reps * threads * sec
    total = sum(times)
    r = {"value": audio_s * len(times) / total, "cores": threads,
         "sample": "%d passes over %d channels x %d s of the same synthetic audio per step, %d timed steps, OpenMP over channels; "
                   "gcc -O3 -march=native build of oracle/oracle.c" % (reps, threads, sec, len(times)),
         "ms_per_step": 1e3 * total / len(times), "frames_per_s": orc.num_evals(sec * FS) * reps * threads * len(times) / total}
    if single_core:   # the faithful analogue of the single-threaded reference CLI (main.swift:126-130): one channel, one thread
        orc.run_multi(x[:1], n_threads=1, want_outputs=False)
        t0 = time.perf_counter()
        n1 = 0
        while time.perf_counter() - t0 < 3.0:
            orc.run_multi(x[:1], n_threads=1, want_outputs=False)
            n1 += 1
        r["single"] = {"value": n1 * sec / (time.perf_counter() - t0), "cores": 1, "unit": UNIT,
                       "sample": "%d passes over 1 channel x %d s, one thread" % (n1, sec)}
    return r


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [p.strip() for p in line.split(",")]))

    def summary(self, t0, t1):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 6] or [r for _, r in self.rows if len(r) >= 6]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(r[0])) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(rows[0][1])), "reasons": reasons, "samples": len(rows)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def dram_traffic_per_launch(name, alg_bytes):
    """dram__bytes_read.sum + dram__bytes_write.sum of the active kernel from the committed `ncu --set full` capture, scaled from
    the captured workload to this launch by the ratio of algorithmic bytes (the capture is a shorter recording)."""
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            t = json.load(f)
        scale = alg_bytes / float(t["algorithmic_bytes"])
        t["bytes_per_launch"] = (t["dram_bytes_read"] + t["dram_bytes_write"]) * scale
        t["scaled_by"] = scale
        return t
    except Exception:
        return None


def stream_latency(args, channels, buffer, seconds, paced, resident=False):
    """BASELINE config 5: live channels with small buffers, sample.txt network; per-buffer latency (submit -> outputs and `seen`
    flags host-visible) measured by the C++ driver cli/syldet_stream_bench.cpp through syldet_stream_submit."""
    exe = os.path.join(ROOT, "syllable-detector-swift_b200", "syldet_stream_bench")
    try:
        env = dict(os.environ)
        if resident:   # the tick blocks stay on their SMs and poll a message in pinned memory
            env["SYLDET_STREAM_RESIDENT"] = "1"
        r = subprocess.run([exe, "-n", SAMPLE_TXT, "-c", str(channels), "-b", str(buffer), "-s", str(seconds), "-p", str(paced),
                            "-d", os.environ.get("LOCAL_RANK", "0")], capture_output=True, text=True, timeout=120 if resident else 600, env=env)
        if r.returncode != 0:
            return {"error": (r.stderr or r.stdout)[-300:]}
        return json.loads(r.stdout)
    except Exception as e:  # noqa: BLE001 - the bench line must still print
        return {"error": repr(e)[:300]}


# ---- NUMA placement of the pinned staging buffers (e2e leg) -----------------------------------------------------------
def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_near_gpu(torch, local):
    """Moves this process next to its GPU before the pinned host buffers are allocated: CPU affinity to the GPU's NUMA node when the
    cpuset allows it, and a preferred-node memory policy (set_mempolicy) in any case, so that cudaHostAlloc's pages and the PCIe
    root port of the GPU sit on the same socket. Returns a description for the bench line."""
    info = {"gpu_numa_node": None, "cpu_bound": False, "mem_policy": None}
    restore = {"cpus": None, "mem": False}

    def undo():
        """back to the original affinity / default memory policy (the CPU baseline and the oracle use every host core)"""
        try:
            if restore["cpus"] is not None:
                os.sched_setaffinity(0, restore["cpus"])
            if restore["mem"]:
                ctypes.CDLL(None).syscall(238, 0, None, ctypes.c_ulong(0))
        except Exception:  # noqa: BLE001
            pass

    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        info["gpu_numa_node"] = node
        info["pci"] = bdf
        if node < 0:
            return info, undo
        before = os.sched_getaffinity(0)
        info["cpus_before"] = len(before)
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            want = _parse_cpulist(f.read()) & before
        if want:
            os.sched_setaffinity(0, want)
            restore["cpus"] = before
            info["cpu_bound"] = True
            info["cpus_after"] = len(want)
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED = 1
        rc = libc.syscall(238, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))   # set_mempolicy (x86-64)
        info["mem_policy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy failed (errno %d)" % ctypes.get_errno()
        restore["mem"] = rc == 0
    except Exception as e:  # noqa: BLE001
        info["error"] = repr(e)[:200]
    return info, undo


def full_parity(np, orc, cfg, x_host, outs_gpu, ev_channel, ev_sample, tol=1e-5, scale_tol=False):
    """Every evaluation of a recording against the oracle: outputs within `tol`, decisions identical except evaluations whose oracle
    output lies within `tol` of the threshold (listed), events == the kernel's own decisions."""
    nch = x_host.shape[0]
    ref, da_ref = orc.run_multi(x_host, n_threads=len(os.sched_getaffinity(0)), want_outputs=True)
    thr = float(cfg.thresholds[0])
    if scale_tol:   # random networks (--shape) may have outputs far above 1: tolerance x output scale, as in the parity tests
        tol = tol * max(1.0, float(np.nanmax(np.abs(ref))) if ref.size else 1.0)
    err = np.abs(outs_gpu - ref)
    both_nan = np.isnan(outs_gpu) & np.isnan(ref)
    err[both_nan] = 0.0
    worst = float(np.nanmax(err)) if err.size else 0.0
    n_nan = int(np.isnan(err).sum())
    near = np.abs(ref[:, :, 0].astype(np.float64) - thr) <= tol
    with np.errstate(invalid="ignore"):
        da_gpu = outs_gpu[:, :, 0].astype(np.float64) >= thr
    flips = da_gpu != da_ref
    flips_far = int((flips & ~near).sum())
    first, hop = cfg.first_output_sample, cfg.hop
    ch_idx, j_idx = np.nonzero(da_gpu)
    want_samples = first + hop * j_idx
    events_ok = bool(len(ev_sample) == ch_idx.size and np.array_equal(ev_channel, ch_idx) and np.array_equal(ev_sample, want_samples))
    near_list = [{"channel": int(c), "evaluation": int(j), "oracle_output": float(ref[c, j, 0]), "gpu_output": float(outs_gpu[c, j, 0])}
                 for c, j in zip(*np.nonzero(near))][:32]
    return {"evaluations_checked": int(ref.shape[0] * ref.shape[1]), "channels": nch, "max_abs_err_vs_oracle": worst,
            "unmatched_nan": n_nan, "tolerance": tol, "near_threshold": int(near.sum()), "near_threshold_frames": near_list,
            "decision_flips": int(flips.sum()), "decision_flips_outside_near_band": flips_far,
            "oracle_detections": int(da_ref.sum()), "gpu_detections": int(da_gpu.sum()),
            "events_equal_own_decisions": events_ok,
            "timestamps_bit_exact": bool(flips_far == 0 and events_ok and n_nan == 0 and worst <= tol)}


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    sd = importlib.import_module("syldet_b200")
    sharding = importlib.import_module("syllable-detector-swift_b200.sharding")
    from tools import synth

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if args.config == 3:
        return run_corpus(args, np, torch, dist, sd, sharding, synth, rank, world, local, dev, barrier)
    if args.config == 4:
        return run_wide(args, np, torch, dist, sd, rank, world, local, dev, barrier)

    shape_text = None
    if args.shape != "sample":   # a second trained-net shape (device-resident numbers + parity only): what the other kernels do
        spec = importlib.util.spec_from_file_location("_cw", os.path.join(ROOT, "syllable-detector-swift_b200", "config_writer.py"))
        cwm = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(cwm)
        shape_text = cwm.random_config(seed=5, threshold=0.3, **SHAPES[args.shape])
        args.no_e2e = True
        args.no_alt = True
    cfg = (sd.SyllableDetectorConfig.from_text(shape_text) if shape_text else sd.SyllableDetectorConfig(SAMPLE_TXT)).validate()
    nch = args.channels
    n = int(round(args.hours * 3600 * FS))
    n -= n % 4
    E = cfg.num_evals(n)
    audio_seconds = nch * n / FS                      # per rank per step
    if shape_text:
        g = torch.Generator(device=dev)
        g.manual_seed(77 + rank)
        x = torch.empty((nch, n), dtype=torch.float32, device=dev).normal_(0.0, 0.05, generator=g)
    else:
        x = synth.make_audio_torch(nch, n, dev, seed=1000 + rank)   # this rank's recording, resident in HBM
    d_out = torch.empty((nch, E, cfg.net_outputs), dtype=torch.float32, device=dev)
    det = sd.BatchDetector(cfg, device=local, kernel=getattr(sd, "KERNEL_" + args.kernel.upper()))
    assert shape_text or det.active_kernel in (sd.KERNEL_FUSED, sd.KERNEL_TENSOR, sd.KERNEL_TENSOR_TF32), "sample.txt must take a fused kernel"
    tf32_env = bool(os.environ.get("SYLDET_TC_TF32_CORR"))
    kernel_name = {sd.KERNEL_FUSED: "fused_detect_kernel<256,4> (SIMT FFT)",
                   sd.KERNEL_TENSOR_TF32: "tc_detect_kernel<4,kFast> (tcgen05 3xTF32 band DFT)",
                   sd.KERNEL_TENSOR: ("tc_detect_kernel<4,kFast> (tcgen05 3xTF32 band DFT)" if tf32_env else
                                      "tc_detect_kernel<4,kFast,kF16> (tcgen05 band DFT on two-term fp16 splits, range-guarded; layer 0 as 3xTF32)")}.get(det.active_kernel, sd.KERNEL_NAMES.get(det.active_kernel))
    if shape_text:
        kernel_name = "%s kernel on shape '%s' %s" % (sd.KERNEL_NAMES.get(det.active_kernel), args.shape, SHAPES[args.shape])
    stream = torch.cuda.current_stream(dev)

    def launch(d=det, outs=d_out):
        d.launch_device(x.data_ptr(), nch, n, n, detect_rule=sd.DETECT_ANY_OUTPUT, d_outputs_ptr=outs.data_ptr() if outs is not None else None,
                        stream=stream.cuda_stream)

    # ---- device-resident throughput: K launches bracketed by barrier + synchronize; per-launch CUDA events -------------
    for _ in range(args.warmup):
        launch()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = det.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_clock0 = time.perf_counter()
    barrier()
    t0 = time.perf_counter()
    for a, b in evs:
        a.record(stream)
        launch()
        b.record(stream)
    barrier()
    t1 = time.perf_counter()
    wall = t1 - t0
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = evs[0][0].elapsed_time(evs[-1][1])       # device time of the whole timed region on the launching stream
    gpu_launches = det.launch_count - launches0
    n_det = det.last_detection_count()                # settles the launch (range flag / event buffer), see include/syldet.h
    events = det.collect(debounce_frames=0)
    range_fallbacks = det.range_fallbacks

    # ---- the all-TF32 variant of the same kernel, same recording (reported beside the default) ---------------------------
    alt = None
    if det.active_kernel == sd.KERNEL_TENSOR and not tf32_env and not args.no_alt:
        det2 = sd.BatchDetector(cfg, device=local, kernel=sd.KERNEL_TENSOR_TF32)
        for _ in range(3):
            launch(det2, None)
        torch.cuda.synchronize(dev)
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(3, args.steps // 2))]
        for a, b in ev2:
            a.record(stream)
            launch(det2, None)
            b.record(stream)
        torch.cuda.synchronize(dev)
        alt = sum(a.elapsed_time(b) for a, b in ev2) / len(ev2)
        del det2

    if args.no_e2e:   # development runs (tools/*.sh): device-resident numbers and a parity check only
        if rank == 0:
            import oracle
            orc = oracle.Oracle(text=shape_text) if shape_text else oracle.Oracle(SAMPLE_TXT)
            seg_n = n if not args.quick_parity else min(n, 300 * FS)
            keep = events.sample < cfg.first_output_sample + cfg.hop * cfg.num_evals(seg_n)
            parity = full_parity(np, orc, cfg, x[:, :seg_n].cpu().numpy(), d_out[:, :cfg.num_evals(seg_n)].cpu().numpy(),
                                 events.channel[keep], events.sample[keep], scale_tol=shape_text is not None)
            parity.pop("near_threshold_frames")
            peaks, _ = measured_peaks()
            k_ms = sum(kernel_ms) / len(kernel_ms)
            alg_bytes = (4 * cfg.hop + 4 * cfg.net_outputs) * E * nch
            print(json.dumps({"dev_only": True, "kernel": kernel_name, "value": audio_seconds * world * args.steps / wall,
                              "roofline": {"kernel_ms": k_ms, "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]),
                                           "tf32_kernel_ms": alt}, "detections_per_step": int(n_det), "range_fallbacks": int(range_fallbacks),
                              "parity": parity, "clocks": clocks.summary(t_clock0, time.perf_counter())}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- end to end through the host API: pinned host PCM -> events --------------------------------------------------------
    numa, numa_undo = bind_near_gpu(torch, local)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    table = sharding.EventTable(cfg.net_outputs) if world > 1 else None   # this rank's detections: page-locked, filled in place

    def timed_e2e(h_np):
        """e2e_steps recordings per rank through the host API, then the job's one gather of every rank's detections on rank 0"""
        for _ in range(max(1 if world > 1 else 0, min(args.warmup, 2))):   # (N > 1: the gather warm-up and the final check need one)
            ev_h = det.run(h_np)
        if world > 1:   # warm-up of the gather as well: the first point-to-point gather sets up NCCL's peer connections (~0.4 s, once per job)
            table.clear()
            for step in range(e2e_steps):
                table.append(rank * e2e_steps + step, ev_h.channel, ev_h.sample, ev_h.outputs)
            sharding.gather_events(table, dist)
        barrier()
        te0 = time.perf_counter()
        if world > 1:   # the detections go straight into this rank's gather table (run_into), then ONE gather for the job
            table.clear()
            for step in range(e2e_steps):
                n_last = det.run_into(table, rank * e2e_steps + step, h_np)
            sharding.gather_events(table, dist)
        else:
            for step in range(e2e_steps):
                ev_h = det.run(h_np)
        barrier()
        wall = time.perf_counter() - te0
        if world > 1:   # the last step's rows, as the caller's checks want them
            _, ch_l, smp_l, out_l = sharding.unpack_events(table.rows[table.n - n_last:])
            assert table.in_order and len(ev_h) == n_last and np.array_equal(ev_h.sample, smp_l) and np.array_equal(ev_h.channel, ch_l) \
                and np.array_equal(ev_h.outputs, out_l)
        return wall, ev_h

    # 16-bit PCM (what a WAV corpus holds; converted on the device as x/32768): the e2e headline
    q = torch.clamp(torch.round(x * 32768.0), -32768, 32767).to(torch.int16)
    h16 = torch.empty((nch, n), dtype=torch.int16, pin_memory=True)
    h16.copy_(q)
    xq = q.to(torch.float32) / 32768.0             # what the device sees after the conversion
    # bare copy ceiling: the same pinned buffer -> device, alone and with all ranks copying at once
    torch.cuda.synchronize(dev)
    pcie = []
    for _ in range(3):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        q.copy_(h16, non_blocking=True)
        b.record(stream)
        torch.cuda.synchronize(dev)
        pcie.append(nch * n * 2 / (a.elapsed_time(b) * 1e-3) / 1e9)
    pcie_gbs = max(pcie)
    del q
    det.launch_device(xq.data_ptr(), nch, n, n, detect_rule=sd.DETECT_ANY_OUTPUT, d_outputs_ptr=None, stream=stream.cuda_stream)
    ev_q = det.collect(debounce_frames=0)
    del xq
    torch.cuda.synchronize(dev)
    e2e16_wall, ev16 = timed_e2e(h16.numpy())
    assert len(ev16) == len(ev_q) and np.array_equal(ev16.sample, ev_q.sample) and np.array_equal(ev16.channel, ev_q.channel)
    d2h16 = len(ev16) * (16 + 4 * cfg.net_outputs) + 16
    del h16

    # float32 PCM
    e2e32 = None
    h_np = None
    if not args.no_f32_e2e:
        h = torch.empty((nch, n), dtype=torch.float32, pin_memory=True)
        h.copy_(x)
        torch.cuda.synchronize(dev)
        h_np = h.numpy()
        e2e32_wall, ev_h = timed_e2e(h_np)
        assert len(ev_h) == len(events) and np.array_equal(ev_h.sample, events.sample)
        e2e32 = (e2e32_wall, len(ev_h) * (16 + 4 * cfg.net_outputs) + 16)
    t_clock1 = time.perf_counter()
    numa_undo()

    # ---- parity against the oracle, outside the timed regions -----------------------------------------------------------------
    import oracle
    orc = oracle.Oracle(SAMPLE_TXT)
    if world == 1 and not args.quick_parity:
        x_host = h_np if h_np is not None else x.cpu().numpy()
        parity = full_parity(np, orc, cfg, x_host, d_out.cpu().numpy(), events.channel, events.sample)
    else:   # every rank checks two minutes of each channel of its own recording (per-shard parity); rank 0 reports the worst
        seg_n = min(n, 120 * FS)
        keep = events.sample < cfg.first_output_sample + cfg.hop * cfg.num_evals(seg_n)
        parity = full_parity(np, orc, cfg, x[:, :seg_n].cpu().numpy(), d_out[:, :cfg.num_evals(seg_n)].cpu().numpy(),
                             events.channel[keep], events.sample[keep])
        parity["scope"] = "first %d s of every channel of every rank's recording" % (seg_n // FS)
        if world > 1:
            t = torch.tensor([parity["max_abs_err_vs_oracle"], parity["decision_flips_outside_near_band"], parity["near_threshold"],
                              0.0 if parity["timestamps_bit_exact"] else 1.0, parity["evaluations_checked"]], dtype=torch.float64, device=dev)
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            parity.update({"max_abs_err_vs_oracle": float(tm[0]), "decision_flips_outside_near_band": int(t[1]), "near_threshold": int(t[2]),
                           "timestamps_bit_exact": bool(tm[3] == 0.0), "evaluations_checked": int(t[4]), "ranks_checked": world})
    del h_np

    # ---- reduce over ranks: max time, summed units ------------------------------------------------------------------------
    t = torch.tensor([wall, e2e16_wall, dev_ms, e2e32[0] if e2e32 else 0.0, -pcie_gbs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, e2e16_wall, dev_ms, e2e32_wall, pcie_min = [float(v) for v in t.tolist()]
    pcie_min = -pcie_min
    total_audio = audio_seconds * world
    value = total_audio * args.steps / wall

    if rank == 0:
        peaks, peak_src = measured_peaks()
        peak = float(peaks["hbm_gbs"])
        k_ms = sum(kernel_ms) / len(kernel_ms)
        alg_bytes = (4 * cfg.hop + 4 * cfg.net_outputs) * E * nch          # per launch, this rank
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        tensor = det.active_kernel != sd.KERNEL_FUSED
        traffic = dram_traffic_per_launch("tc_traffic.json" if tensor else "fused_traffic.json", alg_bytes)
        h2d16 = nch * n * 2
        e2e_ms = 1e3 * e2e16_wall / e2e_steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (Gaussian noise + synthetic syllables, generated on device; sample.txt weights)",
            "config": workload_config(args, world),
            "frames_per_s": E * nch * world * args.steps / wall,
            "device_ms_per_step": dev_ms / args.steps,
            "e2e": {"value": total_audio * e2e_steps / e2e16_wall, "unit": UNIT, "h2d_bytes_per_step": h2d16, "d2h_bytes_per_step": d2h16,
                    "steps": e2e_steps, "ms_per_step": e2e_ms, "input": "pinned host 16-bit PCM (the recording quantised as a WAV file holds it)",
                    "api": "syldet_batch_run_host with SYLDET_PCM_S16 (host PCM in, debounced events out; time-sliced copy/ingest/detect/collect pipeline)"
                           + ("; then ONE sharding.gather_events of all steps' detections to rank 0, inside the timed region" if world > 1 else ""),
                    "h2d_gbs_per_gpu": h2d16 / (e2e_ms * 1e-3) / 1e9,
                    "pcie_peak_gbs": pcie_min, "pcie_frac": h2d16 / (e2e_ms * 1e-3) / 1e9 / pcie_min,
                    "pcie_peak_source": "bare cudaMemcpyAsync of the same pinned buffer, all %d rank(s) copying at once, slowest rank, best of 3" % world,
                    "numa": numa},
            "gpu_launches": int(gpu_launches),
            "kernel": kernel_name,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["bytes_per_launch"] if traffic else None, "traffic_source": traffic,
                         "peak_source": peak_src + ", burst copy", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "532 B per evaluation (4*hop audio read once + 4*outputs written); FP32 work is 9572 FLOP per "
                                 "evaluation => %.2f TFLOP/s achieved" % (9572.0 * E * nch / (k_ms * 1e-3) / 1e12)},
            "detections_per_step": int(n_det), "events_per_step": len(events), "range_fallbacks": int(range_fallbacks), "parity": parity,
            "clocks": clocks.summary(t_clock0, t_clock1),
            "nccl": "barrier + max-reduction of the timings" + ("; one gather of the detection events to rank 0 at the end of the e2e job" if world > 1 else "") + "; no collective on the data path",
        }
        if alt is not None:
            line["roofline"]["variants"] = {
                "tensor (default; fp16 correction pass, range-guarded)": {"kernel_ms": k_ms, "frac": achieved / peak},
                "tensor_tf32 (all three DFT products in TF32)": {"kernel_ms": alt, "frac": alg_bytes / (alt * 1e-3) / 1e9 / peak}}
        if e2e32:
            ms32 = 1e3 * e2e32_wall / e2e_steps
            line["e2e_f32"] = {"value": total_audio * e2e_steps / e2e32_wall, "unit": UNIT, "h2d_bytes_per_step": nch * n * 4,
                               "d2h_bytes_per_step": e2e32[1], "steps": e2e_steps, "ms_per_step": ms32,
                               "h2d_gbs_per_gpu": nch * n * 4 / (ms32 * 1e-3) / 1e9, "pcie_frac": nch * n * 4 / (ms32 * 1e-3) / 1e9 / pcie_min,
                               "api": "syldet_batch_run_host with SYLDET_PCM_F32 (pinned host float32 PCM in)"}
        if world == 1 and not args.no_cpu:
            cb = cpu_reference(args, steps=1, warmup=0, single_core=True)
            line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port", "sample": cb["sample"],
                                    "frames_per_s": cb["frames_per_s"], "single_core": cb.get("single")}
        if world == 1 and not args.no_stream:
            main = stream_latency(args, args.stream_channels, args.stream_buffer, args.stream_seconds, args.stream_paced_seconds)
            main["api"] = ("syldet_stream_submit (one stream_tick_fast_kernel launch per tick that completes an STFT column; samples pulled from and "
                           "outputs written to pinned host memory by the kernel)")
            res = stream_latency(args, args.stream_channels, args.stream_buffer, args.stream_seconds, args.stream_paced_seconds, resident=True)
            res["api"] = ("syldet_stream_submit with SYLDET_STREAM_RESIDENT=1 (opt-in): no launch per tick - the channel blocks stay on their SMs, a "
                          "dispatcher warp polls the tick message in pinned host memory")
            main["resident"] = res
            if not args.no_stream_sweep:   # BASELINE config 5's other points: 128- / 256-frame buffers, 1024 channels (shorter runs)
                main["sweep"] = {"%dch_x_%dframes" % (c, b): stream_latency(args, c, b, 15.0, 2.0)
                                 for c, b in ((64, 128), (64, 256), (1024, 32))}
            line["stream"] = main
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_corpus(args, np, torch, dist, sd, sharding, synth, rank, world, local, dev, barrier):
    """BASELINE config 3: the per-file loop of SyllableDetectorCLI/main.swift:63-131 over a corpus sharded by recording."""
    import oracle
    cfg = sd.SyllableDetectorConfig(SAMPLE_TXT).validate()
    orc = oracle.Oracle(SAMPLE_TXT)
    nch = args.channels
    n = 3600 * FS
    E = cfg.num_evals(n)
    n_rec = max(1, int(round(args.corpus_hours / nch)))          # recordings of 1 h x nch channels
    a, b = sharding.partition([n] * n_rec, world)[rank]
    det = sd.BatchDetector(cfg, device=local, kernel=getattr(sd, "KERNEL_" + args.kernel.upper()))
    stream = torch.cuda.current_stream(dev)
    rows = []
    dev_ms = 0.0
    wall = 0.0
    worst, flips_far, checked = 0.0, 0, 0
    launches0 = det.launch_count
    rng = np.random.default_rng(rank)
    d_out = torch.empty((nch, E, cfg.net_outputs), dtype=torch.float32, device=dev)
    # warm-up on the first recording of the shard
    x = synth.make_audio_torch(nch, n, dev, seed=5000 + a)
    for _ in range(max(3, args.warmup)):
        det.launch_device(x.data_ptr(), nch, n, n, d_outputs_ptr=None, stream=stream.cuda_stream)
    ev_w = det.collect()
    if world > 1:   # the first point-to-point gather sets up NCCL's peer connections: once per job, not part of the corpus
        sharding.gather_events(sharding.pack_events_compact(rank, ev_w.channel, ev_w.sample, ev_w.outputs), dist)
    table = sharding.EventTable(cfg.net_outputs, capacity=max(1 << 20, int(1.25 * len(ev_w) * (b - a))))   # page-locked once, before the clock starts
    if world > 1 and rank == 0:   # and so is rank 0's table of everybody's rows (sized from the warm-up recording)
        sharding.reserve_gather(cfg.net_outputs, int(1.25 * len(ev_w) * n_rec))
    del ev_w
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    t_clock0 = time.perf_counter()
    for rec in range(a, b):
        x = synth.make_audio_torch(nch, n, dev, seed=5000 + rec)     # device synthesis: input creation, not timed
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        det.launch_device(x.data_ptr(), nch, n, n, d_outputs_ptr=d_out.data_ptr(), stream=stream.cuda_stream)
        e1.record(stream)
        t_l = time.perf_counter()
        ev = det.collect(debounce_frames=0)                        # synchronises; events to the host
        wall += time.perf_counter() - t0
        if os.environ.get("SYLDET_E2E_TIMING") and rec < a + 8:
            sys.stderr.write("[c3] recording %d: launch call %.2f ms, collect %.2f ms\n" % (rec, 1e3 * (t_l - t0), 1e3 * (time.perf_counter() - t_l)))
        dev_ms += e0.elapsed_time(e1)
        table.append(rec, ev.channel, ev.sample, ev.outputs)
        # per-shard parity: 400 evaluations of a random channel / offset of this recording against the oracle
        ch, j = int(rng.integers(nch)), int(rng.integers(E - 400))
        seg = x[ch, j * cfg.hop: j * cfg.hop + cfg.first_output_sample + cfg.hop * 399].cpu().numpy()
        ref, da, _ = orc.run(seg)
        got = d_out[ch, j:j + 400].cpu().numpy()
        worst = max(worst, float(np.abs(got - ref).max()))
        near = np.abs(ref[:, 0].astype(np.float64) - cfg.thresholds[0]) <= 1e-5
        flips_far += int((((got[:, 0].astype(np.float64) >= cfg.thresholds[0]) != da) & ~near).sum())
        checked += 400
    rows = table.rows
    barrier()
    t0 = time.perf_counter()
    allrows = sharding.gather_events(table, dist if world > 1 else None)     # the final host gather of detection timestamps
    barrier()
    gather_s = time.perf_counter() - t0
    t_clock1 = time.perf_counter()
    t = torch.tensor([wall + gather_s, dev_ms, worst, float(flips_far)], dtype=torch.float64, device=dev)
    s = torch.tensor([float(checked), float(det.launch_count - launches0), float(rows.shape[0])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    total_s, dev_ms_max, worst, flips_far = [float(v) for v in t.tolist()]
    if rank == 0:
        audio = n_rec * nch * n / FS
        peaks, peak_src = measured_peaks()
        alg = (4 * cfg.hop + 4 * cfg.net_outputs) * E * nch * (b - a)
        line = {"metric": METRIC, "value": audio / total_s, "unit": UNIT, "n_gpus": world, "steps": n_rec, "warmup": max(3, args.warmup),
                "ms_per_step": 1e3 * total_s / max(1, b - a), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic (Gaussian noise + synthetic syllables, one seed per recording, generated on device; sample.txt weights)",
                "config": workload_config(args, world),
                "recordings": n_rec, "recordings_on_slowest_rank": int(b - a), "seconds_total": total_s, "gather_seconds": gather_s,
                "device_ms_slowest_rank": dev_ms_max, "value_device_only": audio / (dev_ms_max * 1e-3),
                "events_gathered": int(allrows.shape[0]) if allrows is not None else 0, "events_all_ranks": int(s[2]),
                "gpu_launches": int(s[1]),
                "roofline": {"bound": "hbm", "achieved": alg / (dev_ms_max * 1e-3) / 1e9, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                             "frac": alg / (dev_ms_max * 1e-3) / 1e9 / float(peaks["hbm_gbs"]), "traffic": None, "peak_source": peak_src,
                             "note": "rank 0's recordings over its summed launch durations"},
                "parity": {"evaluations_checked": int(s[0]), "max_abs_err_vs_oracle": worst, "decision_flips_outside_near_band": int(flips_far),
                           "scope": "400 evaluations at a random channel / offset of every recording of every shard"},
                "clocks": clocks.summary(t_clock0, t_clock1),
                "nccl": "barrier, reductions of the timings, one all_gather of the detection events (sharding.gather_events); no collective on the data path"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def wide_config_text(hidden):
    """BASELINE config 4 as SURVEY.md 8(d) spells it out: fs 44 100, FFT = window = 1024, hop 4 (overlap 1020), band 1-8 kHz (162 bins),
    T = 8 (1296 inputs), `hidden` tansig units, 2 purelin outputs; seeded random weights in the reference's text format. config_writer
    is loaded by path: the reference arm must not import the product package."""
    spec = importlib.util.spec_from_file_location("_cw", os.path.join(ROOT, "syllable-detector-swift_b200", "config_writer.py"))
    cw = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cw)
    return cw.random_config(seed=21, fft_len=1024, overlap=1020, freq_range=(1000.0, 8000.0), time_range=8, hidden=(hidden,), outputs=2,
                            threshold=0.2)


def wide_audio(np, nch, n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    return np.stack([(0.05 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * (2500.0 + 450.0 * ch) * t / FS + 3 * np.sin(2 * np.pi * 2 * t / FS))).astype(np.float32)
                     for ch in range(nch)])


def wide_workload(args, hidden, n_gpus):
    return {"workload": "high-overlap wide-hidden network (FFT 1024, hop 4, 162 bins x 8 columns = 1296 inputs -> %d tansig -> 2) over a "
                        "%g-second %d-channel 44.1 kHz synthetic recording per GPU" % (hidden, args.wide_seconds, args.channels),
            "hidden": hidden, "channels_per_gpu": args.channels, "seconds_per_channel": args.wide_seconds, "sampling_rate": FS,
            "sharding": "by recording, %d rank(s), no collective on the data path" % n_gpus,
            "l2": "band-magnitude planes of a step (%.1f GB) and the audio are far larger than L2" % (args.channels * args.wide_seconds * FS / 4 * 1344 / 1e9)}


def run_wide(args, np, torch, dist, sd, rank, world, local, dev, barrier):
    """BASELINE config 4: the wide-hidden tensor path (stft_planes_kernel + wide_l0_kernel)."""
    import oracle
    H = args.hidden
    text = wide_config_text(H)
    cfg = sd.SyllableDetectorConfig.from_text(text).validate()
    nch = args.channels
    n = int(round(args.wide_seconds * FS))
    n -= n % 4
    E = cfg.num_evals(n)
    audio_seconds = nch * n / FS
    x_host = wide_audio(np, nch, n, 900 + rank)
    x = torch.from_numpy(x_host).to(dev)
    d_out = torch.empty((nch, E, cfg.net_outputs), dtype=torch.float32, device=dev)
    det = sd.BatchDetector(cfg, device=local)
    assert det.active_kernel == sd.KERNEL_WIDE, "config 4 must take the wide-hidden tensor path"
    stream = torch.cuda.current_stream(dev)

    def launch():
        det.launch_device(x.data_ptr(), nch, n, n, detect_rule=sd.DETECT_ANY_OUTPUT, d_outputs_ptr=d_out.data_ptr(), stream=stream.cuda_stream)

    for _ in range(args.warmup):
        launch()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = det.launch_count
    t_clock0 = time.perf_counter()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    phases = []
    for _ in range(args.steps):
        launch()
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    gpu_launches = det.launch_count - launches0
    phases = det.wide_phase_ms()                     # the last step's two kernels
    n_det = det.last_detection_count()
    events = det.collect(debounce_frames=0)

    # end to end through the host API: pinned host float32 PCM -> events
    h = torch.empty((nch, n), dtype=torch.float32, pin_memory=True)
    h.copy_(x)
    torch.cuda.synchronize(dev)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ev_h = det.run(h.numpy())
    barrier()
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        ev_h = det.run(h.numpy())
    barrier()
    e2e_wall = time.perf_counter() - te0
    t_clock1 = time.perf_counter()
    assert len(ev_h) == len(events) and np.array_equal(ev_h.sample, events.sample)

    # parity: the first `ps` seconds of every channel against the oracle (all host threads), every evaluation in it
    orc = oracle.Oracle(text=text)
    ps = int(min(n, args.wide_parity_seconds * FS))
    Ep = cfg.num_evals(ps)
    ref, da_ref = orc.run_multi(x_host[:, :ps], n_threads=len(os.sched_getaffinity(0)), want_outputs=True)
    got = d_out[:, :Ep].cpu().numpy()
    scale = max(1.0, float(np.nanmax(np.abs(ref))))
    tol = 1e-5 * scale
    thr = cfg.thresholds
    near = (np.abs(ref.astype(np.float64) - thr[None, None, :]) <= tol).any(axis=2)
    da_gpu = (got.astype(np.float64) >= thr[None, None, :]).any(axis=2)
    flips = da_gpu != da_ref
    parity = {"evaluations_checked": int(ref.shape[0] * ref.shape[1]), "max_abs_err_vs_oracle": float(np.abs(got - ref).max()), "tolerance": tol,
              "near_threshold": int(near.sum()), "decision_flips": int(flips.sum()), "decision_flips_outside_near_band": int((flips & ~near).sum()),
              "scope": "first %.1f s of every channel" % (ps / FS)}

    t = torch.tensor([wall, e2e_wall, dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, e2e_wall, dev_ms = [float(v) for v in t.tolist()]
    if rank == 0:
        peaks, peak_src = measured_peaks()
        try:
            with open(os.path.join(ROOT, "profiles", "r02_tensor_peaks.json")) as f:
                tp = json.load(f)
            tf32_peak, tf32_src = float(tp["tf32_tflops"]), "measured (profiles/r02_tensor_peaks.json: tools/mma_rate.cu --peak, burst)"
        except Exception:
            tf32_peak, tf32_src = float(peaks["bf16_tflops"]) / 2.0, peak_src + ": bf16 burst / 2 (no TF32 measurement on file)"
        I = cfg.net_inputs
        alg_flops = 2.0 * I * H * E * nch                       # layer 0 as the reference computes it (NeuralNet.swift:366-377), per launch
        l0_ms, stft_ms = phases[1], phases[0]
        achieved = alg_flops / (l0_ms * 1e-3) / 1e12
        total_audio = audio_seconds * world
        line = {
            "metric": METRIC, "value": total_audio * args.steps / wall, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xTF32 tensor products)",
            "data": "synthetic (noise + frequency-modulated tones; seeded random weights in the reference's text format)",
            "config": wide_workload(args, H, world),
            "frames_per_s": E * nch * world * args.steps / wall, "device_ms_per_step": dev_ms / args.steps,
            "e2e": {"value": total_audio * e2e_steps / e2e_wall, "unit": UNIT, "h2d_bytes_per_step": nch * n * 4,
                    "d2h_bytes_per_step": len(ev_h) * (16 + 4 * cfg.net_outputs) + 16, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_wall / e2e_steps,
                    "api": "syldet_batch_run_host (pinned host float32 PCM in, debounced events out)"},
            "gpu_launches": int(gpu_launches),
            "kernel": "wide_l0_kernel (tcgen05 3xTF32, [evaluations x 1296] . [1296 x %d], sliding A operand, TMEM accumulators) after stft_planes_kernel" % H,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak, "traffic": None,
                         "peak_source": tf32_src, "kernel_ms": l0_ms, "stft_kernel_ms": stft_ms,
                         "algorithmic_flops_per_launch": alg_flops,
                         "note": "algorithmic FLOPs = 2 * inputs * hidden per evaluation (layer 0); a 3xTF32 product issues 3 MMA FLOPs per algorithmic "
                                 "FLOP over K padded from 1296 to 1344, so frac <= 0.32; mma_frac counts the issued MMA FLOPs"},
            "detections_per_step": int(n_det), "events_per_step": len(events), "parity": parity,
            "clocks": clocks.summary(t_clock0, t_clock1),
        }
        line["roofline"]["mma_frac"] = 3.0 * achieved * (8 * 168.0 / I) / tf32_peak
        if world == 1 and not args.no_cpu:
            threads = len(os.sched_getaffinity(0))
            of = oracle.Oracle(text=text, fast=True)
            xs = wide_audio(np, threads, int(args.wide_cpu_seconds * FS), 77)
            of.run_multi(xs[:, :FS // 10], n_threads=threads, want_outputs=False)
            tc0 = time.perf_counter()
            of.run_multi(xs, n_threads=threads, want_outputs=False)
            dt = time.perf_counter() - tc0
            line["cpu_baseline"] = {"value": threads * args.wide_cpu_seconds / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d channels x %.2f s of the same kind of audio, one pass, OpenMP over channels; gcc -O3 -march=native build "
                                              "of oracle/oracle.c" % (threads, args.wide_cpu_seconds)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference_wide(args):
    import numpy as np
    import oracle
    world = int(os.environ.get("WORLD_SIZE", "1"))
    text = wide_config_text(args.hidden)
    of = oracle.Oracle(text=text, fast=True)
    threads = len(os.sched_getaffinity(0))
    xs = wide_audio(np, threads, int(args.wide_cpu_seconds * FS), 77)
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        of.run_multi(xs, n_threads=threads, want_outputs=False)
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    v = threads * args.wide_cpu_seconds * len(times) / sum(times)
    sample = "%d channels x %.2f s per step, OpenMP over channels; gcc -O3 -march=native build of oracle/oracle.c" % (threads, args.wide_cpu_seconds)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic (noise + frequency-modulated tones; seeded random weights)", "config": wide_workload(args, args.hidden, world),
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "note": "CPU port (oracle/oracle.c) of the reference's Swift/Accelerate path, which cannot be built on Linux"}), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config == 4:
        return run_reference_wide(args)
    cb = cpu_reference(args, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (Gaussian noise + synthetic syllables; sample.txt weights)",
            "config": workload_config(args, world), "frames_per_s": cb["frames_per_s"],
            "cpu_baseline": {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port", "sample": cb["sample"]},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU port (oracle/oracle.c, gcc -O3 -march=native) of the reference's Swift/Accelerate path, which cannot be built on Linux"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4])
    ap.add_argument("--hours", type=float, default=1.0)
    ap.add_argument("--channels", type=int, default=8)
    ap.add_argument("--corpus-hours", type=float, default=1000.0)
    ap.add_argument("--shape", default="sample", choices=["sample"] + sorted(SHAPES))
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--wide-seconds", type=float, default=30.0)
    ap.add_argument("--wide-parity-seconds", type=float, default=1.0)
    ap.add_argument("--wide-cpu-seconds", type=float, default=0.5)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-step-seconds", type=float, default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-f32-e2e", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-alt", action="store_true")
    ap.add_argument("--quick-parity", action="store_true")
    ap.add_argument("--kernel", default="auto", choices=["auto", "fused", "tensor", "tensor_tf32"])
    ap.add_argument("--no-stream", action="store_true")
    ap.add_argument("--no-stream-sweep", action="store_true")
    ap.add_argument("--stream-channels", type=int, default=64)
    ap.add_argument("--stream-buffer", type=int, default=32)
    ap.add_argument("--stream-seconds", type=float, default=60.0)
    ap.add_argument("--stream-paced-seconds", type=float, default=5.0)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.cpu_step_seconds is None:
        a.cpu_step_seconds = 12.0 if a.impl == "ours" else 3.0
    # The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version line at communicator creation),
    # so file descriptor 1 points at stderr while the benchmark runs and the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)

    def _emit(text, **_kw):
        os.write(_real_stdout, (text + "\n").encode())

    print = _emit  # noqa: A001 - run_ours / run_reference look the name up in the module namespace
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
