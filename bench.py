#!/usr/bin/env python
"""Benchmark of the detection hot path (STFT -> band -> sliding window -> MLP -> detect) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--hours H] [--channels C]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): the sample.txt network over a 1-hour, 8-channel, 44.1 kHz synthetic recording per
GPU (noise + synthetic syllables, generated on the device).  One step = one pass of the whole recording through the hot
path.  With N > 1 every rank owns its own recording (sharding by recording, no collective on the data path): weak scaling.

One JSON line on stdout (rank 0):
  value     audio-seconds processed per second, inputs resident in HBM (kernel launches only), all ranks together
  e2e       same metric through the public host API: pinned host PCM in, events out (H2D + kernels + D2H every step)
  roofline  fused kernel vs the measured HBM peak, algorithmic bytes = 4*hop + 4*outputs per evaluation
  cpu_baseline  the CPU oracle (oracle/oracle.c, a port: the Swift/Accelerate reference cannot be built on Linux),
                timed on this box's host cores on a bounded sample of the same workload
`--impl reference` times that CPU port alone, on all host threads, and prints the same line with "impl": "reference".
"""
import argparse
import importlib
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
METRIC = "audio-sec/sec"
UNIT = "audio-seconds/second"


def workload_config(args, n_gpus):
    return {"workload": "sample.txt network (FFT 256, hop 132, 29 bins x 10 columns -> 4 tansig -> 1) over a %g-hour %d-channel "
                        "44.1 kHz synthetic recording per GPU" % (args.hours, args.channels),
            "channels_per_gpu": args.channels, "seconds_per_channel": args.hours * 3600.0, "sampling_rate": FS,
            "sharding": "by recording, %d rank(s), no collective" % n_gpus,
            "l2": "inputs (%.2f GB per GPU) are far larger than L2, no flush needed" % (args.channels * args.hours * 3600 * FS * 4 / 1e9)}


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference(args, steps, warmup, calibrate_s=0.0):
    """Times the CPU port of the reference path (oracle) with all host threads. One step = `reps` passes over
    n_threads channels x 120 s of the same synthetic audio.  Returns dict(value, cores, sample, ms_per_step, frames)."""
    import numpy as np
    import oracle
    oracle.build()
    synth = importlib.import_module("syllable-detector-swift_b200.synth")
    orc = oracle.Oracle(SAMPLE_TXT)
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
    return {"value": audio_s * len(times) / total, "cores": threads,
            "sample": "%d passes over %d channels x %d s of the same synthetic audio per step, %d timed steps, OpenMP over channels"
                      % (reps, threads, sec, len(times)),
            "ms_per_step": 1e3 * total / len(times), "frames_per_s": orc.num_evals(sec * FS) * reps * threads * len(times) / total}


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


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def dram_traffic_per_launch(tensor, alg_bytes):
    """dram__bytes_read.sum + dram__bytes_write.sum of the active kernel from the committed `ncu --set full` capture, scaled from
    the captured workload to this launch by the ratio of algorithmic bytes (the capture is a shorter recording)."""
    try:
        with open(os.path.join(ROOT, "profiles", "tc_traffic.json" if tensor else "fused_traffic.json")) as f:
            t = json.load(f)
        scale = alg_bytes / float(t["algorithmic_bytes"])
        t["bytes_per_launch"] = (t["dram_bytes_read"] + t["dram_bytes_write"]) * scale
        t["scaled_by"] = scale
        return t
    except Exception:
        return None


def stream_latency(args):
    """BASELINE config 5: 64 live channels, 32-frame buffers, sample.txt network; per-buffer latency (submit -> outputs and
    `seen` flags host-visible) measured by the C++ driver cli/syldet_stream_bench.cpp through syldet_stream_submit."""
    import subprocess
    exe = os.path.join(ROOT, "syllable-detector-swift_b200", "syldet_stream_bench")
    try:
        r = subprocess.run([exe, "-n", SAMPLE_TXT, "-c", str(args.stream_channels), "-b", str(args.stream_buffer),
                            "-s", str(args.stream_seconds), "-p", str(args.stream_paced_seconds),
                            "-d", os.environ.get("LOCAL_RANK", "0")], capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            return {"error": (r.stderr or r.stdout)[-300:]}
        d = json.loads(r.stdout)
        d["api"] = "syldet_stream_submit (one stream_tick_kernel launch per tick that completes an STFT column; samples pulled from and outputs written to pinned host memory by the kernel)"
        return d
    except Exception as e:  # noqa: BLE001 - the bench line must still print
        return {"error": repr(e)[:300]}


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
    synth = importlib.import_module("syllable-detector-swift_b200.synth")

    cfg = sd.SyllableDetectorConfig(SAMPLE_TXT).validate()
    nch = args.channels
    n = int(round(args.hours * 3600 * FS))
    n -= n % 4
    E = cfg.num_evals(n)
    audio_seconds = nch * n / FS                      # per rank per step
    x = synth.make_audio_torch(nch, n, dev, seed=1000 + rank)   # this rank's recording, resident in HBM
    d_out = torch.empty((nch, E, cfg.net_outputs), dtype=torch.float32, device=dev)
    det = sd.BatchDetector(cfg, device=local, kernel=getattr(sd, "KERNEL_" + args.kernel.upper()))
    assert det.active_kernel in (sd.KERNEL_FUSED, sd.KERNEL_TENSOR), "sample.txt must take a fused kernel"
    kernel_name = {sd.KERNEL_FUSED: "fused_detect_kernel<256,4> (SIMT FFT)", sd.KERNEL_TENSOR: ("tc_detect_kernel<4> (tcgen05 3xTF32 band DFT)" if os.environ.get("SYLDET_TC_TF32_CORR") else
                                         "tc_detect_kernel<4,kFast,kF16> (tcgen05 band DFT: TF32 product + one fp16 correction pass)")}[det.active_kernel]
    stream = torch.cuda.current_stream(dev)

    def launch():
        det.launch_device(x.data_ptr(), nch, n, n, detect_rule=sd.DETECT_ANY_OUTPUT, d_outputs_ptr=d_out.data_ptr(),
                          stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

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
    n_det = det.last_detection_count()
    events = det.collect(debounce_frames=0)

    # ---- end to end through the host API: pinned host PCM -> events ------------------------------------------------------
    h = torch.empty((nch, n), dtype=torch.float32, pin_memory=True)
    h.copy_(x)
    torch.cuda.synchronize(dev)
    h_np = h.numpy()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(args.warmup, 2)):
        ev_h = det.run(h_np)
    barrier()
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        ev_h = det.run(h_np)
    barrier()
    te1 = time.perf_counter()
    e2e_wall = te1 - te0
    t_clock1 = time.perf_counter()
    d2h_bytes = len(ev_h) * (16 + 4 * cfg.net_outputs) + 8
    assert len(ev_h) == len(events) and np.array_equal(ev_h.sample, events.sample)
    del h, h_np

    # ---- the same call with 16-bit PCM (what a WAV corpus holds; converted on the device as x/32768): half the PCIe bytes ----
    e2e16 = None
    if not args.no_pcm16:
        q = torch.clamp(torch.round(x * 32768.0), -32768, 32767).to(torch.int16)
        h16 = torch.empty((nch, n), dtype=torch.int16, pin_memory=True)
        h16.copy_(q)
        xq = q.to(torch.float32) / 32768.0             # what the device sees after the conversion
        del q
        det.launch_device(xq.data_ptr(), nch, n, n, detect_rule=sd.DETECT_ANY_OUTPUT, d_outputs_ptr=None, stream=stream.cuda_stream)
        ev_q = det.collect(debounce_frames=0)
        del xq
        torch.cuda.synchronize(dev)
        h16_np = h16.numpy()
        ev16 = det.run(h16_np)
        assert len(ev16) == len(ev_q) and np.array_equal(ev16.sample, ev_q.sample)
        barrier()
        tq0 = time.perf_counter()
        for _ in range(e2e_steps):
            ev16 = det.run(h16_np)
        barrier()
        e2e16 = (time.perf_counter() - tq0, len(ev16) * (16 + 4 * cfg.net_outputs) + 8)
        del h16, h16_np

    # ---- parity spot check against the oracle on slices of this rank's recording (outside the timed regions) -----------
    parity = None
    if rank == 0:
        import oracle
        orc = oracle.Oracle(SAMPLE_TXT)
        rng = np.random.default_rng(0)
        worst, flips = 0.0, 0
        for _ in range(4):
            ch, j = int(rng.integers(nch)), int(rng.integers(max(1, E - 400)))
            seg = x[ch, j * 132: j * 132 + 1444 + 132 * 399].cpu().numpy()
            ref, da, _ = orc.run(seg)
            got = d_out[ch, j:j + ref.shape[0]].cpu().numpy()
            worst = max(worst, float(np.abs(got - ref).max()))
            flips += int(((got[:, 0].astype(np.float64) >= cfg.thresholds[0]) != da).sum())
        parity = {"max_abs_err_vs_oracle": worst, "decision_flips": flips, "evaluations_checked": 1600}

    # ---- reduce over ranks: max time, summed units ------------------------------------------------------------------------
    t = torch.tensor([wall, e2e_wall, dev_ms, e2e16[0] if e2e16 else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, e2e_wall, dev_ms, e2e16_wall = [float(v) for v in t.tolist()]
    total_audio = audio_seconds * world
    value = total_audio * args.steps / wall
    e2e_value = total_audio * e2e_steps / e2e_wall

    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = sum(kernel_ms) / len(kernel_ms)
        alg_bytes = (4 * cfg.hop + 4 * cfg.net_outputs) * E * nch          # per launch, this rank
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        traffic = dram_traffic_per_launch(det.active_kernel == sd.KERNEL_TENSOR, alg_bytes)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (Gaussian noise + synthetic syllables, generated on device; sample.txt weights)",
            "config": workload_config(args, world),
            "frames_per_s": E * nch * world * args.steps / wall,
            "device_ms_per_step": dev_ms / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nch * n * 4, "d2h_bytes_per_step": d2h_bytes,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_wall / e2e_steps,
                    "api": "syldet_batch_run_host (pinned host float32 PCM in, debounced events out; time-sliced copy/detect/collect pipeline)"},
            "gpu_launches": int(gpu_launches),
            "kernel": kernel_name,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["bytes_per_launch"] if traffic else None, "traffic_source": traffic,
                         "peak_source": peak_src, "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "532 B per evaluation (4*hop audio read once + 4*outputs written); FP32 work is 9572 FLOP per "
                                 "evaluation => %.2f TFLOP/s achieved" % (9572.0 * E * nch / (k_ms * 1e-3) / 1e12)},
            "detections_per_step": int(n_det), "events_per_step": len(events), "parity": parity,
            "clocks": clocks.summary(t_clock0, t_clock1),
        }
        if e2e16:
            line["e2e_pcm16"] = {"value": total_audio * e2e_steps / e2e16_wall, "unit": UNIT, "h2d_bytes_per_step": nch * n * 2,
                                 "d2h_bytes_per_step": e2e16[1], "steps": e2e_steps, "ms_per_step": 1e3 * e2e16_wall / e2e_steps,
                                 "api": "syldet_batch_run_host with SYLDET_PCM_S16 (the same recording quantised to 16-bit PCM)"}
        if world == 1 and not args.no_cpu:
            cb = cpu_reference(args, steps=1, warmup=0)
            line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port", "sample": cb["sample"],
                                    "frames_per_s": cb["frames_per_s"]}
        if world == 1 and not args.no_stream:
            line["stream"] = stream_latency(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cb = cpu_reference(args, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (Gaussian noise + synthetic syllables; sample.txt weights)",
            "config": workload_config(args, world), "frames_per_s": cb["frames_per_s"],
            "cpu_baseline": {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port", "sample": cb["sample"]},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU port (oracle/oracle.c) of the reference's Swift/Accelerate path, which cannot be built on Linux"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hours", type=float, default=1.0)
    ap.add_argument("--channels", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-step-seconds", type=float, default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pcm16", action="store_true")
    ap.add_argument("--kernel", default="auto", choices=["auto", "fused", "tensor"])
    ap.add_argument("--no-stream", action="store_true")
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
