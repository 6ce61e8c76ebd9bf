#!/usr/bin/env python
"""bench.py - headline benchmark of the compression round trip (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): Mpixel/s of the full round trip - RGB->YCbCr, 4:2:0 decimation,
8x8 DCT, quantise at Q, dequantise, IDCT, upsample, YCbCr->RGB, PSNR / SSIM / bits-per-
pixel reductions - on 4K (3840x2160) frames, and the fraction of the HBM roofline.

A "step" is one pass of the hot path over one batch of FRAMES distinct synthetic 4K
frames per GPU (fast fp32 mode, Q=50, 4:2:0, full SSIM on R,G,B,Y).  The batch
(16 frames = 398 MB in + 398 MB out) is larger than the 126 MB L2, so every step
streams from HBM.  Weak scaling: each rank owns its own batch (frames shard with no
data-path collective); the per-step metric partials are all-reduced over NCCL.

JSON line (rank 0): value = whole-job Mpixel/s with inputs resident in HBM;
e2e = the same through the public API with pinned HOST buffers (H2D of the frames
and D2H of the reconstructed frames + metrics inside the timed region);
roofline = dominant kernel vs measured HBM peak; cpu_baseline = the oracle port timed
on this box's host cores on a bounded sample; exact = the fp64 bit-exact mode on the
same workload.

--impl reference: the reference's own CPU implementation on the host cores - the unmodified
engines/pipeline.py::compress_reconstruct staged under oracle/_ref by oracle/make_ref.py
(kind "reference"; the bit-exact NumPy port oracle/numpy_port.py beside it, and alone with kind
"port" when nothing is staged), one process per host core, same metric/config.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 2160, 3840
QUALITY, MODE, PREFILTER = 50, "4:2:0", False
FRAMES = int(os.environ.get("JDS_BENCH_FRAMES", "16"))
BYTES_PER_PX = 6.0          # algorithmic: 3 B read + 3 B written per pixel
WORKLOAD = (f"{FRAMES}x 4K (3840x2160) random RGB frames per GPU per step, Q={QUALITY} {MODE} "
            f"prefilter off, round trip + PSNR/SSIM(R,G,B,Y)/bpp, fast fp32 mode")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML, 10 ms period - every
    NVML query takes driver locks that delay kernel launches, measurably so on an 8-GPU box;
    falls back to `nvidia-smi -lms` when pynvml is unusable)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []
        self.samples, self.reasons, self.power, self.max_mhz = [], set(), [], None
        self._stop = threading.Event()
        self.nvml = None

    def _nvml_loop(self):
        n = self.nvml
        h = self.handle
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop.is_set():
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                try:
                    self.power.append(n.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                if get_reasons:
                    r = get_reasons(h)
                    for name, bit in bits.items():
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.010)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # physical index of this process's CUDA device (CUDA_VISIBLE_DEVICES aware)
            idx = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    pass
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)      # warm the call path
            self.nvml = pynvml
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=1)
            sm = self.samples
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                    "power_w_max": max(self.power) if self.power else None,
                    "samples": len(sm), "reasons": sorted(self.reasons), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


class StdoutToStderr:
    """While active, file descriptor 1 points at stderr: NCCL writes its version / debug lines
    to stdout from C, and rank 0 must print exactly ONE JSON line there."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def bind_to_gpu_numa_node(local):
    """Pin this rank's threads to the CPU cores next to its GPU (NVML's ideal-CPU mask) BEFORE any
    pinned host buffer exists: page-locked memory is first-touched by this process, so it then
    lives on the NUMA node of the GPU's PCIe root and eight ranks no longer pull their frames
    through one socket's memory controllers (e2e efficiency 0.20 at 8 GPUs).
    Returns a short description for the JSON line; JDS_BENCH_NUMA=0 switches it off."""
    if os.environ.get("JDS_BENCH_NUMA", "1") == "0":
        return "off (JDS_BENCH_NUMA=0)"
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = local
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                idx = int(vis.split(",")[local])
            except (ValueError, IndexError):
                pass
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return "no usable CPUs in the GPU's affinity mask"
        os.sched_setaffinity(0, allowed)
        node = None
        for n in range(16):                       # which NUMA node holds the first of these cores
            try:
                txt = open(f"/sys/devices/system/node/node{n}/cpulist").read().strip()
            except OSError:
                break
            for part in txt.split(","):
                lo, _, hi = part.partition("-")
                if int(lo) <= allowed[0] <= int(hi or lo):
                    node = n
        return f"bound to {len(allowed)} CPUs near GPU {idx} (NUMA node {node})"
    except Exception as e:                        # noqa: BLE001 - a diagnostic, never fatal
        return f"not bound ({type(e).__name__}: {e})"


def make_frames(rank, n):
    import numpy as np
    out = np.empty((n, H, W, 3), dtype=np.uint8)
    for k in range(n):
        out[k] = np.random.default_rng(4000 + 100 * rank + k).integers(0, 256, (H, W, 3), dtype=np.uint8)
    return out


# ------------------------------------------------------------------------------------
# CPU arms (oracle port = the checker; here only as the timed CPU baseline)
# ------------------------------------------------------------------------------------
def _frame(seed):
    import numpy as np
    return np.random.default_rng(seed).integers(0, 256, (H, W, 3), dtype=np.uint8)


def _cpu_one_frame(seed):
    """one 4K frame through the oracle port (oracle/numpy_port.py)"""
    from oracle import numpy_port as P
    img = _frame(seed)
    t0 = time.perf_counter()
    o = P.compress_reconstruct(img, QUALITY, MODE, PREFILTER, want_maps=False)
    return time.perf_counter() - t0, o["psnr_y"]


def reference_staged():
    """True when the unmodified reference is importable here: /root/reference in the build
    container, oracle/_ref/reference (staged by oracle/make_ref.py) on the GPU box."""
    try:
        from oracle import reference_shim as R
        return R.available()
    except Exception:
        return False


def _ref_one_frame(seed):
    """one 4K frame through the UNMODIFIED reference's engines/pipeline.py::compress_reconstruct
    (imported by oracle/reference_shim.py with the skimage stand-in)"""
    from oracle import reference_shim as R
    ref = R.load()
    img = _frame(seed)
    params = ref.CompressionParams(quality=QUALITY, subsampling_mode=MODE, use_prefilter=PREFILTER)
    t0 = time.perf_counter()
    res, _ = ref.compress_reconstruct(img, params)
    return time.perf_counter() - t0, res.psnr_y


def cpu_baseline_sample():
    """The CPU figure printed beside the GPU one (rank 0, N=1): the reference AS IS on one 4K
    frame of the workload in one process (about 20-30 s; SURVEY 8d "baseline of record"), with
    the bit-exact NumPy port on two frames beside it.  Without a staged reference the port is
    the baseline (kind "port")."""
    t_port = [_cpu_one_frame(4000 + k)[0] for k in range(2)]
    port = {"value": round(2 * H * W / sum(t_port) / 1e6, 4), "unit": "Mpixel/s", "cores": 1,
            "sample": f"2 of the workload's 4K frames through oracle/numpy_port.py, 1 process, {sum(t_port):.1f} s"}
    if reference_staged():
        sec, _ = _ref_one_frame(4000)
        return {"value": round(H * W / sec / 1e6, 4), "unit": "Mpixel/s", "cores": 1, "kind": "reference",
                "sample": f"1 of the workload's 4K frames through the unmodified reference "
                          f"(engines/pipeline.py::compress_reconstruct, skimage stand-in), 1 Python "
                          f"process (OpenCV's own threads as the reference leaves them), {sec:.1f} s",
                "port": port}
    port["kind"] = "port"
    return port


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on the host cores, one process
    per core, each step = one 4K frame per process (a bounded sample of the workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    ncpu = os.cpu_count() or 1
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        pass
    try:
        avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2**30
    except Exception:
        avail_gb = 64.0
    procs = max(1, min(ncpu, 32, int(avail_gb // 6)))     # ~5 GB peak per 4K frame in fp64
    as_is = reference_staged()
    worker = _ref_one_frame if as_is else _cpu_one_frame
    # bounded: the reference needs 20-30 s per 4K frame (the port 5 s), so cap the steps
    steps = min(max(1, args.steps), 2 if as_is else 3)
    warmup = min(max(0, args.warmup), 0 if as_is else 1)
    ctx = mp.get_context("spawn")
    os.environ.setdefault("OMP_NUM_THREADS", "1")         # one core per process: cv2 / BLAS threads off
    os.environ.setdefault("OPENCV_FOR_THREADS_NUM", "1")
    with ctx.Pool(procs) as pool:
        pool.map(_frame, [1] * procs)                      # start the workers (imports) untimed
        for _ in range(warmup):
            pool.map(worker, [4000 + k for k in range(procs)])
        t0 = time.perf_counter()
        for s in range(steps):
            pool.map(worker, [4000 + k for k in range(procs)])
        sec = time.perf_counter() - t0
        port_line = None
        if as_is:                                          # the port beside it, one step
            t1 = time.perf_counter()
            pool.map(_cpu_one_frame, [4000 + k for k in range(procs)])
            sp = time.perf_counter() - t1
            port_line = {"value": round(procs * H * W / sp / 1e6, 4), "unit": "Mpixel/s", "cores": procs,
                         "sample": f"1 step x {procs} 4K frames, oracle/numpy_port.py, {sp:.1f} s"}
    mpx = steps * procs * H * W / sec / 1e6
    kind = "reference" if as_is else "port"
    what = ("the unmodified reference (engines/pipeline.py::compress_reconstruct, skimage stand-in)"
            if as_is else "oracle/numpy_port.py (bit-exact restatement; no staged reference here)")
    line = {
        "impl": "reference", "metric": "4K round-trip Mpixel/s (incl. PSNR/SSIM/bpp)",
        "value": round(mpx, 4), "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": round(sec / steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD.replace("fast fp32 mode", "reference fp64 arithmetic on host CPU"),
                   "step": f"{procs} frames, one per process"},
        "cpu_baseline": {"value": round(mpx, 4), "unit": "Mpixel/s", "cores": procs, "kind": kind,
                         "sample": f"{steps} steps x {procs} 4K frames, {what} in "
                                   f"{procs} processes ({ncpu} host cores visible), {sec:.1f} s",
                         "port": port_line},
        "e2e": {"value": round(mpx, 4), "unit": "Mpixel/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    numa = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # one launch per step: let the whole batch fit the scratch budget of the context
    os.environ.setdefault("JDS_SCRATCH_MB", "8192")
    import jpeg_dsp_studio_b200 as J
    eng = J.Engine(local)
    stream = torch.cuda.current_stream(dev)
    eng.use_stream(stream.cuda_stream)

    frames_np = make_frames(rank, FRAMES)
    host_in = torch.from_numpy(frames_np).pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    d_in = host_in.to(dev, non_blocking=False)
    d_out = torch.empty_like(d_in)
    px_per_step = FRAMES * H * W
    # streamed steps alternate between two contexts on two streams: the ragged last wave of one
    # step's kernels is filled by the next step's (each context has its own scratch and output)
    N_STREAMS = max(1, min(4, int(os.environ.get("JDS_BENCH_STREAMS", "3"))))
    engines, streams, d_outs = [eng], [stream], [d_out]
    for _ in range(N_STREAMS - 1):
        s_k = torch.cuda.Stream(dev)
        eng_k = J.Engine(local)
        eng_k.use_stream(s_k.cuda_stream)
        engines.append(eng_k)
        streams.append(s_k)
        d_outs.append(torch.empty_like(d_in))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # The path's only exchange is the all-reduce of the metric partials (NCCL).  Units are
    # independent, so a job needs ONE all-reduce: "job" (default) accumulates the records of
    # every step on the device and reduces them once, inside the timed region, after the last
    # step (again whenever the ring of RING steps is full).  "step" issues one asynchronous
    # all-reduce per step instead (diagnosis: an NCCL kernel resident next to the SSIM kernel
    # costs ~9 % of the step, measured at 2 and 8 GPUs); "off" never reduces.
    REDUCE_MODE = os.environ.get("JDS_BENCH_REDUCE", "job")      # job | step | off
    RING = 64
    from jpeg_dsp_studio_b200 import _native as NAT
    rec_all = torch.zeros((RING, FRAMES, NAT.JDS_RECORD_FIELDS), dtype=torch.float64, device=dev)
    host_acc = torch.zeros(8, dtype=torch.float64).pin_memory()     # per-call API: host partials
    dev_acc = torch.zeros(8, dtype=torch.float64, device=dev)
    pending = []
    state = {"stream_steps": 0, "host_steps": 0, "reduces": 0}

    def reduce_job():
        """one all-reduce of everything accumulated since the last one"""
        n, m = state["stream_steps"], state["host_steps"]
        state["stream_steps"] = state["host_steps"] = 0
        if world == 1 or REDUCE_MODE == "off":
            return
        if n:
            join_streams()
            dist.all_reduce(rec_all[:n])
            state["reduces"] += 1
        if m:
            dev_acc.copy_(host_acc, non_blocking=True)
            dist.all_reduce(dev_acc)
            state["reduces"] += 1

    def join_streams():
        """the main stream waits for the second one (before a collective / the closing event)"""
        for s_ in streams[1:]:
            stream.wait_stream(s_)

    def fork_streams():
        for s_ in streams[1:]:
            s_.wait_stream(stream)

    def drain():
        pipe_drain()
        while pending:
            pending.pop(0).wait()
        join_streams()
        reduce_job()
        host_acc.zero_()

    def reduce_partials(outs):
        """per-call API: the step's host-side metric structs are added to the job's partials"""
        ms = [o.metrics for o in outs]
        sse = float(sum(m.sse_rgb for m in ms))
        ssey = float(sum(m.sse_y for m in ms))
        bits = float(sum(2 * m.luma_blocks + m.coeff_bits for m in ms))
        ssim = [float(sum(m.ssim_sum[c] for m in ms)) for c in range(4)]
        host_acc.numpy()[:] += [sse, ssey, bits] + ssim + [float(len(outs))]
        state["host_steps"] += 1
        if REDUCE_MODE == "step":
            reduce_job()
            host_acc.zero_()

    def step_device(precision):
        outs = eng.roundtrip_batch(d_in, QUALITY, MODE, PREFILTER, precision=precision,
                                   recon_out=d_out)
        reduce_partials(outs)
        return outs

    def step_stream(precision):
        """the production loop for device-resident frames: the kernels and the metric records of
        a step are enqueued without any host synchronisation (jds_roundtrip_batch_records);
        the records of all steps wait on the device for the job's one all-reduce, and the host
        reads metrics only when it needs them - here once, after the timed region"""
        if state["stream_steps"] == RING:
            reduce_job()
        i = state["stream_steps"]
        state["stream_steps"] += 1
        k = i % state.get("stream_engines", len(engines))
        engines[k].batch_records(d_in, rec_all[i], QUALITY, MODE, PREFILTER, precision=precision,
                                 recon_out=d_outs[k], unit0=rank, unit_step=world)
        if REDUCE_MODE == "step" and world > 1:
            state["stream_steps"] = 0
            pending.append(dist.all_reduce(rec_all[i], async_op=True))
            while len(pending) >= 16:
                pending.pop(0).wait()
        return rec_all[i]

    def step_host(precision):
        outs = eng.roundtrip_batch(host_in, QUALITY, MODE, PREFILTER, precision=precision,
                                   recon_out=host_out)
        reduce_partials(outs)
        return outs

    # consecutive host batches overlapped: two contexts alternate, batch k+1 is enqueued
    # (jds_roundtrip_batch_begin) before batch k is awaited (jds_ctx_finish), so the fill of one
    # batch's PCIe pipeline hides behind the drain of the other's - compress_stream's loop.
    # Every step still uploads its 16 frames from pinned host memory and downloads its 16
    # reconstructions + metric structs; all steps have finished before the closing event.
    from jpeg_dsp_studio_b200.engine import stream_engines
    pipe_engines = stream_engines(local, 2)
    host_out2 = torch.empty_like(host_in).pin_memory()
    pipe = {"pending": None, "k": 0}

    def step_host_pipelined(precision):
        k = pipe["k"]
        pipe["k"] = k + 1
        nxt = pipe_engines[k & 1].roundtrip_batch_begin(host_in, QUALITY, MODE, PREFILTER, precision=precision,
                                                        recon_out=(host_out, host_out2)[k & 1])
        prev, pipe["pending"] = pipe["pending"], nxt
        if prev is not None:
            reduce_partials(prev.result())
        return nxt

    def pipe_drain():
        if pipe["pending"] is not None:
            reduce_partials(pipe["pending"].result())
            pipe["pending"] = None

    def step_host_metrics(precision):
        # what the batch / sweep consumers read (gui/worker.py:66-72: scalars only): frames go up,
        # only the metric structs come back - compress_batch(keep_images=False)
        outs = eng.roundtrip_batch(host_in, QUALITY, MODE, PREFILTER, precision=precision, want_recon=False)
        reduce_partials(outs)
        return outs

    def timed(fn, precision, steps, warmup, sample_clocks=False, stage_timing=False):
        eng.set_stage_timing(stage_timing)
        for _ in range(warmup):
            fn(precision)
        drain()
        if world > 1 and REDUCE_MODE == "job" and fn is step_stream:
            # NCCL sets up protocols / connections lazily per message size: take the job's
            # all-reduce sizes once outside the timed region (the rows are overwritten by the steps)
            for n in {min(steps, RING), steps % RING}:
                if n:
                    dist.all_reduce(rec_all[:n])
            rec_all.zero_()
        # the sampler starts BEFORE the barrier: NVML start-up takes tens of ms, and a rank that
        # enters the timed region late makes every other rank wait in the job's all-reduce
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None   # rank 0's GPU; it prints the line
        if sampler:
            sampler.start()
        barrier()
        eng.stage_times(reset=True) if stage_timing else None
        l0 = sum(e.launch_count() for e in engines)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fork_streams()
        for _ in range(steps):
            outs = fn(precision)
        drain()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if os.environ.get("JDS_BENCH_DEBUG"):
            print(f"[debug] rank {rank} {fn.__name__} {precision}: {ms / steps:.4f} ms/step", file=sys.stderr, flush=True)
        clocks = sampler.stop() if sampler else None
        stages = eng.stage_times(reset=True) if stage_timing else None
        launches = sum(e.launch_count() for e in engines) - l0
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), stages, launches, clocks, outs

    K, Wm = args.steps, max(args.warmup, 3)
    # headline: production configuration (no per-kernel events)
    ms, _, launches, clocks, last_rec = timed(step_stream, "fast", K, Wm, sample_clocks=True)
    value = world * px_per_step * K / (ms / 1e3) / 1e6
    rec_host = last_rec.cpu().numpy()               # the metrics of the last step, read once
    if os.environ.get("JDS_BENCH_ONLY_HEADLINE"):
        print(f"[headline] streams {N_STREAMS}: {ms / K:.4f} ms/step {value:.1f} Mpixel/s", file=sys.stderr, flush=True)
        return None
    # the same batch through the synchronising per-call API (returns host structs every step)
    ms_sync, _, _, _, outs = timed(step_device, "fast", K, Wm)
    sync_value = world * px_per_step * K / (ms_sync / 1e3) / 1e6
    # the streamed records must say what the per-call API says (summed over ranks by the all-reduce)
    if world == 1:
        import numpy as _np
        assert _np.array_equal(rec_host[:, 2], _np.array([float(o.metrics.sse_rgb) for o in outs]))
        assert _np.allclose(rec_host[:, 7], [o.metrics.ssim_sum[3] for o in outs], rtol=1e-12)
    # per-kernel times for the roofline: same steps with CUDA events around every kernel
    Kp = max(1, min(K, 10))
    _, stages, _, _, _ = timed(step_device, "fast", Kp, 2, stage_timing=True)

    ms_e2e, _, _, _, _ = timed(step_host, "fast", K, Wm)
    e2e = world * px_per_step * K / (ms_e2e / 1e3) / 1e6
    ms_e2e_p, _, _, _, _ = timed(step_host_pipelined, "fast", K, Wm)
    e2e_pipe = world * px_per_step * K / (ms_e2e_p / 1e3) / 1e6
    ms_e2e_m, _, _, _, _ = timed(step_host_metrics, "fast", K, Wm)
    e2e_metrics = world * px_per_step * K / (ms_e2e_m / 1e3) / 1e6

    Kx = max(1, min(K, 3))
    ms_x, stages_x, _, _, outs_x = timed(step_device, "exact", Kx, 1, stage_timing=True)
    eng.set_stage_timing(False)
    exact_sync_value = world * px_per_step * Kx / (ms_x / 1e3) / 1e6
    # ... and streamed like the headline, but on ONE context: no host synchronisation between the
    # steps (the per-call API leaves the GPU idle while Python unpacks 16 metric structs); the fp64
    # kernels of concurrent steps would only compete for the same FP64 pipe (three contexts: slower)
    Kxs = max(3, min(K, 10))
    state["stream_engines"] = 1
    ms_xs, _, _, _, _ = timed(step_stream, "exact", Kxs, 2)
    del state["stream_engines"]
    exact_value = world * px_per_step * Kxs / (ms_xs / 1e3) / 1e6
    # the bit-exact mode through the same host->host call (H2D + D2H inside the timed region)
    ms_xe, _, _, _, _ = timed(step_host, "exact", Kx, 1)
    exact_e2e = world * px_per_step * Kx / (ms_xe / 1e3) / 1e6

    # north_star: the fp32 mode "must report its round-half mismatch rate" - frame 0 of this
    # rank's batch, fast against exact: quantised coefficients (round-half decisions) and pixels
    fa = eng.roundtrip(d_in[0], QUALITY, MODE, PREFILTER, precision="fast", want_coeffs=True)
    xa = eng.roundtrip(d_in[0], QUALITY, MODE, PREFILTER, precision="exact", want_coeffs=True)
    mismatch = {"coeff": float((fa.coeffs != xa.coeffs).double().mean().item()),
                "pixel": float((fa.recon != xa.recon).double().mean().item()),
                "d_psnr_y_db": abs(fa.scalars["psnr_y"] - xa.scalars["psnr_y"]),
                "d_ssim_y": abs(fa.scalars["ssim_y"] - xa.scalars["ssim_y"]),
                "d_ssim_rgb": abs(fa.scalars["ssim_rgb"] - xa.scalars["ssim_rgb"]),
                "frame": "frame 0 of rank 0's batch, fast vs exact mode (exact = bit-identical to the reference)"}
    del fa, xa

    # BASELINE config 4 inside the same launch: the 100-point sweep, points sharded over the
    # ranks (strong scaling) - SCALE_rNN.json then carries north_star's ">= 7x on 8 GPUs"
    sweep_rec = measure_sweep(eng, dev, stream, world, rank, max(3, min(K, 10)), 3)

    # uniform-random RGB is the worst case for coefficient density; also report a
    # natural-statistics input (generate_photo(512) tiled to 4K, each frame shifted) at N=1
    natural = None
    if world == 1:
        from jpeg_dsp_studio_b200.utils import test_images as TI
        tile = np.tile(TI.generate_photo(512), (-(-H // 512), -(-W // 512), 1))[:H, :W]
        nat_np = np.stack([np.roll(tile, (11 * k, 37 * k), axis=(0, 1)) for k in range(FRAMES)])
        d_nat = torch.from_numpy(np.ascontiguousarray(nat_np)).to(dev)

        def step_nat(precision):
            outs = eng.roundtrip_batch(d_nat, QUALITY, MODE, PREFILTER, precision=precision,
                                       recon_out=d_out)
            reduce_partials(outs)
            return outs
        Kn = max(1, min(K, 10))
        ms_n, _, _, _, outs_n = timed(step_nat, "fast", Kn, 2)
        sn = outs_n[0].scalars
        natural = {"value": round(px_per_step * Kn / (ms_n / 1e3) / 1e6, 2), "unit": "Mpixel/s",
                   "ms_per_step": round(ms_n / Kn, 4), "steps": Kn,
                   "input": f"generate_photo(512) tiled to {W}x{H}, {FRAMES} shifted copies",
                   "nonzero_fraction": round(sn["nonzero_count"] / sn["total_coeffs"], 4),
                   "bpp": sn["bpp"], "psnr_y": sn["psnr_y"], "ssim_y": sn["ssim_y"]}
        del d_nat

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    peak, peak_src = measured_peak()
    dom = max(stages, key=lambda k: stages[k]["ms"])
    launches_per_step = max(stages[dom]["launches"], 1) / Kp
    dom_ms = stages[dom]["ms"] / max(stages[dom]["launches"], 1)
    kern_ms_step = sum(v["ms"] for v in stages.values()) / Kp
    alg_bytes_step = BYTES_PER_PX * px_per_step
    alg_bytes = alg_bytes_step / launches_per_step          # per launch of the dominant kernel
    achieved = alg_bytes / (dom_ms / 1e3) / 1e9
    path_gbs = alg_bytes_step / (kern_ms_step / 1e3) / 1e9
    kern_ms_step_x = sum(v["ms"] for v in stages_x.values()) / Kx
    # DRAM traffic of the dominant kernel from the committed ncu capture (per pixel, scaled
    # to the pixels of one launch here); null when no capture exists for that kernel
    traffic, traffic_src = None, None
    kmap = {"ssim": "k_ssim_strip", "block_codec": "k_fast_luma", "forward_colour": "k_fast_chroma"}
    path_traffic = None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            per_px = tj[kmap[dom]]["dram_bytes_per_pixel"]
            traffic = round(per_px * px_per_step / launches_per_step, 0)
            traffic_src = tj["source"]
            wp = tj.get("whole_path")
            if wp:
                path_traffic = {"bytes_per_step": round(wp["dram_bytes_per_pixel"] * px_per_step, 0),
                                "dram_bytes_per_pixel": wp["dram_bytes_per_pixel"],
                                "algorithmic_bytes_per_pixel": BYTES_PER_PX, "source": wp.get("source")}
            break
        except Exception:
            continue
    # the CPU baseline is timed on rank 0 at N=1 only (it takes ~15 s of host time)
    cpu = cpu_baseline_sample() if world == 1 else None
    o0 = outs[0]
    line = {
        "metric": "4K round-trip Mpixel/s (incl. PSNR/SSIM/bpp)",
        "value": round(value, 2), "unit": "Mpixel/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": FRAMES,
                   "l2_policy": f"batch of {FRAMES} frames ({FRAMES * H * W * 3 / 1e6:.0f} MB in, "
                                "same out) exceeds the 126 MB L2; no flush needed",
                   "sharding": "frames across ranks, no data-path collective; the job's metric "
                               "records are all-reduced (NCCL) once, inside the timed region, after "
                               "the last step (JDS_BENCH_REDUCE=step: once per step)"},
        "e2e": {"value": round(e2e, 2), "unit": "Mpixel/s",
                "h2d_bytes_per_step": FRAMES * H * W * 3,
                "d2h_bytes_per_step": FRAMES * H * W * 3 + FRAMES * 504,
                "ms_per_step": round(ms_e2e / K, 4),
                "pcie_gbs_per_rank_each_way": round(FRAMES * H * W * 3 / (ms_e2e / K * 1e-3) / 1e9, 2),
                "host_numa": numa,
                "api": "Engine.roundtrip_batch(pinned host uint8 frames) -> jds_roundtrip_batch (C ABI)",
                "pipelined": {"value": round(e2e_pipe, 2), "unit": "Mpixel/s",
                              "ms_per_step": round(ms_e2e_p / K, 4),
                              "h2d_bytes_per_step": FRAMES * H * W * 3,
                              "d2h_bytes_per_step": FRAMES * H * W * 3 + FRAMES * 504,
                              "pcie_gbs_per_rank_each_way": round(FRAMES * H * W * 3 / (ms_e2e_p / K * 1e-3) / 1e9, 2),
                              "api": "compress_stream's loop: Engine.roundtrip_batch_begin on two alternating "
                                     "contexts, PendingBatch.result() one step later (jds_roundtrip_batch_begin / "
                                     "jds_ctx_finish): same bytes up and down every step, consecutive steps overlap"},
                "metrics_only": {"value": round(e2e_metrics, 2), "unit": "Mpixel/s",
                                 "ms_per_step": round(ms_e2e_m / K, 4),
                                 "h2d_bytes_per_step": FRAMES * H * W * 3,
                                 "d2h_bytes_per_step": FRAMES * 504,
                                 "pcie_gbs_per_rank_h2d": round(FRAMES * H * W * 3 / (ms_e2e_m / K * 1e-3) / 1e9, 2),
                                 "api": "the same call with want_recon=False (compress_batch(keep_images=False)): "
                                        "what the reference's batch / sweep consumers read, gui/worker.py:66-72"}},
        "gpu_launches": launches,
        "value_mode": "streamed: per-step kernels and device-resident metric records enqueued without "
                      "host synchronisation (jds_roundtrip_batch_records), steps alternating over "
                      f"{N_STREAMS} context(s) / stream(s); one all-reduce of the job's records after the "
                      "last step; metrics read after the timed region",
        "sync_api": {"value": round(sync_value, 2), "unit": "Mpixel/s", "ms_per_step": round(ms_sync / K, 4),
                     "api": "Engine.roundtrip_batch(device tensors): synchronises and returns host "
                            "metric structs every step"},
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": dom, "kernel_ms_per_launch": round(dom_ms, 4),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "launches_per_step": launches_per_step,
                     "whole_path": {"kernel_ms_per_step": round(kern_ms_step, 4),
                                    "achieved": round(path_gbs, 1),
                                    "frac": round(path_gbs / peak, 4),
                                    "traffic": path_traffic,
                                    "stages_ms_per_step": {k: round(v["ms"] / Kp, 4) for k, v in stages.items()}}},
        "cpu_baseline": cpu,
        "natural_statistics": natural,
        "fast_mode_mismatch": mismatch,
        "sweep": sweep_rec,
        "exact_mode": {"value": round(exact_value, 2), "unit": "Mpixel/s", "dtype": "f64",
                       "ms_per_step": round(ms_xs / Kxs, 4), "steps": Kxs,
                       "value_mode": "streamed like the headline value (jds_roundtrip_batch_records, no host "
                                     "synchronisation per step) on ONE context / stream",
                       "sync_api": {"value": round(exact_sync_value, 2), "unit": "Mpixel/s",
                                    "ms_per_step": round(ms_x / Kx, 4), "steps": Kx,
                                    "note": "per-call API (synchronises and returns host metric structs every "
                                            "step) with CUDA events around every kernel - the run the stage "
                                            "times below come from"},
                       "e2e": {"value": round(exact_e2e, 2), "unit": "Mpixel/s",
                               "ms_per_step": round(ms_xe / Kx, 4),
                               "h2d_bytes_per_step": FRAMES * H * W * 3,
                               "d2h_bytes_per_step": FRAMES * H * W * 3 + FRAMES * 504,
                               "api": "Engine.roundtrip_batch(pinned host frames, precision='exact')"},
                       "hbm_frac_whole_path": round(alg_bytes_step / (kern_ms_step_x / 1e3) / 1e9 / peak, 4),
                       "stages_ms_per_step": {k: round(v["ms"] / Kx, 4) for k, v in stages_x.items()}},
        "results_frame0": {"psnr_y": o0.scalars["psnr_y"], "ssim_y": o0.scalars["ssim_y"],
                           "bpp": o0.scalars["bpp"],
                           "psnr_y_exact_mode": outs_x[0].scalars["psnr_y"],
                           "ssim_y_exact_mode": outs_x[0].scalars["ssim_y"]},
    }
    if world > 1:
        dist.destroy_process_group()
    return line


def measure_sweep(eng, dev, stream, world, rank, K, Wm):
    """BASELINE.json config 4: a 100-point quality sweep (Q=1..100) of one 4K frame, 4:2:0,
    fast mode, metrics only, points dealt round-robin to the ranks (STRONG scaling), one
    all_gather of the device-resident records per sweep.  Three figures, each the max over ranks:
      single_sweep    K sweeps strictly one after the other, CUDA events around the whole loop
                      (host finalisation of every table inside): the latency a GUI sweep sees
      single_device   the same sweeps, but each one's own CUDA-event span (first kernel ->
                      records gathered and copied to the host) summed: the device time of a sweep
      pipelined       consecutive sweeps overlapped (a folder of frames)
    Returns the record on rank 0, None elsewhere."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from jpeg_dsp_studio_b200 import distributed as D
    img = np.random.default_rng(4).integers(0, 256, (H, W, 3), dtype=np.uint8)
    d_img = torch.from_numpy(img).to(dev)
    qs = list(range(1, 101))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    dev_spans = []

    def step():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        h = D.sweep_sharded_begin(eng, d_img, qs, MODE, PREFILTER, precision="fast", device=dev)
        b.record(stream)
        table = h.result()
        dev_spans.append((a, b))
        return table

    def timed(run_steps, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = run_steps(K)
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    def one_at_a_time(K):
        for _ in range(K):
            table = step()
        return table

    def pipelined(K):
        # throughput of MANY sweeps (a folder of frames): sweep i+1 is enqueued before the
        # table of sweep i is finalised, so the host work hides behind the kernels; every
        # sweep still ends in its own all_gather and its own full result table
        prev, table = None, None
        for _ in range(K):
            h = D.sweep_sharded_begin(eng, d_img, qs, MODE, PREFILTER, precision="fast", device=dev)
            if prev is not None:
                table = prev.result()
            prev = h
        return prev.result()

    one_at_a_time(Wm)
    pipelined(Wm)
    dev_spans.clear()
    l0 = eng.launch_count()
    ms_lat, table_lat = timed(one_at_a_time, K)
    launches = eng.launch_count() - l0
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in dev_spans)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev = float(t.item())
    ms, table = timed(pipelined, K)
    for a, b in zip(table, table_lat):      # same table either way (SSIM sums: atomic order only)
        assert a["estimated_bits"] == b["estimated_bits"] and a["psnr_rgb"] == b["psnr_rgb"]
        assert abs(a["ssim_y"] - b["ssim_y"]) < 1e-9
    del d_img
    if rank != 0:
        return None
    px = len(qs) * H * W
    return {
        "workload": "100-point quality sweep Q=1..100 of one random 4K frame, 4:2:0, fast fp32 mode, "
                    "metrics only (BASELINE config 4)",
        "scaling": "strong", "n_gpus": world, "points": len(qs), "sweeps_timed": K,
        "sharding": "sweep points round-robin over ranks; one all_gather of the device-resident "
                    "records per sweep",
        "single_sweep": {"ms_per_sweep": round(ms_lat / K, 4),
                         "value": round(px * K / (ms_lat / 1e3) / 1e6, 2), "unit": "Mpixel/s",
                         "note": "one sweep at a time: enqueue, all_gather, D2H, synchronise, finalise "
                                 "the table on the host, then the next (latency of one GUI sweep)"},
        "single_device": {"ms_per_sweep": round(ms_dev / K, 4),
                          "value": round(px * K / (ms_dev / 1e3) / 1e6, 2), "unit": "Mpixel/s",
                          "note": "the same isolated sweeps, each timed by its own CUDA-event pair on the "
                                  "stream (kernels + all_gather + D2H of the records), host work excluded"},
        "pipelined": {"ms_per_sweep": round(ms / K, 4), "value": round(px * K / (ms / 1e3) / 1e6, 2),
                      "unit": "Mpixel/s",
                      "note": "consecutive sweeps pipelined (one in flight while the previous table is finalised)"},
        "gpu_launches": launches,
        "rd_table_sample": {"q10": table[9], "q50": table[49], "q90": table[89]},
    }


def run_sweep(args):
    """--workload sweep: config 4 on its own (see measure_sweep); value = the pipelined figure."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    os.environ.setdefault("JDS_SCRATCH_MB", "8192")
    import jpeg_dsp_studio_b200 as J
    eng = J.Engine(local)
    stream = torch.cuda.current_stream(dev)
    eng.use_stream(stream.cuda_stream)
    K, Wm = args.steps, max(args.warmup, 3)
    rec = measure_sweep(eng, dev, stream, world, rank, K, Wm)
    line = None
    if rank == 0:
        line = {
            "metric": "4K round-trip Mpixel/s (incl. PSNR/SSIM/bpp)", "value": rec["pipelined"]["value"],
            "unit": "Mpixel/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": rec["pipelined"]["ms_per_sweep"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": rec["workload"], "sharding": rec["sharding"],
                       "pipelining": "value: consecutive sweeps pipelined; single_sweep: strictly one at a time"},
            "gpu_launches": rec["gpu_launches"],
            "single_sweep": rec["single_sweep"], "single_device": rec["single_device"],
            "rd_table_sample": rec["rd_table_sample"],
        }
    if world > 1:
        dist.destroy_process_group()
    return line


def run_configs(args):
    """--workload configs: the other BASELINE.json configurations on this GPU / these GPUs, each
    with its throughput, its fraction of the HBM roofline (algorithmic bytes of THAT config,
    optional outputs included) and a CPU figure beside it:
      cfg2  1920x1080 random, Q=50, 4:4:4 (k_fast_444)
      cfg3  3840x2160 random, Q=75, 4:2:0 + prefilter, with coefficients + histogram + error map
      cfg5  1024 x 1080p, Q=30, 4:2:2, frames sharded over the ranks (aggregate + GB/s per GPU)
    Inputs cycle through more distinct frames than fit the 126 MB L2."""
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    os.environ.setdefault("JDS_SCRATCH_MB", "8192")
    import jpeg_dsp_studio_b200 as J
    from jpeg_dsp_studio_b200 import _native as NAT
    eng = J.Engine(local)
    stream = torch.cuda.current_stream(dev)
    eng.use_stream(stream.cuda_stream)
    peak, peak_src = measured_peak()
    K = max(3, min(args.steps, 20))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n, warm=2):
        for i in range(warm):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(n):
            fn(i)
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n

    def rec(ms, px, bytes_px, note, **extra):
        gbs = bytes_px * px / (ms / 1e3) / 1e9
        out = {"ms": round(ms, 4), "value": round(px / (ms / 1e3) / 1e6, 2), "unit": "Mpixel/s",
               "algorithmic_bytes_per_pixel": bytes_px, "achieved_gbs": round(gbs, 1),
               "hbm_frac": round(gbs / peak, 4), "note": note}
        out.update(extra)
        return out

    def cpu_figure(h, w, seed, q, mode, pf):
        """the oracle port (and the staged reference, if any) on ONE frame of the config"""
        from oracle import numpy_port as P
        img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        t0 = time.perf_counter()
        P.compress_reconstruct(img, q, mode, pf, want_maps=True)
        tp = time.perf_counter() - t0
        out = {"port_mpixel_s": round(h * w / tp / 1e6, 4), "cores": 1}
        if reference_staged():
            from oracle import reference_shim as R
            ref = R.load()
            t0 = time.perf_counter()
            ref.compress_reconstruct(img, ref.CompressionParams(quality=q, subsampling_mode=mode, use_prefilter=pf))
            out["reference_mpixel_s"] = round(h * w / (time.perf_counter() - t0) / 1e6, 4)
        return out

    results = {}
    # ---- cfg2: 1080p 4:4:4, batches of 32 distinct frames (199 MB in + 199 MB out > L2) ----
    h2, w2, n2 = 1080, 1920, 32
    f2 = torch.from_numpy(np.stack([np.random.default_rng(2 + 7 * k).integers(0, 256, (h2, w2, 3), dtype=np.uint8)
                                    for k in range(n2)])).to(dev)
    o2 = torch.empty_like(f2)
    r2 = torch.zeros((n2, NAT.JDS_RECORD_FIELDS), dtype=torch.float64, device=dev)
    cfg2 = {}
    for prec in ("fast", "exact"):
        ms = timed(lambda i: eng.batch_records(f2, r2, 50, "4:4:4", False, precision=prec, recon_out=o2), K if prec == "fast" else 3)
        cfg2[prec] = rec(ms, n2 * h2 * w2, 6.0, f"{n2} frames per call, device resident, recon + PSNR/SSIM/bpp")
    one = eng.roundtrip(f2[0], 50, "4:4:4", False, precision="fast")
    cfg2["results_frame0"] = {k: one.scalars[k] for k in ("psnr_y", "ssim_y", "bpp")}
    del f2, o2
    results["cfg2_1080p_q50_444"] = cfg2

    # ---- cfg3: 4K Q75 4:2:0 + prefilter, GUI outputs; 8 distinct frames cycle (199 MB > L2) ----
    f3 = torch.from_numpy(np.stack([np.random.default_rng(3 + 11 * k).integers(0, 256, (H, W, 3), dtype=np.uint8)
                                    for k in range(8)])).to(dev)
    px3 = H * W
    cfg3 = {}
    for prec in ("fast", "exact"):
        n = K if prec == "fast" else 3
        ms = timed(lambda i: eng.roundtrip(f3[i % 8], 75, "4:2:0", True, precision=prec), n)
        cfg3[f"{prec}_recon_only"] = rec(ms, px3, 6.0, "per-call API (synchronises), recon + metrics")
        ms = timed(lambda i: eng.roundtrip(f3[i % 8], 75, "4:2:0", True, precision=prec, want_coeffs=True, want_hist=True), n)
        cfg3[f"{prec}_coeffs_hist"] = rec(ms, px3, 9.0, "recon + int16 coefficients (3 B/px) + 50-bin histogram")
        ms = timed(lambda i: eng.roundtrip(f3[i % 8], 75, "4:2:0", True, precision=prec, want_coeffs=True, want_hist=True,
                                           want_error_maps=True), n)
        cfg3[f"{prec}_all_intermediates"] = rec(ms, px3, 25.0, "recon + coefficients + histogram + both fp64 error maps "
                                                "(8 B/px each): what IntermediateData holds")
        ms = timed(lambda i: eng.plot_payload(f3[i % 8], 75, "4:2:0", True, precision=prec), n)
        cfg3[f"{prec}_plot_payload"] = rec(ms, px3, 10.0, "recon + uint8 error heat map (1 B/px) + per-value coefficient "
                                           "counts: what the GUI plots draw (coefficients stay on the device, 3 B/px)")
    ms = timed(lambda i: eng.selected_block(f3[i % 8], 75, 0, 0), K)
    cfg3["selected_block_dct_ms"] = round(ms, 4)
    del f3
    results["cfg3_4k_q75_420_pf"] = cfg3

    # ---- cfg5: 1024 x 1080p Q30 4:2:2 sharded by frame; 64 distinct frames per rank, tiled ----
    total5 = 1024
    mine = len(range(rank, total5, world))
    base = torch.from_numpy(np.stack([np.random.default_rng(5000 + rank + world * k).integers(0, 256, (h2, w2, 3), dtype=np.uint8)
                                      for k in range(64)])).to(dev)
    per_call = 64
    o5 = torch.empty_like(base)
    r5 = torch.zeros((per_call, NAT.JDS_RECORD_FIELDS), dtype=torch.float64, device=dev)
    calls = (mine + per_call - 1) // per_call

    def job(i):
        for c in range(calls):
            n = min(per_call, mine - c * per_call)
            eng.batch_records(base[:n], r5[:n], 30, "4:2:2", False, precision="fast", recon_out=o5[:n],
                              unit0=rank, unit_step=world)
        if world > 1:
            dist.all_reduce(r5)              # the job's one metric exchange
    ms5 = timed(job, 3, warm=1)
    px5 = total5 * h2 * w2
    cfg5 = rec(ms5, px5, 6.0, f"{total5} frames over {world} GPU(s): {mine} per rank in calls of {per_call} "
               "(64 distinct frames per rank, re-used: 398 MB in + 398 MB out per call > L2), device resident, "
               "one all-reduce of the records per job")
    cfg5["gbs_per_gpu"] = round(6.0 * px5 / world / (ms5 / 1e3) / 1e9, 1)
    cfg5["hbm_frac_per_gpu"] = round(cfg5["gbs_per_gpu"] / peak, 4)
    cfg5["hbm_frac"] = cfg5["hbm_frac_per_gpu"]
    del base, o5
    results["cfg5_1024x1080p_q30_422"] = cfg5

    line = None
    if rank == 0:
        if world == 1:
            results["cfg2_1080p_q50_444"]["cpu"] = cpu_figure(h2, w2, 2, 50, "4:4:4", False)
            results["cfg3_4k_q75_420_pf"]["cpu"] = cpu_figure(H, W, 3, 75, "4:2:0", True)
            results["cfg5_1024x1080p_q30_422"]["cpu"] = cpu_figure(h2, w2, 5000, 30, "4:2:2", False)
        line = {"metric": "Mpixel/s of the other BASELINE configs (round trip incl. PSNR/SSIM/bpp)",
                "value": results["cfg5_1024x1080p_q30_422"]["value"], "unit": "Mpixel/s", "n_gpus": world,
                "steps": K, "warmup": 2, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "peak_gbs": peak, "peak_source": peak_src,
                "config": {"workload": "BASELINE configs 2, 3, 5; value = cfg5 aggregate"},
                "configs": results}
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="batch", choices=["batch", "sweep", "configs"],
                    help="batch = the headline (16 x 4K frames per GPU, weak scaling; carries the sweep "
                         "sub-record); sweep = BASELINE config 4 alone (100-point sweep, strong "
                         "scaling); configs = BASELINE configs 2, 3 and 5")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    with StdoutToStderr():
        line = {"sweep": run_sweep, "configs": run_configs}.get(args.workload, run_ours)(args)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
