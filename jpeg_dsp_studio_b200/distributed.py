"""Multi-GPU sharding of sweeps and batches (BASELINE.json configs 4 and 5).

One process per GPU (``torchrun``); units - quality points of a sweep, frames of a
batch - are independent, so they shard round-robin (unit k -> rank k mod world,
SURVEY.md §8e) with NO data-path collective.  The only exchange is the metric
records: one ``all_gather`` per sweep (every rank ends up with the whole
rate-distortion table) or one ``all_reduce`` per batch (aggregate partials), over
NCCL on GPUs (gloo in the CPU tests).  Integer partials (SSE_rgb, bits, nnz) are
summed exactly: fp64 holds integers up to 2^53.
"""

from typing import List, Optional, Sequence

import numpy as np

#: fields of one unit's record, in the order they travel
RECORD_FIELDS = ("unit", "quality", "sse_rgb", "sse_y", "ssim_r", "ssim_g", "ssim_b", "ssim_y",
                 "ssim_count", "coeff_bits", "nnz", "total_coeffs", "luma_blocks")


def shard_indices(n_units: int, rank: int, world: int) -> List[int]:
    """Units owned by ``rank``: k with k mod world == rank."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_units, world))


def record_from_metrics(unit: int, quality: int, m) -> np.ndarray:
    """One fp64 row from a ``jds_metrics`` struct (or anything with the same fields)."""
    return np.array([unit, quality, float(m.sse_rgb), float(m.sse_y), float(m.ssim_sum[0]),
                     float(m.ssim_sum[1]), float(m.ssim_sum[2]), float(m.ssim_sum[3]),
                     float(m.ssim_count), float(m.coeff_bits), float(m.nnz),
                     float(m.total_coeffs), float(m.luma_blocks)], dtype=np.float64)


#: numpy view of a ctypes array of jds_metrics (include/jds.h), for vectorised record building
METRICS_DTYPE = np.dtype([("sse_rgb", "<u8"), ("sse_y", "<f8"), ("ssim_sum", "<f8", (4,)),
                          ("ssim_count", "<u8"), ("coeff_bits", "<u8"), ("nnz", "<u8"),
                          ("total_coeffs", "<u8"), ("luma_blocks", "<u8"), ("hist50", "<i8", (50,)),
                          ("gpu_ms", "<f8"), ("reserved", "<u8", (3,))])


def records_from_outputs(units, qualities, outs) -> np.ndarray:
    """(n, F) fp64 rows for a list of RoundTripOutputs of one call.  When the outputs share
    one ctypes array of jds_metrics (the Engine's batch / sweep calls) the rows are built
    from a zero-copy numpy view of it."""
    n = len(outs)
    if n == 0:
        return np.zeros((0, len(RECORD_FIELDS)))
    arr = getattr(outs[0], "metrics_array", None)
    if arr is None:
        return np.stack([record_from_metrics(u, q, o.metrics) for u, q, o in zip(units, qualities, outs)])
    m = np.frombuffer(arr, dtype=METRICS_DTYPE, count=n)
    rows = np.empty((n, len(RECORD_FIELDS)), dtype=np.float64)
    rows[:, 0] = units
    rows[:, 1] = qualities
    rows[:, 2] = m["sse_rgb"]
    rows[:, 3] = m["sse_y"]
    rows[:, 4:8] = m["ssim_sum"]
    rows[:, 8] = m["ssim_count"]
    rows[:, 9] = m["coeff_bits"]
    rows[:, 10] = m["nnz"]
    rows[:, 11] = m["total_coeffs"]
    rows[:, 12] = m["luma_blocks"]
    return rows


def scalars_from_table(table: np.ndarray, height: int, width: int) -> List[dict]:
    """The reference's result floats for every row of a gathered table at once
    (utils/metrics.py formulas, vectorised: one NumPy expression per field)."""
    t = np.asarray(table, dtype=np.float64).reshape(-1, len(RECORD_FIELDS))
    f = {k: t[:, i] for i, k in enumerate(RECORD_FIELDS)}
    n_px = height * width
    with np.errstate(divide="ignore", invalid="ignore"):
        psnr_rgb = 10 * np.log10((255.0 ** 2) / (f["sse_rgb"] / np.float64(3 * n_px)))
        psnr_y = 10 * np.log10((255.0 ** 2) / (f["sse_y"] / np.float64(n_px)))
        cnt = f["ssim_count"]
        ssim_rgb = np.where(cnt > 0, (np.stack([f["ssim_r"], f["ssim_g"], f["ssim_b"]]) / cnt).mean(axis=0), np.nan)
        ssim_y = np.where(cnt > 0, f["ssim_y"] / cnt, np.nan)
    # bit estimate in the reference's float32 arithmetic (utils/metrics.bitrate_from_partials)
    exact = (2 * f["luma_blocks"] + f["coeff_bits"]).astype(np.int64)
    nnz = f["nnz"].astype(np.int64)
    est32 = (2 * f["luma_blocks"]).astype(np.float32) + \
        ((6 * nnz).astype(np.float32) + (f["coeff_bits"].astype(np.int64) - 6 * nnz).astype(np.float32))
    bpp32 = est32 / np.float32(n_px)
    ratio32 = np.float32(n_px * 24) / np.maximum(est32, np.float32(1))
    none = nnz == 0                                   # the reference stays in Python ints then
    bits = np.where(none, exact, est32.astype(np.int64))
    bpp = np.where(none, exact / n_px, bpp32.astype(np.float64))
    ratio = np.where(none, (n_px * 24) / np.maximum(exact, 1), ratio32.astype(np.float64))
    # ndarray.tolist() yields Python ints / floats in C; one dict(zip()) per row (a 100-point table:
    # 0.16 ms with per-element float() / int() calls, 0.05 ms this way - it sits on a sweep's latency)
    cols = (f["quality"].astype(np.int64).tolist(), psnr_rgb.tolist(), psnr_y.tolist(), ssim_rgb.tolist(),
            ssim_y.tolist(), bits.astype(np.int64).tolist(), exact.tolist(), bpp.astype(np.float64).tolist(),
            ratio.astype(np.float64).tolist(), nnz.tolist(), f["total_coeffs"].astype(np.int64).tolist())
    keys = ("quality", "psnr_rgb", "psnr_y", "ssim_rgb", "ssim_y", "estimated_bits", "exact_bits", "bpp",
            "compression_ratio", "nonzero_count", "total_coeffs")
    return [dict(zip(keys, row)) for row in zip(*cols)]


def scalars_from_record(rec: np.ndarray, height: int, width: int) -> dict:
    """One row (see scalars_from_table)."""
    return scalars_from_table(np.asarray(rec).reshape(1, -1), height, width)[0]


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


_gather_cache = {}


def gather_records(local: np.ndarray, n_units: int, device=None) -> np.ndarray:
    """All-gather of per-unit records: every rank returns the (n_units, F) table ordered
    by unit.  ``local``: (n_local, F) rows whose first column is the unit index.  Staging
    buffers (pinned host + device) are cached per (world, capacity, device): one H2D, one
    collective, one D2H, one synchronisation."""
    dist = _dist()
    nf = len(RECORD_FIELDS)
    local = np.asarray(local, dtype=np.float64).reshape(-1, nf)
    if dist is None or dist.get_world_size() == 1:
        table = local
    else:
        import torch
        world = dist.get_world_size()
        cap = (n_units + world - 1) // world                 # rows per rank, padded
        key = (world, cap, str(device))
        bufs = _gather_cache.get(key)
        if bufs is None:
            on_gpu = device is not None and torch.device(device).type == "cuda"
            h_in = torch.empty((cap, nf), dtype=torch.float64)
            h_out = torch.empty((world * cap, nf), dtype=torch.float64)
            if on_gpu:
                h_in, h_out = h_in.pin_memory(), h_out.pin_memory()
                d_in = torch.empty((cap, nf), dtype=torch.float64, device=device)
                d_out = torch.empty((world * cap, nf), dtype=torch.float64, device=device)
            else:
                d_in, d_out = h_in, h_out
            bufs = _gather_cache[key] = (h_in, h_out, d_in, d_out, on_gpu)
        h_in, h_out, d_in, d_out, on_gpu = bufs
        hv = h_in.numpy()
        hv[:] = -1.0
        hv[:len(local)] = local
        if on_gpu:
            d_in.copy_(h_in, non_blocking=True)
        dist.all_gather_into_tensor(d_out, d_in)
        if on_gpu:
            h_out.copy_(d_out, non_blocking=True)
            torch.cuda.current_stream(d_out.device).synchronize()
        table = h_out.numpy().copy()
        table = table[table[:, 0] >= 0]
    order = np.argsort(table[:, 0], kind="stable")
    table = table[order]
    if len(table) != n_units or not np.array_equal(table[:, 0], np.arange(n_units)):
        raise RuntimeError("sharded units do not cover 0..n_units-1 exactly once")
    return table


def reduce_partials(local_sum: np.ndarray, device=None) -> np.ndarray:
    """All-reduce (sum) of aggregate partials; integer-valued entries stay exact < 2^53."""
    dist = _dist()
    v = np.asarray(local_sum, dtype=np.float64)
    if dist is None or dist.get_world_size() == 1:
        return v.copy()
    import torch
    t = torch.from_numpy(v.copy()).to(device) if device is not None else torch.from_numpy(v.copy())
    dist.all_reduce(t)
    return t.cpu().numpy()


class SweepHandle:
    """A sharded sweep in flight (``sweep_sharded_begin``): kernels, all-gather and the D2H
    copy of the table are enqueued; ``result()`` waits for them and returns the per-quality
    dicts.  Several handles may be in flight - each owns its buffers until ``result()``."""

    def __init__(self, key, bufs, event, n_units, hw, keep):
        self._key, self._bufs, self._event = key, bufs, event
        self._n, self._hw, self._keep, self._table = n_units, hw, keep, None

    def table(self) -> np.ndarray:
        """The gathered (n_units, F) record table, ordered by unit."""
        if self._table is None:
            self._event.synchronize()
            t = self._bufs[2].numpy()
            t = t[t[:, 0] >= 0]
            t = t[np.argsort(t[:, 0], kind="stable")]       # fancy indexing: a copy
            _gather_pool.setdefault(self._key, []).append(self._bufs)
            self._bufs = self._keep = None
            if len(t) != self._n or not np.array_equal(t[:, 0], np.arange(self._n)):
                raise RuntimeError("sharded units do not cover 0..n_units-1 exactly once")
            self._table = t
        return self._table

    def result(self) -> List[dict]:
        return scalars_from_table(self.table(), *self._hw)


_gather_pool = {}


def sweep_sharded_begin(engine, image, qualities: Sequence[int], mode="4:2:0", prefilter=False, *,
                        precision="fast", device=None) -> SweepHandle:
    """Device-resident sharded sweep, asynchronous half: the kernels write this rank's
    records straight into the all-gather's send buffer (``jds_sweep_records``), the
    collective and the single D2H copy follow on the same stream, and NOTHING waits for the
    GPU here - no metric read-back, no host-built rows, no H2D in between.  A caller that
    sweeps many frames keeps one sweep in flight while it finalises the previous one."""
    import torch
    from . import _native as N
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    qs = [int(q) for q in qualities]
    mine = shard_indices(len(qs), rank, world)
    dev = torch.device(device if device is not None else f"cuda:{engine.device}")
    cur = torch.cuda.current_stream(dev)
    if getattr(engine, "_stream_handle", None) != cur.cuda_stream:
        engine.use_stream(cur.cuda_stream)      # kernels, collective and copy share one stream
    nf = len(RECORD_FIELDS)
    cap = max(1, (len(qs) + world - 1) // world)
    if cap > N.JDS_SWEEP_RECORDS_MAX:
        raise ValueError(f"at most {N.JDS_SWEEP_RECORDS_MAX} sweep points per rank and call")
    key = ("dev", world, cap, str(dev))
    pool = _gather_pool.setdefault(key, [])
    if pool:
        bufs = pool.pop()
    else:
        d_in = torch.empty((cap, nf), dtype=torch.float64, device=dev)
        d_out = torch.empty((world * cap, nf), dtype=torch.float64, device=dev) if world > 1 else d_in
        h_out = torch.empty((world * cap, nf), dtype=torch.float64).pin_memory()
        bufs = (d_in, d_out, h_out)
    d_in, d_out, h_out = bufs
    keep = engine.sweep_records(image, [qs[i] for i in mine], d_in, mode=mode, prefilter=prefilter,
                                precision=precision, unit0=(mine[0] if mine else 0), unit_step=world)
    if world > 1:
        dist.all_gather_into_tensor(d_out, d_in)
    h_out.copy_(d_out, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(cur)
    return SweepHandle(key, bufs, ev, len(qs), (image.shape[0], image.shape[1]), (keep, image))


def sweep_sharded(engine, image, qualities: Sequence[int], mode="4:2:0", prefilter=False, *,
                  precision="fast", device=None) -> List[dict]:
    """Rate-distortion sweep with the points sharded over the ranks (config 4).
    Every rank returns the full list of per-quality result dicts.  With a CUDA ``device``
    the records never leave the GPU before the all-gather (one synchronisation per sweep)."""
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    qs = [int(q) for q in qualities]
    mine = shard_indices(len(qs), rank, world)
    h, w = image.shape[0], image.shape[1]
    on_gpu = False
    if device is not None and len(qs) > 0:
        import torch
        on_gpu = torch.device(device).type == "cuda"
    from . import _native as N
    if on_gpu and (len(qs) + world - 1) // world <= N.JDS_SWEEP_RECORDS_MAX:
        return sweep_sharded_begin(engine, image, qs, mode, prefilter, precision=precision,
                                   device=device).result()
    rows = np.zeros((0, len(RECORD_FIELDS)))
    if mine:
        my_qs = [qs[i] for i in mine]
        outs = engine.sweep(image, my_qs, mode, prefilter, precision=precision)
        rows = records_from_outputs(mine, my_qs, outs)
    table = gather_records(rows, len(qs), device=device)
    return scalars_from_table(table, h, w)


def batch_sharded(engine, frames_of_rank, n_total: int, quality=50, mode="4:2:0", prefilter=False, *,
                  precision="fast", device=None) -> dict:
    """Batch with the frames sharded over the ranks (config 5): ``frames_of_rank`` holds
    this rank's frames (k mod world == rank, in order).  Returns the aggregate over ALL
    frames (mean PSNR from the summed squared error, mean SSIM, total bits)."""
    outs = engine.roundtrip_batch(frames_of_rank, quality, mode, prefilter, precision=precision,
                                  want_recon=False) if len(frames_of_rank) else []
    h, w = (frames_of_rank.shape[1], frames_of_rank.shape[2]) if len(frames_of_rank) else (1, 1)
    part = np.zeros(11, dtype=np.float64)
    for o in outs:
        m = o.metrics
        part += np.array([1, float(m.sse_rgb), float(m.sse_y), m.ssim_sum[0], m.ssim_sum[1],
                          m.ssim_sum[2], m.ssim_sum[3], float(m.ssim_count), float(m.coeff_bits),
                          float(m.nnz), float(m.luma_blocks)])
    geo = reduce_partials(np.array([h, w], dtype=np.float64) * (1.0 if len(outs) else 0.0), device)
    tot = reduce_partials(part, device=device)
    n = int(round(tot[0]))
    if n != n_total:
        raise RuntimeError(f"ranks processed {n} frames, expected {n_total}")
    from .utils.metrics import psnr_from_sse
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    contributing = max(1, min(world, n_total))
    h, w = int(round(geo[0] / contributing)), int(round(geo[1] / contributing))
    n_px = h * w * n
    bits = 2 * int(round(tot[10])) + int(round(tot[8]))
    return {
        "frames": n, "height": h, "width": w,
        "psnr_rgb": psnr_from_sse(tot[1], 3 * n_px), "psnr_y": psnr_from_sse(tot[2], n_px),
        "ssim_rgb": float((tot[3] + tot[4] + tot[5]) / 3.0 / tot[7]) if tot[7] else float("nan"),
        "ssim_y": float(tot[6] / tot[7]) if tot[7] else float("nan"),
        "estimated_bits": bits, "bpp": float(bits / n_px), "nonzero_count": int(round(tot[9])),
    }


# ---------------------------------------------------------------------------------------
# tile-band sharding of ONE frame (second row): single-image latency
# ---------------------------------------------------------------------------------------
BAND_ALIGN = 16       # rows of one 4:2:0 MCU; band edges of jds_roundtrip_band


def band_bounds(height: int, world: int) -> List[Optional[tuple]]:
    """Rows ``(row0, row1)`` of every rank: the frame's 16-row MCU rows split as evenly as
    possible, in order; ranks beyond the number of MCU rows get ``None``."""
    if height < 1 or world < 1:
        raise ValueError(f"bad height/world {height}/{world}")
    mcu = (height + BAND_ALIGN - 1) // BAND_ALIGN
    out, at = [], 0
    for r in range(world):
        n = mcu // world + (1 if r < mcu % world else 0)
        if n == 0:
            out.append(None)
            continue
        out.append((at * BAND_ALIGN, min((at + n) * BAND_ALIGN, height)))
        at += n
    return out


def merge_band_records(records) -> np.ndarray:
    """Sum of the bands' partial records (``record_from_metrics`` rows) = the frame's record:
    every field after ``unit`` and ``quality`` is a sum over rows, windows or blocks."""
    recs = np.asarray(records, dtype=np.float64).reshape(-1, len(RECORD_FIELDS))
    out = recs.sum(axis=0)
    out[0], out[1] = recs[0, 0], recs[0, 1]
    return out


def frame_banded(engine, image, quality=50, mode="4:2:0", prefilter=False, *, precision="exact",
                 device=None, gather_recon=True) -> dict:
    """One frame with its rows sharded over the ranks: this rank runs its band (plus the halo
    ``jds_roundtrip_band`` adds), the partial metrics are summed with one ``all_reduce`` and -
    if ``gather_recon`` - the reconstructed rows are exchanged with one ``all_gather``, so every
    rank returns the whole frame's scalars (and image).  ``image``: the whole frame on every
    rank (only the band and its halo are read).  Without an initialised process group the
    single rank owns the whole frame."""
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    h, w = int(image.shape[0]), int(image.shape[1])
    bounds = band_bounds(h, world)
    mine = bounds[rank]
    rec = np.zeros(len(RECORD_FIELDS), dtype=np.float64)
    rows = None
    if mine is not None:
        out = engine.roundtrip_band(image, mine[0], mine[1], quality, mode, prefilter, precision=precision)
        rec = record_from_metrics(0, quality, out.metrics)
        rows = out.recon
    rec[0], rec[1] = 0.0, 0.0                       # summed below; restored after the reduce
    total = reduce_partials(rec, device=device)
    total[0], total[1] = 0.0, float(quality)
    result = {"band": mine, "bounds": bounds, "record": total,
              "scalars": scalars_from_record(total, h, w), "recon_rows": rows, "recon": None}
    if gather_recon:
        result["recon"] = _gather_rows(rows, bounds, h, w, dist, device)
    return result


def _gather_rows(rows, bounds, h, w, dist, device):
    """All-gather of the bands' reconstructed rows (padded to the tallest band)."""
    import torch
    if dist is None or dist.get_world_size() == 1:
        return rows
    cap = max((b[1] - b[0]) for b in bounds if b is not None)
    on_gpu = device is not None and torch.device(device).type == "cuda"
    dev = device if on_gpu else "cpu"
    key = ("rows", len(bounds), cap, w, str(dev))
    bufs = _gather_cache.get(key)
    if bufs is None:
        bufs = _gather_cache[key] = (torch.zeros((cap, w, 3), dtype=torch.uint8, device=dev),
                                     torch.empty((len(bounds) * cap, w, 3), dtype=torch.uint8, device=dev))
    send, recv = bufs
    if rows is not None:
        t = rows if isinstance(rows, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(rows))
        send[:t.shape[0]].copy_(t)
    dist.all_gather_into_tensor(recv, send)
    parts = [recv[r * cap:r * cap + (b[1] - b[0])] for r, b in enumerate(bounds) if b is not None]
    full = torch.cat(parts, dim=0)
    assert full.shape[0] == h
    if isinstance(rows, torch.Tensor) and rows.is_cuda:
        return full
    return full.cpu().numpy()
