"""Engine: one libjds context on one GPU, with NumPy / torch buffer handoff.

This is the host-side object behind the drop-in ``compress_reconstruct``
(engines/pipeline.py) and behind the batch / sweep entry points that BASELINE.json's
configs 4 and 5 need.  It owns no arithmetic: every pixel, coefficient and metric
partial comes from the CUDA kernels through the C ABI (``_native``).  PyTorch is
used only where the caller hands over device tensors or pinned host tensors
(``data_ptr()``), never for compute.
"""

import ctypes as C
import threading
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N
from .models import CompressionParams
from .utils.metrics import bitrate_from_partials, metrics_from_partials


class RoundTripOutputs:
    """Raw outputs of one unit (frame or sweep point).

    ``metrics`` is the ``jds_metrics`` record of device partials; ``scalars`` (PSNR / SSIM /
    bpp / compression ratio ... computed from the partials with the reference's final
    formulas) is evaluated on first access."""

    __slots__ = ("recon", "coeffs", "err_y", "err_rgb", "metrics", "metrics_array", "_hw", "_scalars")

    def __init__(self, recon, coeffs, err_y, err_rgb, metrics, hw, metrics_array=None):
        self.recon, self.coeffs, self.err_y, self.err_rgb = recon, coeffs, err_y, err_rgb
        self.metrics, self._hw, self._scalars = metrics, hw, None
        self.metrics_array = metrics_array      # the call's whole ctypes array of jds_metrics

    @property
    def scalars(self) -> dict:
        if self._scalars is None:
            h, w = self._hw
            out = metrics_from_partials(self.metrics, h, w)
            out.update(bitrate_from_partials(self.metrics, h, w))
            self._scalars = out
        return self._scalars


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


_PINNED_LIMIT = 1 << 30      # do not pin more than this per output array


def host_array(shape, dtype) -> np.ndarray:
    """A fresh NumPy array for results that come back from the device.

    Backed by page-locked memory from torch's caching host allocator when available: the
    device->host copy then runs at PCIe speed instead of the ~6 GB/s of a pageable
    destination (a 4K frame's fp64 error maps + coefficients are 183 MB), and blocks freed
    by earlier results are recycled without a new cudaHostAlloc.  The array owns its
    storage like any other (the tensor is its base); plain ``np.empty`` is the fallback."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if 0 < n <= _PINNED_LIMIT:
        try:
            import torch
            if torch.cuda.is_available():
                tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.int16): torch.int16,
                       np.dtype(np.float64): torch.float64}[np.dtype(dtype)]
                return torch.empty(tuple(shape), dtype=tdt, pin_memory=True).numpy()
        except Exception:
            pass
    return np.empty(shape, dtype=dtype)


def _mode_code(mode) -> int:
    try:
        return N.SUBSAMPLING[mode]
    except (KeyError, TypeError):
        # same exception type and text as engines/color_space.py:51
        raise ValueError(f"Unknown subsampling mode: {mode}") from None


def _check_qualities(qs) -> None:
    """The reference's validation / message (models/compression_params.py:16-17) for every sweep
    point, without constructing a dataclass per point (a 100-point sweep spent 0.2 ms there)."""
    lo, hi = CompressionParams.QUALITY_RANGE
    for q in qs:
        if q < lo or q > hi:
            CompressionParams(quality=q)            # raises with the reference's message


def _precision_code(precision) -> int:
    try:
        return N.PRECISION[precision]
    except (KeyError, TypeError):
        raise ValueError(f"precision must be 'exact' or 'fast', got {precision!r}") from None


def histogram_from_values(value_hist: np.ndarray, bins: int = 50):
    """``np.histogram(coeffs, bins)`` - what matplotlib's ``ax.hist(coeffs, bins=bins)``
    computes (gui/widgets/mpl_canvas.py:96) - from the per-value counts: the distinct
    values go through NumPy's own binning as an int16 array weighted by their counts, so the
    bin edges and every count equal those of the full coefficient array (counts < 2^53 are
    exact as fp64 weights).  Returns (int64 counts, fp64 edges)."""
    value_hist = np.asarray(value_hist)
    present = np.nonzero(value_hist)[0]
    if present.size == 0:
        return np.histogram(np.zeros(0, dtype=np.int16), bins=bins)
    values = (present - len(value_hist) // 2).astype(np.int16)
    counts, edges = np.histogram(values, bins=bins, range=(float(values.min()), float(values.max())),
                                 weights=value_hist[present].astype(np.float64))
    return counts.astype(np.int64), edges


class PlotPayload:
    """Result of Engine.plot_payload: ``reconstructed_image`` (uint8 H x W x 3),
    ``error_heat_y`` / ``error_heat_rgb`` (uint8 H x W = trunc(clip(error_map * 10, 0, 255)),
    gui/widgets/mpl_canvas.py:116-118), ``value_hist`` (int64[2048], value v at v + 1024),
    ``hist_counts`` / ``hist_edges`` (the 50-bin histogram matplotlib would draw), and the
    scalar metrics through ``outputs.scalars``."""
    __slots__ = ("reconstructed_image", "error_heat_y", "error_heat_rgb", "value_hist",
                 "hist_counts", "hist_edges", "outputs")

    def __init__(self, recon, heat_y, heat_rgb, value_hist, counts, edges, outputs):
        self.reconstructed_image = recon
        self.error_heat_y = heat_y
        self.error_heat_rgb = heat_rgb
        self.value_hist = value_hist
        self.hist_counts = counts
        self.hist_edges = edges
        self.outputs = outputs

    @property
    def scalars(self):
        return self.outputs.scalars


class PendingBatch:
    """A batch in flight (``Engine.roundtrip_batch_begin``); keeps its buffers alive."""

    def __init__(self, engine, frames, recon, metrics, hw, loc):
        self._engine, self._frames, self._recon, self._ms, self._hw, self._loc = engine, frames, recon, metrics, hw, loc
        self._outs = None

    def result(self) -> List[RoundTripOutputs]:
        if self._outs is None:
            self._engine._finish()                 # synchronises the context's streams
            r, ms = self._recon, self._ms
            self._outs = [RoundTripOutputs(r[i] if r is not None else None, None, None, None, ms[i], self._hw, ms)
                          for i in range(len(ms))]
        return self._outs


class Engine:
    """A libjds context bound to one CUDA device."""

    def __init__(self, device: int = 0):
        self._lib = N.load()
        self._ctx = C.c_void_p()
        self.device = int(device)
        N.check(self._lib.jds_ctx_create(self.device, C.byref(self._ctx)))
        self._lock = threading.Lock()        # one call in flight per context

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.jds_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------
    def launch_count(self) -> int:
        n = C.c_uint64()
        N.check(self._lib.jds_ctx_launch_count(self._ctx, C.byref(n)))
        return int(n.value)

    STAGES = ("forward_colour", "block_codec", "inverse_colour", "ssim")

    def set_stage_timing(self, enable: bool):
        """Per-kernel CUDA events on/off (off by default; while on, launch sequences run on
        one stream so that kernel times do not overlap)."""
        N.check(self._lib.jds_ctx_stage_timing(self._ctx, int(bool(enable))))
        self._stage_timing = bool(enable)

    def stage_times(self, reset: bool = True) -> dict:
        """Accumulated device ms and launch count per stage kernel since the last reset.
        The first call switches the per-kernel events on (they are off by default)."""
        if not getattr(self, "_stage_timing", False):
            N.check(self._lib.jds_ctx_stage_timing(self._ctx, 1))
            self._stage_timing = True
        ms = (C.c_double * 4)()
        ln = (C.c_uint64 * 4)()
        N.check(self._lib.jds_ctx_stage_times(self._ctx, ms, ln, int(reset)))
        return {name: {"ms": float(ms[i]), "launches": int(ln[i])}
                for i, name in enumerate(self.STAGES)}

    def use_stream(self, cuda_stream: int):
        """Issue this context's work on the caller's cudaStream_t (an int handle)."""
        N.check(self._lib.jds_ctx_set_stream(self._ctx, C.c_void_p(cuda_stream)))
        self._stream_handle = int(cuda_stream)

    def synchronize(self):
        N.check(self._lib.jds_ctx_synchronize(self._ctx))

    @staticmethod
    def _frame_geometry(image) -> tuple:
        shape = tuple(image.shape)
        if len(shape) < 3:
            # the reference indexes rgb[:, :, 0] on a 2-D array (color_space.py:10)
            raise IndexError("too many indices for array: array is 2-dimensional, "
                             "but 3 were indexed")
        if shape[-1] != 3:
            raise ValueError(f"expected an RGB image with 3 channels, got shape {shape}")
        return shape

    def _params(self, h, w, quality, mode, prefilter, precision, outputs) -> "N.JdsParams":
        return N.JdsParams(int(h), int(w), int(quality), _mode_code(mode), int(bool(prefilter)),
                           _precision_code(precision), int(outputs), 0)

    def _order_after_torch(self, t):
        """A CUDA tensor handed to this engine was produced on torch's current stream of its
        device, while the context issues its kernels on its own stream: make the context's
        stream wait for what torch has enqueued so far (an event, no host synchronisation).
        Nothing to do when the context already runs on that stream (``use_stream``)."""
        import torch
        if t.device.index != self.device:
            raise ValueError(f"tensor lives on cuda:{t.device.index} but this engine is bound to "
                             f"cuda:{self.device} (use get_engine({t.device.index}))")
        cur = torch.cuda.current_stream(t.device)
        if getattr(self, "_stream_handle", None) == cur.cuda_stream:
            return
        ev = torch.cuda.Event()
        ev.record(cur)
        N.check(self._lib.jds_ctx_wait_event(self._ctx, C.c_void_p(ev.cuda_event)))

    def _order_torch_after(self, device_tensor):
        """After a NON-synchronising entry point: torch's current stream waits for the
        context's stream, so the caller may consume the results with ordinary torch ops."""
        import torch
        cur = torch.cuda.current_stream(device_tensor.device)
        if getattr(self, "_stream_handle", None) == cur.cuda_stream:
            return
        ev = torch.cuda.Event()
        ev.record(cur)                 # creates the underlying cudaEvent_t
        N.check(self._lib.jds_ctx_record_event(self._ctx, C.c_void_p(ev.cuda_event)))
        cur.wait_event(ev)

    def _in_ptr(self, image):
        """(pointer, location, keepalive) for a uint8 NumPy array or torch tensor."""
        if _is_torch(image):
            import torch
            if image.dtype != torch.uint8:
                raise TypeError(f"image tensor must be uint8, got {image.dtype}")
            t = image.contiguous()
            loc = N.JDS_DEVICE if t.is_cuda else N.JDS_HOST
            if t.is_cuda:
                self._order_after_torch(t)
            return C.c_void_p(t.data_ptr()), loc, t
        a = np.ascontiguousarray(image)
        if a.dtype != np.uint8:
            raise TypeError(f"image must be uint8, got {a.dtype}")
        return C.c_void_p(a.ctypes.data), N.JDS_HOST, a

    # -- single frame ----------------------------------------------------------------
    def roundtrip(self, image, quality=50, mode="4:2:0", prefilter=False, *,
                  precision="exact", want_coeffs=False, want_error_maps=False,
                  want_hist=False, want_ssim=True, recon_out=None) -> RoundTripOutputs:
        """One frame through the kernels.  ``image``: uint8 H x W x 3 NumPy array
        (host) or torch tensor (host or CUDA).  Outputs follow the input's location
        (CUDA tensor in -> CUDA tensors out)."""
        h, w, _ = self._frame_geometry(image)
        ptr, loc, keep = self._in_ptr(image)
        flags = N.JDS_OUT_RECON | N.JDS_OUT_PSNR
        flags |= N.JDS_OUT_COEFFS if want_coeffs else 0
        flags |= (N.JDS_OUT_ERR_Y | N.JDS_OUT_ERR_RGB) if want_error_maps else 0
        flags |= N.JDS_OUT_HIST if want_hist else 0
        flags |= N.JDS_OUT_SSIM if want_ssim else 0
        p = self._params(h, w, quality, mode, prefilter, precision, flags)
        ncoef = C.c_uint64()
        N.check(self._lib.jds_coeff_count(h, w, p.subsampling, C.byref(ncoef)))
        m = N.JdsMetrics()
        if loc == N.JDS_DEVICE:
            import torch
            dev = keep.device
            recon = recon_out if recon_out is not None else torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
            coeffs = torch.empty(ncoef.value, dtype=torch.int16, device=dev) if want_coeffs else None
            ey = torch.empty((h, w), dtype=torch.float64, device=dev) if want_error_maps else None
            ergb = torch.empty((h, w), dtype=torch.float64, device=dev) if want_error_maps else None
            gp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        else:
            recon = recon_out if recon_out is not None else host_array((h, w, 3), np.uint8)
            coeffs = host_array((ncoef.value,), np.int16) if want_coeffs else None
            ey = host_array((h, w), np.float64) if want_error_maps else None
            ergb = host_array((h, w), np.float64) if want_error_maps else None
            if _is_torch(recon):
                gp = lambda t: C.c_void_p(t.data_ptr() if _is_torch(t) else t.ctypes.data) if t is not None else None
            else:
                gp = lambda t: C.c_void_p(t.ctypes.data) if t is not None else None
        with self._lock:
            N.check(self._lib.jds_roundtrip(self._ctx, C.byref(p), ptr, loc, gp(recon), gp(coeffs),
                                            gp(ey), gp(ergb), loc, C.byref(m)))
        return RoundTripOutputs(recon, coeffs, ey, ergb, m, (h, w))

    # -- tile-band sharding of one frame -----------------------------------
    def roundtrip_band(self, image, row0: int, row1: int, quality=50, mode="4:2:0", prefilter=False, *,
                       precision="exact", want_coeffs=False, want_ssim=True) -> RoundTripOutputs:
        """Rows ``[row0, row1)`` of ``image`` as one rank's share of the round trip
        (``jds_roundtrip_band``): ``recon`` holds only those rows, ``coeffs`` only the band's
        blocks (Y | Cb | Cr), ``metrics`` the band's PARTIAL sums - add the partials of all
        bands (``distributed.merge_band_records``) before turning them into PSNR / SSIM / bpp.
        ``row0`` / ``row1``: multiples of 16, or the frame height."""
        h, w, _ = self._frame_geometry(image)
        ptr, loc, keep = self._in_ptr(image)
        flags = N.JDS_OUT_RECON | N.JDS_OUT_PSNR
        flags |= N.JDS_OUT_COEFFS if want_coeffs else 0
        flags |= N.JDS_OUT_SSIM if want_ssim else 0
        p = self._params(h, w, quality, mode, prefilter, precision, flags)
        rows = int(row1) - int(row0)
        if rows <= 0 or row0 < 0 or row1 > h:
            raise ValueError(f"band [{row0}, {row1}) outside the {h}-row frame")
        ncoef = 0
        if want_coeffs:
            v = 2 if mode == "4:2:0" else 1
            hdiv = 1 if mode == "4:4:4" else 2
            nb = lambda n: (n + 7) // 8
            by = (nb(h) if row1 == h else row1 // 8) - row0 // 8
            cy = (nb(h // v) if row1 == h else row1 // (8 * v)) - row0 // (8 * v)
            ncoef = 64 * (by * nb(w) + 2 * cy * nb(w // hdiv))
        m = N.JdsMetrics()
        if loc == N.JDS_DEVICE:
            import torch
            recon = torch.empty((rows, w, 3), dtype=torch.uint8, device=keep.device)
            coeffs = torch.empty(ncoef, dtype=torch.int16, device=keep.device) if want_coeffs else None
            gp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        else:
            recon = host_array((rows, w, 3), np.uint8)
            coeffs = host_array((ncoef,), np.int16) if want_coeffs else None
            gp = lambda t: C.c_void_p(t.ctypes.data) if t is not None else None
        with self._lock:
            N.check(self._lib.jds_roundtrip_band(self._ctx, C.byref(p), ptr, loc, int(row0), int(row1),
                                                 gp(recon), gp(coeffs), loc, C.byref(m)))
        if loc == N.JDS_DEVICE:
            self._order_torch_after(recon)
        return RoundTripOutputs(recon, coeffs, None, None, m, (h, w))

    # -- GUI plot payload -------------------------------------------------
    def plot_payload(self, image, quality=50, mode="4:2:0", prefilter=False, *,
                     precision="exact", want_heat_rgb=False, want_ssim=True,
                     bins=50) -> "PlotPayload":
        """What the reference's analysis plots draw (gui/compression_tab.py:653-676), reduced
        on the device: the coefficient histogram as matplotlib's ``ax.hist(coeffs, bins)``
        would compute it and the x10 clipped error map as uint8 - about 1 byte per pixel
        back instead of the 3 B/px int16 coefficients and 8 B/px fp64 error maps."""
        h, w, _ = self._frame_geometry(image)
        ptr, loc, keep = self._in_ptr(image)
        flags = N.JDS_OUT_RECON | N.JDS_OUT_PSNR | (N.JDS_OUT_SSIM if want_ssim else 0)
        p = self._params(h, w, quality, mode, prefilter, precision, flags)
        m = N.JdsMetrics()
        vhist = host_array((N.JDS_VALUE_HIST_BINS,), np.int64)
        if loc == N.JDS_DEVICE:
            import torch
            dev = keep.device
            recon = torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
            heat_y = torch.empty((h, w), dtype=torch.uint8, device=dev)
            heat_rgb = torch.empty((h, w), dtype=torch.uint8, device=dev) if want_heat_rgb else None
            d_hist = torch.empty(N.JDS_VALUE_HIST_BINS, dtype=torch.int64, device=dev)
            hist_ptr = C.c_void_p(d_hist.data_ptr())
        else:
            recon = host_array((h, w, 3), np.uint8)
            heat_y = host_array((h, w), np.uint8)
            heat_rgb = host_array((h, w), np.uint8) if want_heat_rgb else None
            d_hist = None
            hist_ptr = C.c_void_p(vhist.ctypes.data)
        gp = lambda t: None if t is None else C.c_void_p(t.data_ptr() if _is_torch(t) else t.ctypes.data)
        with self._lock:
            N.check(self._lib.jds_plot_payload(self._ctx, C.byref(p), ptr, loc, gp(recon), gp(heat_y),
                                               gp(heat_rgb), hist_ptr, loc, C.byref(m)))
        if d_hist is not None:
            vhist = d_hist.cpu().numpy()
        counts, edges = histogram_from_values(vhist, bins)
        return PlotPayload(recon, heat_y, heat_rgb, np.asarray(vhist), counts, edges,
                           RoundTripOutputs(recon, None, None, None, m, (h, w)))

    # -- preview downscale (gui/compression_tab.py:532-552) ----------------
    def resize_area(self, image, out_h: int, out_w: int):
        """cv2.resize(image, (out_w, out_h), interpolation=cv2.INTER_AREA) for a uint8
        H x W x 3 frame being shrunk; output follows the input's location."""
        h, w, _ = self._frame_geometry(image)
        ptr, loc, keep = self._in_ptr(image)
        if loc == N.JDS_DEVICE:
            import torch
            out = torch.empty((out_h, out_w, 3), dtype=torch.uint8, device=keep.device)
            optr = C.c_void_p(out.data_ptr())
        else:
            out = host_array((out_h, out_w, 3), np.uint8)
            optr = C.c_void_p(out.ctypes.data)
        with self._lock:
            N.check(self._lib.jds_resize_area(self._ctx, ptr, loc, h, w, optr, int(out_h), int(out_w), loc))
        return out

    # -- chroma-aliasing demo (gui/dialogs/aliasing_demo_dialog.py:98-166) ----
    def aliasing_demo_arm(self, image, quality=50, prefilter=False, *, precision="exact",
                          want_subsampled=False, want_diff=True) -> dict:
        """One arm of the reference's AliasingDemoWorker: explicit float32 chroma subsampling
        (optionally behind the 5x5 Gaussian), the hot path at 4:4:4, ``compute_metrics``
        against the original and the x10 difference image.  Returns ``recon``, ``diff``,
        ``subsampled`` (uint8 H x W x 3 or None) and ``metrics`` (psnr_y, ssim_y, psnr_rgb,
        ssim_rgb as the reference's ``compute_metrics`` defines them: Y is OpenCV's integer
        luma)."""
        h, w, _ = self._frame_geometry(image)
        ptr, loc, keep = self._in_ptr(image)
        CompressionParams(quality=int(quality))
        if loc == N.JDS_DEVICE:
            import torch
            mk = lambda: torch.empty((h, w, 3), dtype=torch.uint8, device=keep.device)
        else:
            mk = lambda: host_array((h, w, 3), np.uint8)
        recon = mk()
        diff = mk() if want_diff else None
        sub = mk() if want_subsampled else None
        gp = lambda t: None if t is None else C.c_void_p(t.data_ptr() if _is_torch(t) else t.ctypes.data)
        m_rgb, m_luma = N.JdsMetrics(), N.JdsMetrics()
        with self._lock:
            N.check(self._lib.jds_aliasing_demo(self._ctx, ptr, loc, h, w, int(quality), int(bool(prefilter)),
                                                _precision_code(precision), gp(sub), gp(recon), gp(diff),
                                                loc, C.byref(m_rgb), C.byref(m_luma)))
        metrics = self._aliasing_metrics_dict(m_rgb, m_luma, h, w)
        return {'recon': recon, 'diff': diff, 'subsampled': sub, 'metrics': metrics}

    @staticmethod
    def _aliasing_metrics_dict(m_rgb, m_luma, h, w) -> dict:
        from .utils.metrics import psnr_from_sse
        n_px = h * w
        cnt = m_rgb.ssim_count
        return {
            'psnr_y': psnr_from_sse(m_luma.sse_rgb // 3, n_px),
            'ssim_y': float(np.float64(m_luma.ssim_sum[0]) / np.float64(cnt)) if cnt else float('nan'),
            'psnr_rgb': psnr_from_sse(m_rgb.sse_rgb, 3 * n_px),
            'ssim_rgb': float((np.array([m_rgb.ssim_sum[0], m_rgb.ssim_sum[1], m_rgb.ssim_sum[2]],
                                        dtype=np.float64) / np.float64(cnt)).mean()) if cnt else float('nan'),
        }

    def aliasing_metrics(self, original, reconstructed) -> dict:
        """The aliasing demo's ``compute_metrics`` (gui/dialogs/aliasing_demo_dialog.py:69-83)."""
        h, w, _ = self._frame_geometry(original)
        if tuple(original.shape) != tuple(reconstructed.shape):
            raise ValueError('Input images must have the same dimensions.')
        pa, loc, ka = self._in_ptr(original)
        pb, loc_b, kb = self._in_ptr(reconstructed)
        if loc != loc_b:
            raise ValueError('both frames must live on the same side (host or device)')
        m_rgb, m_luma = N.JdsMetrics(), N.JdsMetrics()
        with self._lock:
            N.check(self._lib.jds_aliasing_metrics(self._ctx, pa, pb, loc, h, w, C.byref(m_rgb), C.byref(m_luma)))
        return self._aliasing_metrics_dict(m_rgb, m_luma, h, w)

    # -- entropy-coded size (second half) ------------------------------------
    def entropy_bits(self, coeffs, height: int, width: int, mode="4:2:0") -> List[int]:
        """Exact bits of the three baseline-JPEG scans (Y, Cb, Cr; Annex K Huffman tables) of
        an ``all_quantized_coeffs`` array (int16 NumPy array or torch tensor, host or CUDA)."""
        sub = _mode_code(mode)
        ptr, loc, keep, _ = self._coeff_ptr(coeffs, height, width, sub, mode)
        out = (C.c_uint64 * 3)()
        with self._lock:
            N.check(self._lib.jds_entropy_bits(self._ctx, ptr, loc, int(height), int(width), sub, out))
        return [int(out[0]), int(out[1]), int(out[2])]

    def _coeff_ptr(self, coeffs, height, width, sub, mode):
        n = C.c_uint64()
        N.check(self._lib.jds_coeff_count(int(height), int(width), sub, C.byref(n)))
        if _is_torch(coeffs):
            import torch
            if coeffs.dtype != torch.int16:
                raise TypeError(f"coeffs must be int16, got {coeffs.dtype}")
            t = coeffs.contiguous()
            ptr, loc, keep, size = C.c_void_p(t.data_ptr()), (N.JDS_DEVICE if t.is_cuda else N.JDS_HOST), t, t.numel()
            if t.is_cuda:
                self._order_after_torch(t)
        else:
            a = np.ascontiguousarray(coeffs)
            if a.dtype != np.int16:
                raise TypeError(f"coeffs must be int16, got {a.dtype}")
            ptr, loc, keep, size = C.c_void_p(a.ctypes.data), N.JDS_HOST, a, a.size
        if size != n.value:
            raise ValueError(f"expected {n.value} coefficients for {height}x{width} {mode}, got {size}")
        return ptr, loc, keep, int(n.value)

    # -- entropy-coded bytes (the bitstream itself, coded on the device) --------
    def entropy_encode(self, coeffs, height: int, width: int, mode="4:2:0"):
        """The three baseline-JPEG scans (Y, Cb, Cr) of an ``all_quantized_coeffs`` array as
        entropy-coded bytes (zig-zag of utils/constants.py:18-27, Annex K Huffman tables, byte
        stuffing and padding included): ``([bytes_Y, bytes_Cb, bytes_Cr], [bits_Y, bits_Cb, bits_Cr])``."""
        sub = _mode_code(mode)
        ptr, loc, keep, ncoef = self._coeff_ptr(coeffs, height, width, sub, mode)
        nbytes, nbits = (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
        cap = 2 * ncoef           # the raw int16 array; real scans are far smaller
        with self._lock:
            while True:
                out = np.empty(cap, dtype=np.uint8)
                rc = self._lib.jds_entropy_encode(self._ctx, ptr, loc, int(height), int(width), sub,
                                                  C.c_void_p(out.ctypes.data), N.JDS_HOST, cap, nbytes, nbits)
                if rc != N.JDS_ERR_CAPACITY:
                    N.check(rc)
                    break
                cap = int(sum(nbytes))
        ends = np.cumsum([int(b) for b in nbytes])
        scans = [out[e - int(b):e].tobytes() for e, b in zip(ends, nbytes)]
        return scans, [int(b) for b in nbits]

    def jfif_encode(self, coeffs, height: int, width: int, mode="4:2:0", quality: int = 50, qtable=None):
        """A complete baseline JFIF file (bytes) of a round trip's coefficients, entropy-coded on
        the device; ``qtable`` (8x8, integers 1..255) defaults to the round trip's own table for
        ``quality`` (engines/quantizer.py:7-19).  Returns ``(file_bytes, scan_bits)``."""
        sub = _mode_code(mode)
        ptr, loc, keep, ncoef = self._coeff_ptr(coeffs, height, width, sub, mode)
        q = (C.c_double * 64)()
        if qtable is None:
            N.check(self._lib.jds_quant_table(int(quality), q))
        else:
            flat = np.asarray(qtable, dtype=np.float64).reshape(-1)
            if flat.size != 64:
                raise ValueError("qtable must have 64 entries")
            q[:] = flat.tolist()
        nbytes, nbits = C.c_uint64(), (C.c_uint64 * 3)()
        cap = 2 * ncoef + 1024
        with self._lock:
            while True:
                out = np.empty(cap, dtype=np.uint8)
                rc = self._lib.jds_jfif_encode(self._ctx, ptr, loc, int(height), int(width), sub, q,
                                               C.c_void_p(out.ctypes.data), cap, C.byref(nbytes), nbits)
                if rc != N.JDS_ERR_CAPACITY:
                    N.check(rc)
                    break
                cap = int(nbytes.value)
        return out[:int(nbytes.value)].tobytes(), [int(b) for b in nbits]

    def selected_block(self, image, quality, block_row, block_col):
        """IntermediateData.selected_block_* (engines/pipeline.py:126-151) or None."""
        h, w, _ = self._frame_geometry(image)
        ptr, loc, keep = self._in_ptr(image)
        p = self._params(h, w, quality, "4:4:4", False, "exact", 0)
        arrs = {k: np.empty((8, 8), dtype=np.float64)
                for k in ("original", "shifted", "dct", "dequantized", "reconstructed")}
        arrs["quantized"] = np.empty((8, 8), dtype=np.int16)
        present = C.c_int(0)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        with self._lock:
            N.check(self._lib.jds_selected_block(
                self._ctx, C.byref(p), ptr, loc, int(block_row), int(block_col),
                vp(arrs["original"]), vp(arrs["shifted"]), vp(arrs["dct"]), vp(arrs["quantized"]),
                vp(arrs["dequantized"]), vp(arrs["reconstructed"]), C.byref(present)))
        return arrs if present.value else None

    # -- batch (BASELINE config 5) -----------------------------------------------------
    def roundtrip_batch(self, frames, quality=50, mode="4:2:0", prefilter=False, *,
                        precision="fast", want_recon=True, want_coeffs=False,
                        want_hist=False, want_ssim=True, recon_out=None) -> List[RoundTripOutputs]:
        """N frames of one geometry: ``frames`` is uint8 N x H x W x 3 (NumPy, pinned
        or pageable torch host tensor, or CUDA tensor)."""
        shape = tuple(frames.shape)
        if len(shape) != 4 or shape[-1] != 3:
            raise ValueError(f"expected N x H x W x 3 frames, got shape {shape}")
        n, h, w, _ = shape
        ptr, loc, keep = self._in_ptr(frames)
        flags = N.JDS_OUT_PSNR | (N.JDS_OUT_RECON if want_recon else 0)
        flags |= N.JDS_OUT_COEFFS if want_coeffs else 0
        flags |= N.JDS_OUT_HIST if want_hist else 0
        flags |= N.JDS_OUT_SSIM if want_ssim else 0
        p = self._params(h, w, quality, mode, prefilter, precision, flags)
        ncoef = C.c_uint64()
        N.check(self._lib.jds_coeff_count(h, w, p.subsampling, C.byref(ncoef)))
        ms = (N.JdsMetrics * n)()
        recon = coeffs = None
        if loc == N.JDS_DEVICE:
            import torch
            if want_recon:
                recon = recon_out if recon_out is not None else torch.empty((n, h, w, 3), dtype=torch.uint8, device=keep.device)
            if want_coeffs:
                coeffs = torch.empty((n, ncoef.value), dtype=torch.int16, device=keep.device)
        else:
            if want_recon:
                recon = recon_out if recon_out is not None else np.empty((n, h, w, 3), dtype=np.uint8)
            if want_coeffs:
                coeffs = np.empty((n, ncoef.value), dtype=np.int16)

        def gp(t):
            if t is None:
                return None
            return C.c_void_p(t.data_ptr() if _is_torch(t) else t.ctypes.data)
        with self._lock:
            N.check(self._lib.jds_roundtrip_batch(self._ctx, C.byref(p), n, ptr, loc, gp(recon),
                                                  gp(coeffs), loc, ms))
        return [RoundTripOutputs(recon[i] if recon is not None else None,
                                 coeffs[i] if coeffs is not None else None, None, None, ms[i],
                                 (h, w), ms) for i in range(n)]

    def roundtrip_batch_begin(self, frames, quality=50, mode="4:2:0", prefilter=False, *,
                              precision="fast", want_recon=True, want_ssim=True, recon_out=None):
        """``roundtrip_batch`` without the wait: enqueues the batch (copies included) and returns
        a ``PendingBatch``; ``.result()`` waits and returns the list of ``RoundTripOutputs``.
        One pending batch per engine - alternate two engines to overlap consecutive batches
        (``engines.pipeline.compress_stream`` does)."""
        shape = tuple(frames.shape)
        if len(shape) != 4 or shape[-1] != 3:
            raise ValueError(f"expected N x H x W x 3 frames, got shape {shape}")
        n, h, w, _ = shape
        ptr, loc, keep = self._in_ptr(frames)
        flags = N.JDS_OUT_PSNR | (N.JDS_OUT_RECON if want_recon else 0) | (N.JDS_OUT_SSIM if want_ssim else 0)
        p = self._params(h, w, quality, mode, prefilter, precision, flags)
        ms = (N.JdsMetrics * n)()
        recon = None
        if want_recon:
            if recon_out is not None:
                recon = recon_out
            elif loc == N.JDS_DEVICE:
                import torch
                recon = torch.empty((n, h, w, 3), dtype=torch.uint8, device=keep.device)
            else:
                recon = host_array((n, h, w, 3), np.uint8)
        rp = None if recon is None else C.c_void_p(recon.data_ptr() if _is_torch(recon) else recon.ctypes.data)
        with self._lock:
            N.check(self._lib.jds_roundtrip_batch_begin(self._ctx, C.byref(p), n, ptr, loc, rp, None, loc, ms))
        return PendingBatch(self, keep, recon, ms, (h, w), loc)

    def _finish(self):
        with self._lock:
            N.check(self._lib.jds_ctx_finish(self._ctx))

    # -- quality sweep (BASELINE config 4; gui/worker.py:55-74) -------------------------
    def sweep(self, image, qualities: Sequence[int], mode="4:2:0", prefilter=False, *,
              precision="fast", want_recon=False, want_ssim=True) -> List[RoundTripOutputs]:
        h, w, _ = self._frame_geometry(image)
        ptr, loc, keep = self._in_ptr(image)
        qs = [int(q) for q in qualities]
        _check_qualities(qs)
        nq = len(qs)
        if nq == 0:
            return []
        flags = N.JDS_OUT_PSNR | (N.JDS_OUT_RECON if want_recon else 0)
        flags |= N.JDS_OUT_SSIM if want_ssim else 0
        p = self._params(h, w, 50, mode, prefilter, precision, flags)
        ms = (N.JdsMetrics * nq)()
        qarr = (C.c_int32 * nq)(*qs)
        recon = None
        if want_recon:
            if loc == N.JDS_DEVICE:
                import torch
                recon = torch.empty((nq, h, w, 3), dtype=torch.uint8, device=keep.device)
            else:
                recon = np.empty((nq, h, w, 3), dtype=np.uint8)
        rp = None if recon is None else C.c_void_p(recon.data_ptr() if _is_torch(recon) else recon.ctypes.data)
        with self._lock:
            N.check(self._lib.jds_sweep(self._ctx, C.byref(p), qarr, nq, ptr, loc, rp, loc, ms))
        return [RoundTripOutputs(recon[i] if recon is not None else None, None, None, None, ms[i],
                                 (h, w), ms) for i in range(nq)]

    def sweep_records(self, image, qualities: Sequence[int], records, *, mode="4:2:0",
                      prefilter=False, precision="fast", unit0=0, unit_step=1, want_ssim=True):
        """Sweep whose metric records stay on the device (``jds_sweep_records``): one row of
        ``N.JDS_RECORD_FIELDS`` fp64 per quality is written to ``records`` - a CUDA fp64
        tensor of shape (capacity, 13) - and rows past ``len(qualities)`` are marked empty
        (unit -1).  Returns WITHOUT synchronising: the work is ordered on this engine's
        stream (``use_stream``); the caller enqueues its collective / copy on the same
        stream and synchronises once.  At most ``N.JDS_SWEEP_RECORDS_MAX`` rows per call."""
        h, w, _ = self._frame_geometry(image)
        ptr, loc, keep = self._in_ptr(image)
        qs = [int(q) for q in qualities]
        _check_qualities(qs)
        if not _is_torch(records) or not records.is_cuda or not records.is_contiguous():
            raise TypeError("records must be a contiguous CUDA fp64 tensor")
        cap = int(records.shape[0])
        if tuple(records.shape[1:]) != (N.JDS_RECORD_FIELDS,) or records.element_size() != 8:
            raise ValueError(f"records must have shape (capacity, {N.JDS_RECORD_FIELDS}) fp64")
        flags = N.JDS_OUT_PSNR | (N.JDS_OUT_SSIM if want_ssim else 0)
        p = self._params(h, w, 50, mode, prefilter, precision, flags)
        qarr = (C.c_int32 * max(len(qs), 1))(*qs)
        with self._lock:
            N.check(self._lib.jds_sweep_records(self._ctx, C.byref(p), qarr, len(qs), ptr, loc,
                                                int(unit0), int(unit_step),
                                                C.c_void_p(records.data_ptr()), cap))
        self._order_torch_after(records)
        return keep

    def batch_records(self, frames, records, quality=50, mode="4:2:0", prefilter=False, *,
                      precision="fast", recon_out=None, unit0=0, unit_step=1, want_ssim=True):
        """``roundtrip_batch`` for CUDA tensors whose metric records stay on the device
        (``jds_roundtrip_batch_records``; rows as in ``distributed.RECORD_FIELDS``).  Returns
        WITHOUT synchronising - see ``sweep_records``."""
        shape = tuple(frames.shape)
        if len(shape) != 4 or shape[-1] != 3:
            raise ValueError(f"expected N x H x W x 3 frames, got shape {shape}")
        n, h, w, _ = shape
        ptr, loc, keep = self._in_ptr(frames)
        if loc != N.JDS_DEVICE:
            raise TypeError("frames must be a CUDA uint8 tensor")
        if not _is_torch(records) or not records.is_cuda or not records.is_contiguous():
            raise TypeError("records must be a contiguous CUDA fp64 tensor")
        if tuple(records.shape[1:]) != (N.JDS_RECORD_FIELDS,) or records.element_size() != 8:
            raise ValueError(f"records must have shape (capacity, {N.JDS_RECORD_FIELDS}) fp64")
        CompressionParams(quality=int(quality))
        flags = N.JDS_OUT_PSNR | (N.JDS_OUT_SSIM if want_ssim else 0) | (N.JDS_OUT_RECON if recon_out is not None else 0)
        p = self._params(h, w, quality, mode, prefilter, precision, flags)
        rp = None
        if recon_out is not None:
            if not recon_out.is_cuda or tuple(recon_out.shape) != shape:
                raise ValueError("recon_out must be a CUDA uint8 tensor of the frames' shape")
            rp = C.c_void_p(recon_out.data_ptr())
        with self._lock:
            N.check(self._lib.jds_roundtrip_batch_records(self._ctx, C.byref(p), n, ptr, rp, int(unit0),
                                                          int(unit_step), C.c_void_p(records.data_ptr()),
                                                          int(records.shape[0])))
        self._order_torch_after(records)
        return keep


_engines = {}
_engines_lock = threading.Lock()


def stream_engines(device: Optional[int] = None, n: int = 2) -> List["Engine"]:
    """``n`` engines on one device for callers that keep several batches in flight: the
    process-wide engine plus ``n - 1`` more contexts (created on first use, kept)."""
    first = get_engine(device)
    with _engines_lock:
        extra = _extra_engines.setdefault(first.device, [])
        while len(extra) < n - 1:
            extra.append(Engine(first.device))
        return [first] + extra[:n - 1]


_extra_engines = {}


def get_engine(device: Optional[int] = None) -> Engine:
    """Process-wide engine per device (created on first use)."""
    if device is None:
        device = 0
        try:
            import os
            device = int(os.environ.get("JDS_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        except ValueError:
            device = 0
    with _engines_lock:
        eng = _engines.get(device)
        if eng is None:
            eng = _engines[device] = Engine(device)
        return eng
