"""Drop-in for the reference's ``engines/pipeline.py``: the compression round trip.

``compress_reconstruct(image_rgb, params, selected_block_idx=(0, 0))`` has the
reference's signature, return types and error behaviour
(engines/pipeline.py:17-21, 102-167), but every stage - RGB->YCbCr, prefilter and
chroma decimation, 8x8 DCT, quantise/dequantise, IDCT, chroma upsampling,
YCbCr->RGB with clamp and truncation, PSNR/SSIM/bit-count/histogram reductions -
runs in the CUDA kernels of libjds.so.  Keyword-only extras (not in the reference)
select the arithmetic mode and the device.
"""

import time
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from ..engine import get_engine
from ..models.compression_params import CompressionParams
from ..models.compression_result import CompressionResult
from ..models.intermediate_data import IntermediateData
from ..utils.metrics import BITRATE_LABEL


def _validate(image_rgb, params: CompressionParams):
    """Raise what the reference raises, where it can be known before launching."""
    if not hasattr(image_rgb, "shape"):
        image_rgb = np.asarray(image_rgb)
    if image_rgb.ndim < 3:
        # reference: rgb[:, :, 0] on a 2-D array (engines/color_space.py:10)
        raise IndexError(f"too many indices for array: array is {image_rgb.ndim}-dimensional, "
                         "but 3 were indexed")
    if image_rgb.shape[2] != 3:
        # reference: np.stack of 3 planes vs image_float with C channels does not
        # broadcast (engines/pipeline.py:121) / skimage channel mismatch
        raise ValueError(f"operands could not be broadcast together with shapes "
                         f"{tuple(image_rgb.shape)} {tuple(image_rgb.shape[:2]) + (3,)}")
    if params.block_size != 8:
        # reference: (B,B) DCT block divided by the (8,8) table (engines/quantizer.py:24)
        b = params.block_size
        raise ValueError(f"operands could not be broadcast together with shapes ({b},{b}) (8,8) ")
    if params.subsampling_mode not in ('4:4:4', '4:2:2', '4:2:0'):
        raise ValueError(f"Unknown subsampling mode: {params.subsampling_mode}")
    if min(image_rgb.shape[:2]) < 7:
        # skimage.metrics.structural_similarity on an image smaller than its 7x7 window
        raise ValueError(
            "win_size exceeds image extent. Either ensure that your images are at least "
            "7x7; or pass win_size explicitly in the function call, with an odd value "
            "less than or equal to the smaller side of your images.")
    return image_rgb


def compress_reconstruct(
    image_rgb: np.ndarray,
    params: CompressionParams,
    selected_block_idx: Tuple[int, int] = (0, 0),
    *,
    precision: str = "exact",
    device: Optional[int] = None,
    intermediates: bool = True,
) -> Tuple[CompressionResult, IntermediateData]:
    """Run the full JPEG-like compression and reconstruction round trip on the GPU.

    ``precision='exact'`` (default) reproduces the reference's int16 coefficients and
    uint8 pixels bit for bit; ``'fast'`` is the fp32 mode.  ``intermediates=False``
    skips the bulky IntermediateData arrays (coefficients, error maps, histogram)
    that the GUI plots - the returned IntermediateData then only echoes the index.
    """
    t_start = time.perf_counter()
    image_checked = _validate(image_rgb, params)
    eng = get_engine(device)
    src = image_checked
    if isinstance(src, np.ndarray) and src.dtype != np.uint8:
        # the reference casts whatever it gets to float64 and its final uint8 cast
        # implies 8-bit data; non-uint8 input is outside the path's contract
        raise TypeError(f"image_rgb must be uint8, got {src.dtype}")

    out = eng.roundtrip(src, params.quality, params.subsampling_mode, params.use_prefilter,
                        precision=precision, want_coeffs=intermediates,
                        want_error_maps=intermediates, want_hist=intermediates)
    s = out.scalars
    gpu_ms = float(out.metrics.gpu_ms)

    block_row, block_col = selected_block_idx
    sel = eng.selected_block(src, params.quality, block_row, block_col)

    wall_ms = (time.perf_counter() - t_start) * 1000.0
    result = CompressionResult(
        original_image=image_rgb,
        reconstructed_image=out.recon,
        psnr_y=s['psnr_y'],
        ssim_y=s['ssim_y'],
        psnr_rgb=s['psnr_rgb'],
        ssim_rgb=s['ssim_rgb'],
        bpp=s['bpp'],
        compression_ratio=s['compression_ratio'],
        nonzero_coeffs=s['nonzero_count'],
        total_coeffs=s['total_coeffs'],
        # the GUI shows encode+decode (gui/compression_tab.py:626): here the sum is the
        # wall time of this call, split into device kernel time and everything else
        encode_time_ms=gpu_ms,
        decode_time_ms=max(wall_ms - gpu_ms, 0.0),
        bitrate_label=BITRATE_LABEL,
    )
    hist = None
    if intermediates:
        hist = np.array(list(out.metrics.hist50), dtype=np.int64)
    intermediate = IntermediateData(
        selected_block_idx=selected_block_idx,
        selected_block_original=sel['original'] if sel else None,
        selected_block_shifted=sel['shifted'] if sel else None,
        selected_block_dct=sel['dct'] if sel else None,
        selected_block_quantized=sel['quantized'] if sel else None,
        selected_block_dequantized=sel['dequantized'] if sel else None,
        selected_block_reconstructed=sel['reconstructed'] if sel else None,
        error_map_y=out.err_y,
        error_map_rgb=out.err_rgb,
        quantized_histogram=hist,
        all_quantized_coeffs=out.coeffs,
    )
    return result, intermediate


def quality_sweep(image_rgb, base_params: CompressionParams,
                  qualities: Sequence[int] = tuple(range(10, 91, 10)), *,
                  precision: str = "fast", device: Optional[int] = None,
                  keep_images: bool = False,
                  progress: Optional[Callable[[int, int], None]] = None,
                  progress_points: int = 0):
    """Rate-distortion sweep: the loop of ``BatchSweepWorker.run`` (gui/worker.py:55-74)
    as one call.  Returns ``[(quality, CompressionResult), ...]`` like the worker's
    ``finished`` payload; ``reconstructed_image`` is ``None`` unless ``keep_images``
    (the sweep's consumers read only bpp/PSNR/SSIM: gui/compression_tab.py:752-754,
    gui/main_window.py:286-293).

    ``progress(done, total)`` mirrors the worker's ``progress`` signal (gui/worker.py:44,70):
    it is called once per finished quality point, in order.  The points are then computed in
    groups of ``progress_points`` (default: a tenth of the sweep, at least one) so that a GUI
    sees movement while the device works; without a callback the sweep is one native call."""
    image_checked = _validate(image_rgb, base_params)
    eng = get_engine(device)
    qualities = list(qualities)
    t0 = time.perf_counter()
    if progress is None or not qualities:
        outs = eng.sweep(image_checked, qualities, base_params.subsampling_mode,
                         base_params.use_prefilter, precision=precision, want_recon=keep_images)
    else:
        for q in qualities:
            CompressionParams(quality=int(q))       # fail before the first group, like one call would
        total = len(qualities)
        group = int(progress_points) if progress_points and progress_points > 0 else max(1, -(-total // 10))
        outs = []
        for g0 in range(0, total, group):
            part = eng.sweep(image_checked, qualities[g0:g0 + group], base_params.subsampling_mode,
                             base_params.use_prefilter, precision=precision, want_recon=keep_images)
            outs.extend(part)
            for i in range(g0, g0 + len(part)):
                progress(i + 1, total)
    wall_ms = (time.perf_counter() - t0) * 1000.0
    results = []
    for q, o in zip(qualities, outs):
        s = o.scalars
        results.append((int(q), CompressionResult(
            original_image=image_rgb, reconstructed_image=o.recon,
            psnr_y=s['psnr_y'], ssim_y=s['ssim_y'], psnr_rgb=s['psnr_rgb'],
            ssim_rgb=s['ssim_rgb'], bpp=s['bpp'], compression_ratio=s['compression_ratio'],
            nonzero_coeffs=s['nonzero_count'], total_coeffs=s['total_coeffs'],
            encode_time_ms=float(o.metrics.gpu_ms),
            decode_time_ms=max(wall_ms / max(len(outs), 1) - float(o.metrics.gpu_ms), 0.0),
            bitrate_label=BITRATE_LABEL)))
    return results


def compress_batch(frames, params: CompressionParams, *, precision: str = "fast",
                   device: Optional[int] = None, keep_images: bool = True):
    """A batch of same-sized frames (N x H x W x 3 uint8) through one call; returns a
    list of CompressionResult (BASELINE.json config 5)."""
    if frames.ndim != 4:
        raise ValueError(f"expected N x H x W x 3 frames, got shape {tuple(frames.shape)}")
    _validate(frames[0], params)
    eng = get_engine(device)
    t0 = time.perf_counter()
    outs = eng.roundtrip_batch(frames, params.quality, params.subsampling_mode,
                               params.use_prefilter, precision=precision,
                               want_recon=keep_images)
    wall_ms = (time.perf_counter() - t0) * 1000.0
    results = []
    for i, o in enumerate(outs):
        s = o.scalars
        results.append(CompressionResult(
            original_image=frames[i], reconstructed_image=o.recon,
            psnr_y=s['psnr_y'], ssim_y=s['ssim_y'], psnr_rgb=s['psnr_rgb'],
            ssim_rgb=s['ssim_rgb'], bpp=s['bpp'], compression_ratio=s['compression_ratio'],
            nonzero_coeffs=s['nonzero_count'], total_coeffs=s['total_coeffs'],
            encode_time_ms=float(o.metrics.gpu_ms),
            decode_time_ms=max(wall_ms / max(len(outs), 1) - float(o.metrics.gpu_ms), 0.0),
            bitrate_label=BITRATE_LABEL))
    return results


def compress_stream(batches, params: CompressionParams, *, precision: str = "fast",
                    device: Optional[int] = None, keep_images: bool = True, recon_out=None):
    """A stream of same-sized batches (an iterable of N x H x W x 3 uint8 arrays - a folder of
    frames, a video): yields ``compress_batch``'s list of results for every batch, in order,
    with consecutive batches overlapped - batch k+1 is enqueued on a second context before the
    results of batch k are awaited, so its host->device copies run while batch k's last
    reconstructions are still on their way back.  ``recon_out``: optional pair of N x H x W x 3
    uint8 host buffers to receive the reconstructions alternately (pinned memory makes the
    copies asynchronous); the arrays of a yielded batch are valid until two batches later."""
    from ..engine import stream_engines
    engines = stream_engines(device, 2)

    def finish(pending, frames, t0):
        outs = pending.result()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        results = []
        for i, o in enumerate(outs):
            s = o.scalars
            results.append(CompressionResult(
                original_image=frames[i], reconstructed_image=o.recon,
                psnr_y=s['psnr_y'], ssim_y=s['ssim_y'], psnr_rgb=s['psnr_rgb'],
                ssim_rgb=s['ssim_rgb'], bpp=s['bpp'], compression_ratio=s['compression_ratio'],
                nonzero_coeffs=s['nonzero_count'], total_coeffs=s['total_coeffs'],
                encode_time_ms=float(o.metrics.gpu_ms),
                decode_time_ms=max(wall_ms / max(len(outs), 1) - float(o.metrics.gpu_ms), 0.0),
                bitrate_label=BITRATE_LABEL))
        return results

    prev = None
    try:
        for k, frames in enumerate(batches):
            if frames.ndim != 4:
                raise ValueError(f"expected N x H x W x 3 frames, got shape {tuple(frames.shape)}")
            _validate(frames[0], params)
            t0 = time.perf_counter()
            pending = engines[k & 1].roundtrip_batch_begin(
                frames, params.quality, params.subsampling_mode, params.use_prefilter, precision=precision,
                want_recon=keep_images, recon_out=None if recon_out is None else recon_out[k & 1])
            done, prev = prev, (pending, frames, t0)
            if done is not None:
                yield finish(*done)
        done, prev = prev, None
        if done is not None:
            yield finish(*done)
    finally:
        # the consumer stopped early or a batch was rejected: collect the batch still in flight so
        # that its context accepts calls again
        if prev is not None:
            try:
                prev[0].result()
            except Exception:
                pass


def plot_payload(image_rgb, params: CompressionParams, *, precision: str = "exact",
                 device: Optional[int] = None, want_heat_rgb: bool = False, bins: int = 50):
    """The round trip for the GUI's analysis plots (gui/compression_tab.py:653-676) without
    the bulky intermediates: returns ``(CompressionResult, PlotPayload)`` where the payload
    holds the 50-bin coefficient histogram exactly as ``ax.hist(all_quantized_coeffs, 50)``
    computes it (gui/widgets/mpl_canvas.py:96) and ``clip(error_map_y * 10, 0, 255)``
    (mpl_canvas.py:118) as uint8 - bit-identical to reducing the reference's arrays in
    ``precision="exact"``."""
    image_rgb = _validate(image_rgb, params)
    eng = get_engine(device)
    t0 = time.perf_counter()
    pay = eng.plot_payload(image_rgb, params.quality, params.subsampling_mode, params.use_prefilter,
                           precision=precision, want_heat_rgb=want_heat_rgb, bins=bins)
    wall_ms = (time.perf_counter() - t0) * 1000.0
    s = pay.scalars
    gpu_ms = float(pay.outputs.metrics.gpu_ms)
    result = CompressionResult(
        original_image=image_rgb, reconstructed_image=pay.reconstructed_image,
        psnr_y=s['psnr_y'], ssim_y=s['ssim_y'], psnr_rgb=s['psnr_rgb'], ssim_rgb=s['ssim_rgb'],
        bpp=s['bpp'], compression_ratio=s['compression_ratio'],
        nonzero_coeffs=s['nonzero_count'], total_coeffs=s['total_coeffs'],
        encode_time_ms=gpu_ms, decode_time_ms=max(wall_ms - gpu_ms, 0.0),
        bitrate_label=BITRATE_LABEL)
    return result, pay
