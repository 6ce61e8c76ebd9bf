"""Command-line entry: the reference's ``main.py --cli`` (main.py:46-91) on the GPU path.

    python -m jpeg_dsp_studio_b200 [--cli] [image] [--quality N] [--mode 4:2:0] [--prefilter]
                                   [--fast] [--sweep START:END:STEP]

Without an image a 256x256 coloured checkerboard is generated, exactly like the reference.
Reading / writing image files uses OpenCV when it is installed (file I/O is outside the
accelerated path); without it the reconstruction is saved as .npy.
"""

import argparse
import sys

import numpy as np

from . import CompressionParams, compress_reconstruct, quality_sweep
from .utils.test_images import generate_colored_checkerboard


def _load(path):
    try:
        import cv2
    except ImportError:
        raise SystemExit("reading image files needs OpenCV (cv2); pass no image for the demo board")
    bgr = cv2.imread(path, cv2.IMREAD_COLOR)
    if bgr is None:
        raise SystemExit(f"Could not load image: {path}")
    return np.ascontiguousarray(bgr[:, :, ::-1])


def _save(path, rgb):
    try:
        import cv2
        cv2.imwrite(path, np.ascontiguousarray(rgb[:, :, ::-1]))
        return path
    except ImportError:
        np.save(path + ".npy", rgb)
        return path + ".npy"


def main(argv=None):
    ap = argparse.ArgumentParser(prog="jpeg_dsp_studio_b200", description=__doc__.splitlines()[0])
    ap.add_argument("--cli", action="store_true", help="accepted for compatibility with main.py")
    ap.add_argument("image", nargs="?", help="image file (default: 256x256 checkerboard)")
    ap.add_argument("--quality", type=int, default=50)
    ap.add_argument("--mode", default="4:2:0", choices=["4:4:4", "4:2:2", "4:2:0"])
    ap.add_argument("--prefilter", action="store_true")
    ap.add_argument("--fast", action="store_true", help="fp32 mode (default: bit-exact fp64 mode)")
    ap.add_argument("--sweep", default=None, metavar="START:END:STEP",
                    help="rate-distortion sweep instead of a single run (BatchSweepWorker)")
    ap.add_argument("--output", default="reconstructed.png")
    ap.add_argument("--jpeg", default=None, metavar="FILE",
                    help="also write the coefficients as a baseline JPEG file (entropy-coded on the GPU)")
    args = ap.parse_args(argv)

    print("JPEG-DSP Studio - CLI Mode (B200 path)")
    print("=" * 40)
    if args.image:
        image = _load(args.image)
        print(f"Loaded: {args.image} ({image.shape[1]}x{image.shape[0]})")
    else:
        image = generate_colored_checkerboard(256)
        print("Using test checkerboard 256x256")
    params = CompressionParams(quality=args.quality, block_size=8, subsampling_mode=args.mode,
                               use_prefilter=args.prefilter)
    precision = "fast" if args.fast else "exact"
    if args.sweep:
        a, b, st = (int(v) for v in args.sweep.split(":"))
        print(f"Sweep Q={a}..{b} step {st}, {args.mode}, prefilter {args.prefilter}, {precision} mode")
        print(" Q     bpp   PSNR(Y)  SSIM(Y)  PSNR(RGB) SSIM(RGB)  ratio")
        for q, r in quality_sweep(image, params, range(a, b + 1, st), precision=precision):
            print(f"{q:3d} {r.bpp:7.3f} {r.psnr_y:8.2f} {r.ssim_y:8.4f} {r.psnr_rgb:9.2f} "
                  f"{r.ssim_rgb:9.4f} {r.compression_ratio:6.2f}x")
        return 0
    result, inter = compress_reconstruct(image, params, precision=precision)
    print(f"PSNR (Y): {result.psnr_y:.2f} dB")
    print(f"SSIM (Y): {result.ssim_y:.4f}")
    print(f"BPP: {result.bpp:.3f}")
    print(f"Compression Ratio: {result.compression_ratio:.2f}x")
    print(f"Runtime: {result.encode_time_ms + result.decode_time_ms:.1f} ms")
    print(f"Saved: {_save(args.output, result.reconstructed_image)}")
    if args.jpeg:
        from .utils.metrics import encode_jfif
        data = encode_jfif(inter.all_quantized_coeffs, image.shape[:2], args.mode, args.quality)
        with open(args.jpeg, "wb") as f:
            f.write(data)
        print(f"JPEG: {args.jpeg} ({len(data)} bytes, {8 * len(data) / (image.shape[0] * image.shape[1]):.3f} bpp "
              "with entropy coding)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
