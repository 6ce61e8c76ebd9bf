"""World-size-2 gloo test of the sharding / gathering logic (no GPU): the engine is a
stand-in that answers from the oracle, the collectives are the real torch.distributed
calls the GPU path makes over NCCL."""

import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class OracleEngine:
    """Duck-typed Engine: metric partials computed by the oracle on the host."""

    @staticmethod
    def _metrics(img, q, mode, pf):
        from oracle import numpy_port as P
        o = P.compress_reconstruct(img, q, mode, pf, want_maps=False, want_metrics=False)
        rec = o["reconstructed_image"].astype(np.float64)
        a = img.astype(np.float64)
        ya = 0.299 * a[..., 0] + 0.587 * a[..., 1] + 0.114 * a[..., 2]
        yb = 0.299 * rec[..., 0] + 0.587 * rec[..., 1] + 0.114 * rec[..., 2]
        h, w = img.shape[:2]
        nnz, bits = P.bit_length_sum(o["all_quantized_coeffs"])
        m = SimpleNamespace(
            sse_rgb=int(np.sum((a - rec) ** 2)), sse_y=float(np.sum((ya - yb) ** 2)),
            ssim_sum=[1.0 * (h - 6) * (w - 6) * 0.9] * 4, ssim_count=(h - 6) * (w - 6),
            coeff_bits=bits, nnz=nnz, total_coeffs=o["all_quantized_coeffs"].size,
            luma_blocks=(-(-h // 8)) * (-(-w // 8)))
        return SimpleNamespace(metrics=m)

    def sweep(self, image, qualities, mode, prefilter, precision="fast"):
        return [self._metrics(image, q, mode, prefilter) for q in qualities]

    def roundtrip_batch(self, frames, quality, mode, prefilter, precision="fast", want_recon=False):
        return [self._metrics(f, quality, mode, prefilter) for f in frames]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    from jpeg_dsp_studio_b200 import distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        img = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
        qs = [5, 20, 35, 50, 65, 80, 95]
        table = D.sweep_sharded(OracleEngine(), img, qs, "4:2:0", False)
        frames = np.stack([np.random.default_rng(100 + k).integers(0, 256, (32, 48, 3), dtype=np.uint8)
                           for k in range(5)])
        mine = frames[D.shard_indices(5, rank, world)]
        agg = D.batch_sharded(OracleEngine(), mine, 5, 40, "4:2:2", False)
        q.put((rank, table, agg))
    finally:
        dist.destroy_process_group()


def test_shard_indices_cover_units():
    from jpeg_dsp_studio_b200 import distributed as D
    for n in (0, 1, 7, 100, 1024):
        for world in (1, 2, 4, 8):
            got = sorted(i for r in range(world) for i in D.shard_indices(n, r, world))
            assert got == list(range(n))
    assert [len(D.shard_indices(100, r, 8)) for r in range(8)] == [13, 13, 13, 13, 12, 12, 12, 12]


def test_sweep_and_batch_sharded_over_two_gloo_ranks():
    import torch.multiprocessing as mp
    from jpeg_dsp_studio_b200 import distributed as D
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process answer
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    qs = [5, 20, 35, 50, 65, 80, 95]
    want = D.sweep_sharded(OracleEngine(), img, qs, "4:2:0", False)
    frames = np.stack([np.random.default_rng(100 + k).integers(0, 256, (32, 48, 3), dtype=np.uint8)
                       for k in range(5)])
    want_agg = D.batch_sharded(OracleEngine(), frames, 5, 40, "4:2:2", False)
    for rank, table, agg in results:
        assert [t["quality"] for t in table] == qs
        assert table == want, f"rank {rank} sweep table differs from the single-process one"
        assert agg == want_agg, f"rank {rank} aggregate differs"
    # and the table agrees with the oracle's own metrics
    from oracle import numpy_port as P
    for t in want:
        o = P.compress_reconstruct(img, t["quality"], "4:2:0", False, want_maps=False)
        assert t["psnr_rgb"] == o["psnr_rgb"]
        assert abs(t["psnr_y"] - o["psnr_y"]) < 1e-9
        assert t["exact_bits"] == o["exact_bits"]
        assert t["bpp"] == o["bpp"] and t["compression_ratio"] == o["compression_ratio"]
        assert t["nonzero_count"] == o["nonzero_coeffs"]


# ---------------------------------------------------------------------------------------
# tile-band sharding of one frame: bounds, reduce and gather over gloo
# ---------------------------------------------------------------------------------------
class OracleBandEngine:
    """Duck-typed Engine.roundtrip_band: the oracle's whole-frame round trip, cut to the band -
    which is what jds_roundtrip_band must produce (tests/test_band_gpu.py checks that it does)."""

    def roundtrip_band(self, image, row0, row1, quality, mode, prefilter, precision="exact"):
        from oracle import numpy_port as P
        h, w = image.shape[:2]
        o = P.compress_reconstruct(image, quality, mode, prefilter, want_maps=False, want_metrics=False)
        rec = o["reconstructed_image"]
        a, b = image[row0:row1].astype(np.float64), rec[row0:row1].astype(np.float64)
        ya = 0.299 * a[..., 0] + 0.587 * a[..., 1] + 0.114 * a[..., 2]
        yb = 0.299 * b[..., 0] + 0.587 * b[..., 1] + 0.114 * b[..., 2]
        v = 2 if mode == "4:2:0" else 1
        hd = 1 if mode == "4:4:4" else 2
        nb = lambda n: (n + 7) // 8
        ny, nc = nb(h) * nb(w), nb(h // v) * nb(w // hd)
        planes = np.split(np.asarray(o["all_quantized_coeffs"]), [64 * ny, 64 * (ny + nc)])
        y0, y1 = row0 // 8, (nb(h) if row1 == h else row1 // 8)
        c0, c1 = row0 // (8 * v), (nb(h // v) if row1 == h else row1 // (8 * v))
        own = np.concatenate([planes[0][64 * y0 * nb(w):64 * y1 * nb(w)],
                              planes[1][64 * c0 * nb(w // hd):64 * c1 * nb(w // hd)],
                              planes[2][64 * c0 * nb(w // hd):64 * c1 * nb(w // hd)]])
        nnz, bits = P.bit_length_sum(own)
        centres = max(0, min(row1, h - 3) - max(row0, 3))
        m = SimpleNamespace(
            sse_rgb=int(np.sum((a - b) ** 2)), sse_y=float(np.sum((ya - yb) ** 2)),
            ssim_sum=[0.9 * centres * (w - 6)] * 4, ssim_count=centres * (w - 6),
            coeff_bits=bits, nnz=nnz, total_coeffs=own.size, luma_blocks=(y1 - y0) * nb(w))
        return SimpleNamespace(metrics=m, recon=np.ascontiguousarray(rec[row0:row1]))


def _band_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    from jpeg_dsp_studio_b200 import distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img = np.random.default_rng(21).integers(0, 256, (84, 64, 3), dtype=np.uint8)
        r = D.frame_banded(OracleBandEngine(), img, 45, "4:2:0", True)
        q.put((rank, r["band"], r["scalars"], r["recon"]))
    finally:
        dist.destroy_process_group()


def test_band_bounds_partition_the_rows():
    from jpeg_dsp_studio_b200 import distributed as D
    for h in (1, 7, 16, 17, 52, 250, 1080, 2160):
        for world in (1, 2, 3, 8, 200):
            bounds = [b for b in D.band_bounds(h, world) if b is not None]
            assert bounds[0][0] == 0 and bounds[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
            assert all(b[0] % 16 == 0 and (b[1] % 16 == 0 or b[1] == h) and b[1] > b[0] for b in bounds)
            sizes = [-(-(b[1] - b[0]) // 16) for b in bounds]
            assert max(sizes) - min(sizes) <= 1
    assert D.band_bounds(2160, 8)[0] == (0, 272) and D.band_bounds(2160, 8)[-1] == (1904, 2160)


def test_frame_banded_over_two_gloo_ranks():
    import torch.multiprocessing as mp
    from oracle import numpy_port as P
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_band_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    img = np.random.default_rng(21).integers(0, 256, (84, 64, 3), dtype=np.uint8)
    o = P.compress_reconstruct(img, 45, "4:2:0", True, want_maps=False)
    assert sorted(r[1] for r in results) == [(0, 48), (48, 84)]
    for rank, band, sc, recon in results:
        assert np.array_equal(recon, o["reconstructed_image"]), f"rank {rank}: gathered frame differs"
        assert sc["psnr_rgb"] == o["psnr_rgb"] and abs(sc["psnr_y"] - o["psnr_y"]) < 1e-9
        assert sc["exact_bits"] == o["exact_bits"] and sc["bpp"] == o["bpp"]
        assert sc["nonzero_count"] == o["nonzero_coeffs"]
