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
