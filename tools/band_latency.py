"""Single-frame latency with the frame's rows sharded over the ranks.
    python tools/band_latency.py                                   # 1 GPU: whole frame
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/band_latency.py                  # N GPUs: N bands
Prints one JSON line on rank 0: ms per frame (device-resident input, max over ranks, barrier
on both sides), for the metrics-only exchange and with the reconstructed rows gathered."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import jpeg_dsp_studio_b200 as J
from jpeg_dsp_studio_b200 import distributed as D

eng = J.Engine(local)
out = {"world": world, "cases": []}
for (h, w) in ((2160, 3840), (4320, 7680)):
    img = np.random.default_rng(3).integers(0, 256, (h, w, 3), dtype=np.uint8)
    d = torch.from_numpy(img).to(dev)
    for precision, pf in (("fast", False), ("fast", True), ("exact", False)):
        row = {"frame": f"{h}x{w}", "precision": precision, "prefilter": pf}
        for gather in (False, True):
            def once():
                return D.frame_banded(eng, d, 50, "4:2:0", pf, precision=precision, device=dev, gather_recon=gather)
            for _ in range(3):
                r = once()
            ts = []
            for _ in range(10):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = once()
                torch.cuda.synchronize()
                t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ts.append(float(t.item()) * 1e3)
            row["ms_gather" if gather else "ms_metrics"] = round(float(np.median(ts)), 4)
            row["psnr_rgb"] = r["scalars"]["psnr_rgb"]; row["bpp"] = r["scalars"]["bpp"]
            row["ssim_rgb"] = r["scalars"]["ssim_rgb"]
        out["cases"].append(row)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
