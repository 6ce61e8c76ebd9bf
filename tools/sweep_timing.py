"""Phase timing of the sharded sweep (debug): torchrun --nproc-per-node N tools/sweep_timing.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
os.environ.setdefault("JDS_SCRATCH_MB", "8192")
import jpeg_dsp_studio_b200 as J
from jpeg_dsp_studio_b200 import distributed as D
eng = J.Engine(local); eng.use_stream(torch.cuda.current_stream(dev).cuda_stream)
img = torch.from_numpy(np.random.default_rng(4).integers(0, 256, (2160, 3840, 3), dtype=np.uint8)).to(dev)
qs = list(range(1, 101)); mine = D.shard_indices(100, rank, world); my_qs = [qs[i] for i in mine]
for it in range(6):
    torch.cuda.synchronize(); 
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    outs = eng.sweep(img, my_qs, "4:2:0", False, precision="fast")
    t1 = time.perf_counter()
    rows = D.records_from_outputs(mine, my_qs, outs)
    t2 = time.perf_counter()
    table = D.gather_records(rows, 100, device=dev)
    t3 = time.perf_counter()
    res = [D.scalars_from_record(r, 2160, 3840) for r in table]
    t4 = time.perf_counter()
    if it >= 3:
        print(f"rank {rank} it {it}: sweep {1e3*(t1-t0):.3f} ms (gpu {sum(o.metrics.gpu_ms for o in outs):.3f}) records {1e3*(t2-t1):.3f} gather {1e3*(t3-t2):.3f} scalars {1e3*(t4-t3):.3f}", flush=True)
if world > 1: dist.destroy_process_group()
