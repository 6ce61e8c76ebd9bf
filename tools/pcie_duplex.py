"""Host-memory / PCIe ceiling of the box, without any kernel of ours: every rank copies a pinned
398 MB buffer host->device and another device->host AT THE SAME TIME (two streams), like the
host->host batch call does, and reports GB/s per direction.

    python tools/pcie_duplex.py                                             # one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/pcie_duplex.py   # all eight at once

If eight concurrent ranks get ~1/5 of the single-rank figure each, the e2e scaling of bench.py is
the box's host-memory path, not the code."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import bind_to_gpu_numa_node  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
numa = bind_to_gpu_numa_node(local)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = 16 * 3840 * 2160 * 3
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.random_(0, 256)
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(duplex, reps=10):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream(dev))
    s2.wait_stream(torch.cuda.current_stream(dev))
    for _ in range(reps):
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        if duplex:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.current_stream(dev).wait_stream(s1)
    torch.cuda.current_stream(dev).wait_stream(s2)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for _ in range(2):
    run(True, 2)
ms_h2d = run(False)
ms_dup = run(True)
if rank == 0:
    gb = n / 1e9
    print(f"ranks {world} ({numa}): H2D alone {gb / ms_h2d * 1e3:.1f} GB/s per rank; duplex {gb / ms_dup * 1e3:.1f} GB/s "
          f"per rank each way = {world * 2 * gb / ms_dup * 1e3:.0f} GB/s through host memory in total "
          f"({ms_dup:.2f} ms per 398 MB each way)")
if world > 1:
    dist.destroy_process_group()
