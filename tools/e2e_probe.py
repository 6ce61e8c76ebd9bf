"""Host->host batch call (16 x 4K, pinned buffers) against the box's own duplex PCIe ceiling:
    [JDS_PIPE_CHUNK=n] python tools/e2e_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jpeg_dsp_studio_b200 as J
eng = J.Engine(0)
n = 16
frames = np.stack([np.random.default_rng(4000 + k).integers(0, 256, (2160, 3840, 3), dtype=np.uint8) for k in range(n)])
h_in = torch.from_numpy(frames).pin_memory()
h_out = torch.empty_like(h_in).pin_memory()
for _ in range(3):
    eng.roundtrip_batch(h_in, 50, "4:2:0", False, precision="fast", recon_out=h_out)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 10
for _ in range(reps):
    eng.roundtrip_batch(h_in, 50, "4:2:0", False, precision="fast", recon_out=h_out)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / reps * 1e3
gb = h_in.numel() / 1e9
print(f"JDS_PIPE_CHUNK={os.environ.get('JDS_PIPE_CHUNK', '-')}: {ms:.3f} ms per 16 x 4K host->host = {gb / ms * 1e3:.1f} GB/s each way, "
      f"{n * 2160 * 3840 / ms / 1e3:.0f} Mpixel/s")
# the same batches overlapped over two contexts (compress_stream's loop)
from jpeg_dsp_studio_b200.engine import stream_engines
engs = stream_engines(0, 2)
h_out2 = torch.empty_like(h_in).pin_memory()
outs = (h_out, h_out2)
def pipelined(reps):
    prev = None
    for k in range(reps):
        nxt = engs[k & 1].roundtrip_batch_begin(h_in, 50, "4:2:0", False, precision="fast", recon_out=outs[k & 1])
        if prev is not None:
            prev.result()
        prev = nxt
    prev.result()
pipelined(4)
torch.cuda.synchronize()
t0 = time.perf_counter()
pipelined(reps)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / reps * 1e3
print(f"pipelined over two contexts: {ms:.3f} ms per 16 x 4K host->host = {gb / ms * 1e3:.1f} GB/s each way, "
      f"{n * 2160 * 3840 / ms / 1e3:.0f} Mpixel/s; results equal: {bool(torch.equal(h_out, h_out2))}")
