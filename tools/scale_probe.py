"""Quick multi-GPU probe (torchrun): frames resident in HBM, K timed calls per rank, prints per-rank and max wall time.
Used to tune the host side (worker threads per rank) on a multi-GPU host; bench.py is the contract benchmark."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sift_b200 import capi, shard  # noqa: E402
from sift_b200.synth import synth_frame  # noqa: E402

W, H, B, N = 1920, 1080, 512, 32
rank, local_rank, world = shard.world()
torch.cuda.set_device(local_rank)
dist = shard.init_process_group("nccl") if world > 1 else None
mode = sys.argv[1] if len(sys.argv) > 1 else "device"
host = torch.empty((N, H, W), dtype=torch.float32).pin_memory()
for i in range(N):
    host[i] = torch.from_numpy(synth_frame(W, H, rank * N + i))
if mode == "u8":
    host = host.to(torch.uint8).pin_memory()
frames = host.to(torch.device("cuda", local_rank)) if mode == "device" else host
g = capi.SiftGpu(3, 5, 1.6, capi.SQRT2_F32, False, max_width=W, max_height=H, max_batch=16, device=local_rank)
esz = frames.element_size()
arr = (capi.Image * B)()
for j in range(B):
    arr[j] = capi.Image(frames.data_ptr() + (j % N) * W * H * esz, W, H, 0, capi.DTYPE_U8 if mode == "u8" else capi.DTYPE_F32,
                        capi.MEM_DEVICE if mode == "device" else capi.MEM_HOST, None)
for _ in range(3):
    g.run_raw(arr, B)
torch.cuda.synchronize()
if dist:
    dist.barrier()
t0 = time.perf_counter()
K = 5
for _ in range(K):
    g.run_raw(arr, B)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / K
mx, _ = shard.reduce_max_sum(dist, torch.device("cuda", local_rank), [wall], [1.0])
print(f"rank {rank}: {1e3 * wall:.1f} ms/step", flush=True)
if rank == 0:
    print(f"== {mode} threads={os.environ.get('SIFT_GPU_HOST_THREADS', 'default')}: max {1e3 * mx[0]:.1f} ms/step -> {world * B / mx[0]:.0f} images/s", flush=True)
g.close()
if dist:
    dist.barrier()
    dist.destroy_process_group()
