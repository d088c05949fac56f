"""BASELINE configs[2]: one synthetic 5x5x512x512 LF scene (1024 patches), patch-grid rows sharded over the ranks of a
torchrun job, stitched stripes all-gathered over NCCL. Prints scene/s and patches/s; rank 0 checks the gathered mosaic
against its own single-GPU result on a crop.   torchrun --nproc-per-node N profiles/run_scene_sharded.py [model] [view]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200

model = sys.argv[1] if len(sys.argv) > 1 else "EPIT"
view = int(sys.argv[2]) if len(sys.argv) > 2 else 512
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
A, s = 5, 4
torch.manual_seed(1234)
net = lfsr_b200.load_net(model, A, s).eval().to(dev)
lr = torch.from_numpy(np.random.RandomState(0).random_sample((A * view, A * view)).astype(np.float32)).to(dev)
mb = 32 if model == "EPIT" else 64
for _ in range(2):
    sr = lfsr_b200.scene.super_resolve_scene(net, lr, A, s, minibatch=mb)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    sr = lfsr_b200.scene.super_resolve_scene(net, lr, A, s, minibatch=mb)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
npatch = ((view + 15) // 16) ** 2
if rank == 0:
    # reference: rows of the patch grid computed alone on this GPU (first 2 rows) must equal the gathered mosaic there
    part = lfsr_b200.scene.super_resolve_rows(net, lr, A, s, minibatch=mb, rows=(0, 2))
    v = sr.view(A, view * s, A * view * s)[:, : 2 * 64]
    w = part.view(A, view * s, A * view * s)[:, : 2 * 64]
    print(f"{model} scene {view}x{view} views, {npatch} patches on {world} GPU(s): {ms.item():.1f} ms/scene, "
          f"{npatch / ms.item() * 1e3:.0f} patches/s, max|gathered - local| = {(v - w).abs().max().item():.2e}")
if world > 1:
    dist.destroy_process_group()
