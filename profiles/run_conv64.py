"""One 3x3 d=5 64->64 conv (DistgSSR.py:78-81 shape) on the tcgen05 path, batch 16 - ncu target."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K
ops = K.CudaOps()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = torch.rand(B, 160, 160, 64, device="cuda")
y = torch.empty_like(x)
w = (torch.rand(64, 64, 3, 3) - 0.5) * 0.1
pc = K.pack_conv(w, dil=(5, 5), pad=(5, 5), device="cuda", tc=True)
for _ in range(3):
    ops.conv(x, pc, y, act=2, slope=0.1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.conv(x, pc, y, act=2, slope=0.1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"conv3x3 d5 64->64 @160x160 batch {B}: {ms:.3f} ms, {2 * B * 160 * 160 * 64 * 4 / ms / 1e6:.0f} GB/s algorithmic")
