"""Stem conv (1 -> 60, 3x3 d5) on the Track-2 model's shape (ncu target / timing)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K
ops = K.CudaOps()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.rand(B, 160, 160, 1, device="cuda")
y = torch.empty(B, 160, 160, 60, device="cuda")
pc = K.pack_conv((torch.rand(60, 1, 3, 3) - 0.5), torch.rand(60), dil=(5, 5), pad=(5, 5), device="cuda")
for _ in range(3):
    ops.conv(x, pc, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv(x, pc, y)
e1.record()
torch.cuda.synchronize()
print(f"stem 1->60 batch {B}: {e0.elapsed_time(e1) / 10:.3f} ms")
