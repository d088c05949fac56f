"""sa_modulate (SA-modulator tail of the Track-2 model) alone: ncu target / timing. usage: run_sa.py [batch] [out16]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K
ops = K.CudaOps()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x, res, out = (torch.rand(B, 160, 160, 60, device="cuda") for _ in range(3))
o16 = K.alloc_nhwc16(B, 160, 160, 32, "cuda") if len(sys.argv) > 2 else None
dw, bs, bb = torch.rand(9, 60, device="cuda"), torch.rand(60, device="cuda") + 0.5, torch.rand(60, device="cuda")
am = torch.rand(B, 5, 5, 60, device="cuda")
fn = lambda: ops.sa_modulate(x, dw, bs, bb, am, 0.4, 0.6, res, out, 5, out16=o16)
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"sa_modulate c60 d5 batch {B}{' + fp16 copy' if o16 is not None else ''}: {ms:.3f} ms, {3 * x.numel() * 4 / ms / 1e6:.0f} GB/s algorithmic")
