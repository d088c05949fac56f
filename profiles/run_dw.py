"""Depthwise kernels on the shapes of the two Track-2 models (ncu target / timing). usage: run_dw.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K
ops = K.CudaOps()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
def timeit(name, fn, nbytes):
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
    print(f"{name}: {ms:.3f} ms, {nbytes / ms / 1e6:.0f} GB/s algorithmic")
# SA modulator of MyEfficientLFNet: 60-channel grouped trunk, dilation 5
x, res, out = (torch.rand(B, 160, 160, 60, device="cuda") for _ in range(3))
dw, bs, bb = torch.rand(9, 60, device="cuda"), torch.rand(60, device="cuda") + 0.5, torch.rand(60, device="cuda")
am = torch.rand(B, 5, 5, 60, device="cuda")
timeit("sa_modulate c60 d5", lambda: ops.sa_modulate(x, dw, bs, bb, am, 0.4, 0.6, res, out, 5), 3 * x.numel() * 4)
# FastConvSSM dilated depthwise quartet and MultiScaleSpatial of MyEfficientLFNetV4_5
g = torch.rand(B, 160, 160, 128, device="cuda")
cat4 = torch.empty(B, 160, 160, 256, device="cuda")
br = [dict(w=torch.rand(9, 64, device="cuda"), kh=3, kw=3, dil=(d, d), in_c0=64, out_c0=k * 64, c=64) for k, d in enumerate((1, 2, 4, 8))]
timeit("dwconv_multi 4 x 3x3 d1/2/4/8 c64", lambda: ops.dwconv_multi(g, cat4, br), (64 + 256) * B * 160 * 160 * 4)
for d in (1, 8):
    timeit(f"dwconv 3x3 d{d} c64", lambda: ops.dwconv_multi(g, cat4, [dict(w=br[0]["w"], kh=3, kw=3, dil=(d, d), in_c0=64, out_c0=0, c=64)]),
           128 * B * 160 * 160 * 4)
f = torch.rand(B, 160, 160, 64, device="cuda")
ms_ = torch.empty(B, 160, 160, 64, device="cuda")
br2 = [dict(w=torch.ones(1, 16, device="cuda"), kh=1, kw=1, in_c0=0, out_c0=0, c=16)]
for j, k in enumerate((3, 5, 7), 1):
    br2.append(dict(w=torch.rand(k * k, 16, device="cuda"), kh=k, kw=k, in_c0=16 * j, out_c0=16 * j, c=16))
timeit("dwconv_multi 1/3/5/7 c16 each", lambda: ops.dwconv_multi(f, ms_, br2), 128 * B * 160 * 160 * 4)
