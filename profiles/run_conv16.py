"""tcgen05 conv with fp32 / TF32 operands vs fp16 operands (fp16 activations in HBM) on one layer shape.
usage: python profiles/run_conv16.py cin cout k dil batch hw"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K, _native as N
cin, cout, k, dil, B, hw = (int(v) for v in sys.argv[1:7])
ops = K.CudaOps()
w = (torch.rand(cout, cin, k, k) - 0.5) * 0.1
p = dil * (k // 2)
pc = K.pack_conv(w, dil=(dil, dil), pad=(p, p), device="cuda", tc=True, tc16=True)
x = torch.rand(B, hw, hw, cin, device="cuda")
y = torch.empty(B, hw, hw, cout, device="cuda")
x16 = K.alloc_nhwc16(B, hw, hw, cin, "cuda"); x16.copy_(x)
y16 = K.alloc_nhwc16(B, hw, hw, cout, "cuda")
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20
fl = 2 * B * hw * hw * cin * cout * k * k
for name, fn in (("fp32 in (tf32 MMA) -> fp32 out", lambda: ops.conv(x, pc, y, act=N.ACT_LRELU, slope=0.1)),
                 ("fp16 in -> fp32 out", lambda: ops.conv(x16, pc, y, act=N.ACT_LRELU, slope=0.1)),
                 ("fp16 in -> fp16 out", lambda: ops.conv(x16, pc, None, out16=y16, act=N.ACT_LRELU, slope=0.1)),
                 ("fp16 in -> fp32 + fp16 out", lambda: ops.conv(x16, pc, y, out16=y16, act=N.ACT_LRELU, slope=0.1))):
    ms = t(fn)
    print(f"conv {k}x{k} d{dil} {cin}->{cout} @{hw} batch {B}  {name:30s} {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
