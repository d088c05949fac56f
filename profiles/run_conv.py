"""Micro-benchmark of one lfsr conv launch on the tcgen05 path (run on the GPU box).
usage: python profiles/run_conv.py cin cout k dil batch hw [res] [mul] [act=N]
prints device ms/launch (events around 20 back-to-back launches), host us/call and the algorithmic GB/s."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K

a = sys.argv[1:]
cin, cout, k, dil, B, hw = (int(v) for v in a[:6])
flags = a[6:]
act = next((int(f[4:]) for f in flags if f.startswith("act=")), 0)
ops = K.CudaOps()
x = torch.rand(B, hw, hw, cin, device="cuda")
y = torch.empty(B, hw, hw, cout, device="cuda")
res = torch.rand(B, hw, hw, cout, device="cuda") if "res" in flags else None
mul = torch.rand(B, hw, hw, cout, device="cuda") if "mul" in flags else None
w = (torch.rand(cout, cin, k, k) - 0.5) * 0.1
p = dil * (k // 2)
pc = K.pack_conv(w, dil=(dil, dil), pad=(p, p), device="cuda", tc=True)
kw = dict(act=act, slope=0.1, res=res, mul=mul)
for _ in range(3):
    ops.conv(x, pc, y, **kw)
torch.cuda.synchronize()
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    ops.conv(x, pc, y, **kw)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
byt = B * hw * hw * 4 * (cin + cout * (1 + (res is not None) + (mul is not None)))
print(f"conv {k}x{k} d{dil} {cin}->{cout} @{hw}x{hw} batch {B} {' '.join(flags)}: {ms:.3f} ms/launch, host {1e6 * (t1 - t0) / n:.0f} us/call, "
      f"{byt / ms / 1e6:.0f} GB/s algorithmic, {2 * B * hw * hw * cin * cout * k * k / ms / 1e9:.1f} TFLOP/s")
