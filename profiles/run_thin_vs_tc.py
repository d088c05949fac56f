"""18 -> 20 3x3 d5 slice conv of the Track-2 trunk: FFMA2 thin kernel vs the tcgen05 path (usage: run_thin_vs_tc.py [batch])"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
trunk = torch.rand(B, 160, 160, 60, device="cuda")
cat = torch.zeros(B, 160, 160, 60, device="cuda")
pc = K.pack_conv((torch.rand(20, 18, 3, 3) - 0.5) * 0.2, torch.rand(20), dil=(5, 5), pad=(5, 5), device="cuda", tc=True)
outs = {}
for thin in (True, False):
    ops = K.CudaOps()
    ops.use_thin = thin
    run = lambda: ops.conv(trunk[..., 0:18], pc, cat[..., 0:20], act=2, slope=0.1)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    outs[thin] = cat[..., 0:20].clone()
    print(f"18->20 3x3 d5 batch {B}: {'thin FFMA2' if thin else 'tcgen05 tf32'} {e0.elapsed_time(e1) / 10:.3f} ms")
print("max |thin - tc| =", (outs[True] - outs[False]).abs().max().item())
