"""Whole-forward throughput of one network at one batch (graph replay). usage: run_minibatch2.py MODEL SCALE BATCH"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
name, scale, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(1234)
net = lfsr_b200.load_net(name, 5, scale).eval().to("cuda")
x = torch.rand(B, 1, 160, 160, device="cuda")
for _ in range(2):
    net.forward_static(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    net.forward_static(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{name} x{scale} batch {B} LFSR_FP16_OPS={os.environ.get('LFSR_FP16_OPS', '1')}: {ms:.2f} ms, {B / ms * 1e3:.0f} patches/s")
