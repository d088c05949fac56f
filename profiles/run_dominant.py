"""Launch only the dominant kernel of BASELINE configs[1] (upsampler conv 54->216 @320^2 + PixelShuffle + LReLU)
a few times - the target of the `ncu --set full` capture.  usage: python profiles/run_dominant.py [batch] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
net = lfsr_b200.load_net("MyEfficientLFNet", 5, 4).eval()
net = net.to(dev)
call, info = net.dominant_kernel(batch)
for _ in range(reps):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{info['name']}: batch {batch}: {ms:.3f} ms/launch, {info['bytes'] / ms / 1e6:.1f} GB/s algorithmic, "
      f"{info['flops'] / ms / 1e9:.1f} TFLOP/s")
