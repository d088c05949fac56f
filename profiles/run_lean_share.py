"""how many tensor-core conv launches of one forward take the two-CTAs-per-SM kernel, and the forward time.
usage: python profiles/run_lean_share.py [batch]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = K.CudaOps().lib
for name, scale in (("MyEfficientLFNet", 4), ("EPIT", 4), ("DistgSSR", 4), ("DistgSSR", 2), ("LF_InterNet", 4), ("MyEfficientLFNetV4_5", 4)):
    torch.manual_seed(1234)
    net = lfsr_b200.load_net(name, 5, scale).eval().to("cuda")
    x = torch.rand(B, 1, 160, 160, device="cuda")
    with torch.no_grad():
        l0, c0 = lib.lfsr_launch_count(), lib.lfsr_conv_tc_lean_count()
        net.forward_static(x)          # first call: allocation + eager launches (counted), then graph capture / replay
        l1, c1 = lib.lfsr_launch_count(), lib.lfsr_conv_tc_lean_count()
        for _ in range(3): net.forward_static(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): net.forward_static(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
    print(f"{name} x{scale} batch {B}: {l1 - l0} launches, {c1 - c0} on the lean kernel, {ms:.2f} ms, {B / ms * 1e3:.0f} patches/s", flush=True)
    del net
    torch.cuda.empty_cache()
