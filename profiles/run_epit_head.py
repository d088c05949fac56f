"""EPIT's reconstruction head alone: 1x1 64->1024 + PixelShuffle(4) + LReLU with the 3x3 head conv's tap projection in the
epilogue, then the 9-tap gather. usage: python profiles/run_epit_head.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K, _native as N
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ops = K.default_ops()
torch.manual_seed(1234)
net = lfsr_b200.load_net("EPIT", 5, 4).eval().to("cuda")
pk = net._get_packed(torch.device("cuda", 0), ops)
fb = torch.rand(B, 160, 160, 64, device="cuda")
taps = torch.empty(B, 640, 640, 12, device="cuda")[..., :9]
Y = torch.zeros(B, 640, 640, 1, device="cuda")
sh = (4, 4, N.SHUF_CHANNEL_MAJOR)
def t(name, fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms (batch {B})")
t("up0 1x1 64->1024 + PS4 + LReLU + tail projection", lambda: ops.conv(fb, pk["up0"], taps, act=N.ACT_LRELU, slope=0.2, shuffle=sh, tail=(pk["tail_w"], 9, 64)))
t("tap_gather", lambda: ops.tap_gather(taps, 3, 3, None, Y, Y))
