import os, sys, torch
sys.path.insert(0, "/root/repo")
import lfsr_b200
from lfsr_b200 import kernels as K, _native as N
ops = K.CudaOps()
B = 64
trunk = torch.rand(B, 160, 160, 60, device="cuda")
cat60 = torch.zeros(B, 160, 160, 60, device="cuda")
cat72 = torch.zeros(B, 160, 160, 72, device="cuda")
ang5 = torch.rand(B, 32, 32, 20, device="cuda")
w20 = (torch.rand(5, 5, 18, 20, device="cuda") - 0.5) * 0.3
w24 = (torch.rand(5, 5, 18, 24, device="cuda") - 0.5) * 0.3
def t(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/10:.3f} ms")
t("ang_expand -> 20-float group of 60 (80 B slices)", lambda: ops.ang_expand(ang5[..., :18], w20, trunk[..., 20:40], cat60[..., 20:40], 5, N.ACT_LRELU, 0.1, 0.1))
t("ang_expand -> 24-float group of 72 (96 B = 3 sectors)", lambda: ops.ang_expand(ang5[..., :18], w24, trunk[..., 20:44], cat72[..., 24:48], 5, N.ACT_LRELU, 0.1, 0.1))
t("ang_expand -> 24-float group of 72, no residual", lambda: ops.ang_expand(ang5[..., :18], w24, None, cat72[..., 24:48], 5, N.ACT_LRELU, 0.1, 0.1))
t("ang_expand -> 20-float group of 60, no residual", lambda: ops.ang_expand(ang5[..., :18], w20, None, cat60[..., 20:40], 5, N.ACT_LRELU, 0.1, 0.1))
vm60 = torch.empty(B, 5, 5, 60, device="cuda"); vm72 = torch.empty(B, 5, 5, 72, device="cuda")
t("block_mean c60", lambda: ops.block_mean(cat60, vm60, 32, 32))
t("block_mean c72", lambda: ops.block_mean(cat72, vm72, 32, 32))
