"""EPI attention on EPIT's shape (ncu target / timing). usage: run_attn.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K
ops = K.CudaOps()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
A, S, E, heads = 5, 32, 128, 8
HW = A * S
T = B * HW * HW
qk = torch.rand(T, 2 * E, device="cuda") - 0.5
v = torch.rand(T, E, device="cuda")
out = torch.empty(T, E, device="cuda")
for name, args in (("h", dict(stride_a=S * HW, stride_s=HW, stride_b=HW * HW, stride_p=S, stride_q=1)),
                   ("v", dict(stride_a=S, stride_s=1, stride_b=HW * HW, stride_p=S * HW, stride_q=HW))):
    run = lambda: ops.epi_attention(qk, v, out, heads, E // heads, A, S, 5, B, A, S, **args)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"epi_attention direction {name} batch {B}: {e0.elapsed_time(e1) / 5:.3f} ms")
