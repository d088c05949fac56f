"""mel_epi_branch on the Track-2 model's shape (ncu target / timing), CUDA-core and tcgen05 variants.
usage: run_epi.py [batch] [tc|f32|both]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K
ops = K.CudaOps()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
which = sys.argv[2] if len(sys.argv) > 2 else "both"
trunk = torch.rand(B, 160, 160, 60, device="cuda")
cat = torch.zeros(B, 160, 160, 60, device="cuda")
w = (torch.rand((2 * 11 + 9) * 18 + 6 * 18 * 18, device="cuda") - 0.5) * 0.3
for tc in ([False, True] if which == "both" else [which == "tc"]):
    for _ in range(3):
        ops.mel_epi_branch(trunk[..., 40:58], w, cat[..., 40:58], 11, 5, 0.1, tc=tc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.mel_epi_branch(trunk[..., 40:58], w, cat[..., 40:58], 11, 5, 0.1, tc=tc)
    e1.record()
    torch.cuda.synchronize()
    print(f"mel_epi_branch{'_tc' if tc else ''} batch {B}: {e0.elapsed_time(e1) / 10:.3f} ms")
# all-tensor-core variant (depthwise taps as shifted-row MMAs)
img = ops.mel_epi_pack(w, 11, "cuda")
trunk16 = K.alloc_nhwc16(B, 160, 160, 64, "cuda")
trunk16[..., :60].copy_(trunk)
if img is not None:
    for _ in range(3):
        ops.mel_epi_branch_mma(trunk[..., 40:58], trunk16[..., 40:58], img, cat[..., 40:58], 11, 5, 0.1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.mel_epi_branch_mma(trunk[..., 40:58], trunk16[..., 40:58], img, cat[..., 40:58], 11, 5, 0.1)
    e1.record()
    torch.cuda.synchronize()
    print(f"mel_epi_branch_mma batch {B}: {e0.elapsed_time(e1) / 10:.3f} ms")
