"""Per-role cycle accounting of the tiled tcgen05 conv kernel (producer / MMA-issuer waits), via LFSR_TC_DBG_PTR."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
os.environ["LFSR_TC_DBG_PTR"] = hex(dbg.data_ptr())
import lfsr_b200
from lfsr_b200 import kernels as K
ops = K.CudaOps()
B = 16
for (cin, cout, hw, dil) in ((64, 64, 160, 5), (56, 224, 320, 1)):
    x = torch.rand(B, hw, hw, cin, device="cuda")
    y = torch.empty(B, hw, hw, cout, device="cuda")
    w = (torch.rand(cout, cin, 3, 3) - 0.5) * 0.1
    pc = K.pack_conv(w, dil=(dil, dil), pad=(dil, dil), device="cuda", tc=True)
    for _ in range(2):
        ops.conv(x, pc, y)
    torch.cuda.synchronize()
    d = dbg.view(148, 8).double().mean(0).tolist()
    ntile = B * hw * hw / 128 / 148
    print(f"{cin}->{cout} @{hw} d{dil}: tiles/CTA {ntile:.0f} | "
          f"MMA thread: wait-full {d[2]:.0f}, wait-acc {d[3]:.0f}, fence+issue+commit {d[4]:.0f} of {d[5]:.0f} cyc | "
          f"per k-stage {d[5]/ntile/18:.0f}")
