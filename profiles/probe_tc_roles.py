"""Per-role cycle accounting of the tiled tcgen05 conv kernel (MMA-issuer and epilogue waits), via LFSR_TC_DBG_PTR.
usage: python profiles/probe_tc_roles.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
os.environ["LFSR_TC_DBG_PTR"] = hex(dbg.data_ptr())
os.environ.setdefault("LFSR_PROBE_LIB", "1")   # liblfsr_probe.so: probe kernels + debug hooks (not in the product library)
import lfsr_b200
from lfsr_b200 import kernels as K
ops = K.CudaOps()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for (cin, cout, hw, k, dil, res) in ((64, 64, 160, 1, 1, False), (64, 64, 160, 1, 1, True), (128, 64, 160, 1, 1, False),
                                     (64, 128, 160, 1, 1, False), (256, 64, 160, 1, 1, False),
                                     (64, 64, 160, 3, 5, False), (56, 224, 320, 3, 1, False)):
    x = torch.rand(B, hw, hw, cin, device="cuda")
    y = torch.empty(B, hw, hw, cout, device="cuda")
    r = torch.rand(B, hw, hw, cout, device="cuda") if res else None
    w = (torch.rand(cout, cin, k, k) - 0.5) * 0.1
    p = dil * (k // 2)
    pc = K.pack_conv(w, dil=(dil, dil), pad=(p, p), device="cuda", tc=True)
    for _ in range(2):
        ops.conv(x, pc, y, res=r)
    torch.cuda.synchronize()
    d = dbg[:148 * 8].view(148, 8).double().mean(0).tolist()
    ntile = B * hw * hw / 128 / 148
    print(f"{k}x{k} {cin}->{cout} @{hw} d{dil} res={int(res)}: tiles/CTA {ntile:.0f} | "
          f"MMA thread: wait-full {d[2]/ntile:.0f}, wait-acc {d[3]/ntile:.0f}, issue {d[4]/ntile:.0f} of {d[5]/ntile:.0f} cyc/tile | "
          f"epilogue warp0: wait-acc-full {d[6]/ntile:.0f}, staging-buffer wait {d[0]/ntile:.0f}, tile work {d[1]/ntile:.0f} of {d[7]/ntile:.0f} cyc/tile")
