"""MMA-lane cycle accounting of the dominant kernel (CTA pairs, fp16 operands, tail projection) - probe library.
usage: python profiles/probe_dominant_roles.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
os.environ["LFSR_TC_DBG_PTR"] = hex(dbg.data_ptr())
os.environ.setdefault("LFSR_PROBE_LIB", "1")
import lfsr_b200
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = lfsr_b200.load_net("MyEfficientLFNet", 5, 4).eval().to("cuda")
call, info = net.dominant_kernel(batch)
for _ in range(3): call()
torch.cuda.synchronize()
d = dbg[:148 * 8].view(148, 8).double()
lead = d[0::2].mean(0).tolist()
tiles = batch * 320 * 320 / 128 / 148          # per SM
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{info['name']}: batch {batch}: {ms:.3f} ms (DBG instantiation), {ms * 1.965e6 / tiles:.0f} cycles per 128-pixel tile per SM")
print(f"leader MMA lane per tile pair: wait-full {lead[2]/tiles:.0f}, wait-acc {lead[3]/tiles:.0f}, issue {lead[4]/tiles:.0f} of {lead[5]/tiles:.0f}")
print(f"epilogue warp 0 of the leader: total {lead[7]/tiles:.0f} cycles per tile")
