"""Role cycle accounting for the 18-channel branch convs of the Track-2 model (slices of the grouped 60-channel trunk)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
os.environ["LFSR_TC_DBG_PTR"] = hex(dbg.data_ptr())
os.environ["LFSR_TC_VERBOSE"] = "1"
os.environ.setdefault("LFSR_PROBE_LIB", "1")   # liblfsr_probe.so: probe kernels + debug hooks (not in the product library)
import lfsr_b200
from lfsr_b200 import kernels as K
ops = K.CudaOps()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
trunk = torch.rand(B, 160, 160, 60, device="cuda")
cat = torch.zeros(B, 160, 160, 60, device="cuda")
for (cin, cout, k, dil, src, dst) in ((18, 18, 3, 5, trunk[..., 0:18], cat[..., 0:18]), (18, 20, 3, 5, trunk[..., 0:18], cat[..., 0:20]),
                                      (54, 60, 3, 5, torch.rand(B, 160, 160, 56, device="cuda")[..., :54], cat),
                                      (60, 54, 1, 1, trunk, torch.zeros(B, 160, 160, 56, device="cuda")[..., :54]),
                                      (60, 56, 1, 1, trunk, torch.zeros(B, 160, 160, 56, device="cuda"))):
    w = (torch.rand(cout, cin, k, k) - 0.5) * 0.1
    p = dil * (k // 2)
    pc = K.pack_conv(w, dil=(dil, dil), pad=(p, p), device="cuda", tc=True)
    for _ in range(2):
        ops.conv(src, pc, dst, act=2, slope=0.1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv(src, pc, dst, act=2, slope=0.1)
    e1.record()
    torch.cuda.synchronize()
    d = dbg[:148 * 8].view(148, 8).double().mean(0).tolist()
    ntile = B * 160 * 160 / 128 / 148
    print(f"{k}x{k} d{dil} {cin}->{cout}: {e0.elapsed_time(e1) / 10:.3f} ms | tiles/CTA {ntile:.0f} | MMA thread: wait-full {d[2]/ntile:.0f}, "
          f"wait-acc {d[3]/ntile:.0f}, issue {d[4]/ntile:.0f} of {d[5]/ntile:.0f} cyc/tile | epilogue: wait {d[6]/ntile:.0f}, work {d[1]/ntile:.0f} of {d[7]/ntile:.0f}")
