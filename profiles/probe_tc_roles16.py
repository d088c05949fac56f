"""Per-role cycle accounting of the tcgen05 conv kernel on fp16 operands (probe library, LFSR_TC_DBG_PTR).
usage: python profiles/probe_tc_roles16.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
os.environ["LFSR_TC_DBG_PTR"] = hex(dbg.data_ptr())
os.environ.setdefault("LFSR_PROBE_LIB", "1")
import lfsr_b200
from lfsr_b200 import kernels as K, _native as N
ops = K.CudaOps()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
def ev(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (cin, cout, hw, k, dil) in ((64, 64, 160, 1, 1), (64, 64, 160, 3, 1), (64, 64, 160, 3, 5), (128, 64, 160, 3, 5), (144, 64, 160, 1, 1)):
    w = (torch.rand(cout, cin, k, k) - 0.5) * 0.1
    p = dil * (k // 2)
    pc = K.pack_conv(w, dil=(dil, dil), pad=(p, p), device="cuda", tc=True, tc16=True)
    x16 = K.alloc_nhwc16(B, hw, hw, cin, "cuda"); x16.copy_(torch.rand(B, hw, hw, cin, device="cuda"))
    y = torch.empty(B, hw, hw, cout, device="cuda")
    y16 = K.alloc_nhwc16(B, hw, hw, cout, "cuda")
    r = torch.rand(B, hw, hw, cout, device="cuda")
    for name, fn in (("f16->f32", lambda: ops.conv(x16, pc, y, act=N.ACT_LRELU, slope=0.1)),
                     ("f16->f16", lambda: ops.conv(x16, pc, None, out16=y16, act=N.ACT_LRELU, slope=0.1)),
                     ("f16->f32+f16+res", lambda: ops.conv(x16, pc, y, out16=y16, res=r))):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        d = dbg[:148 * 8].view(148, 8).double().mean(0).tolist()
        e = (dbg[148 * 8:].view(148, 8).double().mean(0) / (B * hw * hw / 128 / 148)).tolist()
        ntile = B * hw * hw / 128 / 148
        ms = ev(fn)
        print(f"{k}x{k} d{dil} {cin}->{cout} @{hw} x{B} {name:18s} {ms:.3f} ms ({ms * 1.965e6 / ntile:.0f} cyc/tile) | MMA lane: wait-full {d[2]/ntile:.0f}, "
              f"wait-acc {d[3]/ntile:.0f}, issue {d[4]/ntile:.0f} of {d[5]/ntile:.0f} | epi warp0: wait-acc-full {d[6]/ntile:.0f}, "
              f"staging wait {d[0]/ntile:.0f}, work {d[1]/ntile:.0f} of {d[7]/ntile:.0f} = prefetch {e[1]:.0f} + acc wait {e[2]:.0f} + tmem ld {e[3]:.0f} + math {e[4]:.0f} "
              f"+ stores {e[5]:.0f} + fence/tma {e[6]:.0f}", flush=True)
