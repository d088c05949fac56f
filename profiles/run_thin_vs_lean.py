"""the 18-channel spatial-branch convs of the Track-2 stages: FFMA2 thin kernel (fp32) vs the two-CTA tcgen05 kernel on fp16
activations. usage: python profiles/run_thin_vs_lean.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K, _native as N
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ops = K.CudaOps()
hw = 160
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20
g = torch.Generator().manual_seed(1)
w0 = (torch.rand(18, 18, 3, 3, generator=g) - 0.5) * 0.2
def padw(w, to):
    return torch.cat([w, w.new_zeros((to - w.shape[0],) + tuple(w.shape[1:]))], 0)
feat = torch.rand(B, hw, hw, 60, device="cuda")
cat = torch.zeros(B, hw, hw, 60, device="cuda")
t18 = torch.zeros(B, hw, hw, 20, device="cuda")
pk_thin = K.pack_conv(padw(w0, 20), dil=(5, 5), pad=(5, 5), device="cuda", tc=True)
ms = t(lambda: ops.conv(feat[..., 0:18], pk_thin, t18, act=N.ACT_LRELU, slope=0.1))
print(f"18->20 3x3 d5 fp32 (current path): {ms:.3f} ms")
lib = ops.lib
for cout_pad, name in ((24, "18->24 f16 -> f16 only"), (20, "18->20 f16 -> fp32 slice of cat")):
    pk = K.pack_conv(padw(w0, cout_pad), dil=(5, 5), pad=(5, 5), device="cuda", tc=True, tc16=True)
    x16 = K.alloc_nhwc16(B, hw, hw, 24, "cuda")
    x16[..., :18].copy_(feat[..., :18])
    c0 = lib.lfsr_conv_tc_lean_count()
    if cout_pad == 24:
        o16 = K.alloc_nhwc16(B, hw, hw, 24, "cuda")
        fn = lambda: ops.conv(x16[..., 0:18], pk, None, out16=o16, act=N.ACT_LRELU, slope=0.1)
    else:
        fn = lambda: ops.conv(x16[..., 0:18], pk, cat[..., 0:20])
    ms = t(fn)
    print(f"{name}: {ms:.3f} ms (lean launches {lib.lfsr_conv_tc_lean_count() - c0} of 23)")

for cin_pad in (24, 32, 64):
    wp = torch.cat([padw(w0, 24), torch.zeros(24, cin_pad - 18, 3, 3)], 1)
    pk = K.pack_conv(wp, dil=(5, 5), pad=(5, 5), device="cuda", tc=True, tc16=True)
    x16 = K.alloc_nhwc16(B, hw, hw, cin_pad, "cuda")
    x16.zero_()
    x16[..., :18].copy_(feat[..., :18])
    o16 = K.alloc_nhwc16(B, hw, hw, 24, "cuda")
    ms = t(lambda: ops.conv(x16, pk, None, out16=o16, act=N.ACT_LRELU, slope=0.1))
    print(f"{cin_pad}(18 real)->24 f16 -> f16 only, ld {x16.stride(2)}: {ms:.3f} ms")

xs16 = K.alloc_nhwc16(B, hw, hw, 32, "cuda")
print(f"to_f16 feat[..., 0:32] -> xs16: {t(lambda: ops.to_f16(feat[..., 0:32], xs16)):.3f} ms")
wp = torch.cat([padw(w0, 20), torch.zeros(20, 14, 3, 3)], 1)
pk = K.pack_conv(wp, dil=(5, 5), pad=(5, 5), device="cuda", tc=True, tc16=True)
print(f"32(18 real)->20 f16 -> fp32 cat[..., 0:20] (ld 60): {t(lambda: ops.conv(xs16, pk, cat[..., 0:20])):.3f} ms")
d20 = torch.zeros(B, hw, hw, 20, device="cuda")
print(f"32(18 real)->20 f16 -> fp32 dense 20-channel buffer: {t(lambda: ops.conv(xs16, pk, d20)):.3f} ms")
wp = torch.cat([padw(w0, 32), torch.zeros(32, 14, 3, 3)], 1)
pk = K.pack_conv(wp, dil=(5, 5), pad=(5, 5), device="cuda", tc=True, tc16=True)
d32 = torch.zeros(B, hw, hw, 32, device="cuda")
print(f"32(18 real)->32 f16 -> fp32 dense 32-channel buffer: {t(lambda: ops.conv(xs16, pk, d32)):.3f} ms")
