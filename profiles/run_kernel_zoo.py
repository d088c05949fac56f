"""One launch of every kernel family of the hot path on its production shape (Track-2 model / EPIT / DistgSSR layers at
batch B), each preceded by one warm-up launch: the target of the per-kernel `ncu --set full` capture
(profiles/rNN_kernel_zoo_ncu.json) and, run plainly, a CUDA-event timing table with algorithmic GB/s and TFLOP/s.
usage: python profiles/run_kernel_zoo.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K, _native as N

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"
ops = K.CudaOps()
rows = []


def run(name, fn, nbytes, flops=0.0, reps=int(os.environ.get("LFSR_ZOO_REPS", "5"))):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rows.append((name, ms, nbytes / ms / 1e6, flops / ms / 1e9))
    print(f"{name:58s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.0f} GB/s alg  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


def conv_case(name, cin, cout, k, dil, hw, real=None, **kw):
    x = torch.rand(B, hw, hw, cin, device=dev)
    y = torch.empty(B, hw, hw, cout, device=dev)
    w = (torch.rand(cout, cin, k, k) - 0.5) * 0.1
    p = dil * (k // 2)
    pc = K.pack_conv(w, dil=(dil, dil), pad=(p, p), device=dev, tc=True)
    rc_in, rc_out = real or (cin, cout)
    run(name, lambda: ops.conv(x, pc, y, **kw), B * hw * hw * 4 * (rc_in + rc_out), 2.0 * B * hw * hw * rc_in * rc_out * k * k)


# ---- patch pipeline (128x128 views -> 64 patches; bytes = SURVEY 8d per-patch figures) ----
h0 = 128
lr = torch.rand(5 * h0, 5 * h0, device=dev)
sub = torch.empty(64, 1, 160, 160, device=dev)
run("divide_kernel (64 patches)", lambda: ops.divide_rows(lr, sub, 5, h0, h0, 32, 16, 0, 8), 64 * 128e3)
sr = torch.rand(64, 1, 640, 640, device=dev)
mos = torch.empty(5 * h0 * 4, 5 * h0 * 4, device=dev)
run("integrate_kernel (64 patches)", lambda: ops.integrate_rows(sr, mos, 5, 128, 64, h0 * 4, h0 * 4, 8, 8, 0, 8), 64 * 819.2e3)
hr = torch.rand_like(mos)
acc = torch.zeros(50, dtype=torch.float64, device=dev)
run("metric_kernel (5x5 views of 512^2)", lambda: ops.metric_sums(hr, mos, 5, h0 * 4, h0 * 4, acc), 2 * mos.numel() * 4)
xin = torch.rand(B, 1, 160, 160, device=dev)
yout = torch.empty(B, 1, 640, 640, device=dev)
run("interp_kernel bicubic x4", lambda: ops.interp(xin, yout, B, 160, 160, 4, N.INTERP_BICUBIC, 160, 160), B * (160 * 160 + 640 * 640) * 4)

# ---- Track-2 trunk ----
conv_case("conv_tc 3x3 d5 54->54 (60->56 padded) @160", 60, 56, 3, 5, 160, real=(54, 54))
conv_case("conv_tc 1x1 54->54 (60->56 padded) @160 LReLU", 60, 56, 1, 1, 160, real=(54, 54), act=N.ACT_LRELU, slope=0.1)
conv_case("conv_tc 3x3 64->64 d5 @160 (DistgSSR / LF-InterNet)", 64, 64, 3, 5, 160, act=N.ACT_LRELU, slope=0.1)
conv_case("conv_tc 1x1 64->64 @160", 64, 64, 1, 1, 160)
x20 = torch.rand(B, 160, 160, 20, device=dev)
y20 = torch.empty(B, 160, 160, 20, device=dev)
w20 = (torch.rand(20, 18, 3, 3) - 0.5) * 0.1
pc20 = K.pack_conv(w20, dil=(5, 5), pad=(5, 5), device=dev, tc=True)
run("conv_thin 3x3 d5 18->20 @160", lambda: ops.conv(x20[..., :18], pc20, y20, act=N.ACT_LRELU, slope=0.1),
    B * 160 * 160 * 4 * 36, 2.0 * B * 160 * 160 * 18 * 18 * 9)
trunk = torch.rand(B, 160, 160, 60, device=dev)
cat = torch.zeros(B, 160, 160, 60, device=dev)
wep = (torch.rand((2 * 11 + 9) * 18 + 6 * 18 * 18, device=dev) - 0.5) * 0.3
run("mel_epi_branch @160", lambda: ops.mel_epi_branch(trunk[..., 40:58], wep, cat[..., 40:58], 11, 5, 0.1),
    B * 160 * 160 * 4 * 36, 2.0 * B * 160 * 160 * (18 * 31 + 6 * 18 * 18))
run("mel_epi_branch_tc @160 (1x1s on tcgen05)", lambda: ops.mel_epi_branch(trunk[..., 40:58], wep, cat[..., 40:58], 11, 5, 0.1, tc=True),
    B * 160 * 160 * 4 * 36, 2.0 * B * 160 * 160 * (18 * 31 + 6 * 18 * 18))
epi_img = ops.mel_epi_pack(wep, 11, dev)
trunk16 = K.alloc_nhwc16(B, 160, 160, 64, dev)
trunk16[..., :60].copy_(trunk)
run("mel_epi_branch_mma @160 (taps as shifted-row MMAs, opt-in)",
    lambda: ops.mel_epi_branch_mma(trunk[..., 40:58], trunk16[..., 40:58], epi_img, cat[..., 40:58], 11, 5, 0.1),
    B * 160 * 160 * (2 * 18 + 4 * 18), 2.0 * B * 160 * 160 * (18 * 31 + 6 * 18 * 18))
ang5 = torch.rand(B, 32, 32, 20, device=dev)
wex = (torch.rand(5, 5, 18, 20, device=dev) - 0.5) * 0.3
run("ang_expand 18->20 x5x5 + residual @32 -> @160", lambda: ops.ang_expand(ang5[..., :18], wex, trunk[..., 20:40], cat[..., 20:40], 5, N.ACT_LRELU, 0.1, 0.1),
    B * 160 * 160 * 4 * 36 + B * 32 * 32 * 72, 2.0 * B * 160 * 160 * 18 * 18)
res, out = torch.rand_like(trunk), torch.empty_like(trunk)
dw, bs, bb = torch.rand(9, 60, device=dev), torch.rand(60, device=dev) + 0.5, torch.rand(60, device=dev)
am = torch.rand(B, 5, 5, 60, device=dev)
run("sa_tile (SA modulator tail) c54 d5 @160", lambda: ops.sa_modulate(trunk, dw, bs, bb, am, 0.4, 0.6, res, out, 5),
    3 * B * 160 * 160 * 54 * 4)
xs16 = K.alloc_nhwc16(B, 160, 160, 32, dev)
run("sa_tile (SA modulator tail + fp16 copy of 32 channels)", lambda: ops.sa_modulate(trunk, dw, bs, bb, am, 0.4, 0.6, res, out, 5, out16=xs16),
    3 * B * 160 * 160 * 54 * 4 + B * 160 * 160 * 64)
vm = torch.empty(B, 5, 5, 60, device=dev)
run("block_mean 32x32 c54 @160", lambda: ops.block_mean(trunk, vm, 32, 32), B * 160 * 160 * 54 * 4)
x1 = torch.rand(B, 160, 160, 1, device=dev)
ws = (torch.rand(60, 1, 3, 3) - 0.5) * 0.3
pcs = K.pack_conv(ws, torch.rand(60) - 0.5, dil=(5, 5), pad=(5, 5), device=dev)
run("conv_stem 1->54 (60) 3x3 d5 @160", lambda: ops.conv(x1, pcs, trunk), B * 160 * 160 * 55 * 4)

# ---- EPIT tokens ----
E, heads = 128, 8
qk = torch.rand(B, 160, 160, 2 * E, device=dev)
vv = torch.rand(B, 160, 160, E, device=dev)
ao = torch.empty(B, 160, 160, E, device=dev)
run("epi_attention5 (A=5, S=32, 8 heads x 16)",
    lambda: ops.epi_attention(qk, vv, ao, heads, E // heads, 5, 32, 5, B, 5, 32, 32 * 160, 160, 160 * 160, 32, 1),
    B * 160 * 160 * 4 * 4 * E, 2.0 * 2 * B * 160 * 160 * 55 * E)
g_, b_ = torch.rand(E, device=dev), torch.rand(E, device=dev)
tok = lambda t: t.view(1, 1, -1, t.shape[3])
run("layernorm 128 @tokens", lambda: ops.layernorm(tok(vv), g_, b_, 1e-5, tok(ao)), 2 * B * 160 * 160 * E * 4)
print("done")
