"""MyEfficientLFNet v4.5 (FastConvSSM branch) on liblfsr_b200 kernels - mirror of
/root/reference/model/SR/MyEfficientLFNetV4_5.py as it runs when mamba_ssm is absent (:17-25), which is
the only way it runs in the reference's own environment list and in this image.

64-channel trunk at LR resolution, NHWC fp32. Per MambaLFBlock (:143-148), 11 launches:

  MultiScaleSpatial (:262-282)   one lfsr_dwconv_multi launch: identity / 3x3 / 5x5 / 7x7 depthwise branches
                                 on the four 16-channel slices (the slice-0 1x1 conv is composed into the
                                 pointwise conv at pack time), then lfsr_conv2d_tc 1x1 64->64 with folded
                                 BN + LReLU + residual
  FastConvSSM (:208-244)         BN folded into gate_conv (weights and a bias); ONE 1x1 64->128 conv
                                 with GELU writes [gate | y]; the four dilated depthwise 3x3 in one
                                 lfsr_dwconv_multi launch into the four windows of a 256-channel buffer
                                 (the torch.cat); fuse 1x1 256->64 with
                                 the `* silu(gate)` product in its epilogue (mul, mul_act); proj 1x1
                                 with alpha = scale and residual x
  fuse + ChannelAttention        1x1 128->64; two-level lfsr_block_mean; fc1/fc2 on the [B,1,1,C] means;
  (:285-299)                     lfsr_scale_add: out = fused * attn + x, written straight into this
                                 block's 64-channel window of the fuse_early / fuse_late input

so no torch.cat / chunk of the reference ever copies. Tail: fuse convs, refine 3x3, the PixelShuffle
upsampler with shuffle + LReLU fused into the conv epilogues, output conv added onto the bicubic skip.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native as N
from .. import kernels as K
from .common import LFNetBase, bn_affine, slots, tail_table, upsample_tail


def _c(cin, cout, k, bias=False, **kw):
    return nn.Conv2d(cin, cout, k, bias=bias, **kw)


class _LPE(nn.Module):          # LocalPixelEnhancement (:285-294)
    def __init__(self, ch):
        super().__init__()
        self.dw = _c(ch, ch, 3, padding=1, groups=ch)
        self.bn = nn.BatchNorm2d(ch)
        self.pw = _c(ch, ch, 1)


class _MultiScaleSpatial(nn.Module):
    def __init__(self, ch):
        super().__init__()
        c = ch // 4
        self.c = c
        self.conv1 = _c(c, c, 1)
        self.conv3 = _c(c, c, 3, padding=1, groups=c)
        self.conv5 = _c(c, c, 5, padding=2, groups=c)
        self.conv7 = _c(c, c, 7, padding=3, groups=c)
        self.pw = _c(ch, ch, 1)
        self.bn = nn.BatchNorm2d(ch)


class _FastConvSSM(nn.Module):
    DILS = (1, 2, 4, 8)

    def __init__(self, ch):
        super().__init__()
        self.norm = nn.BatchNorm2d(ch)
        self.gate_conv = _c(ch, 2 * ch, 1)
        for d in self.DILS:
            setattr(self, f"conv{d}", _c(ch, ch, 3, padding=d, dilation=d, groups=ch))
        self.fuse = _c(4 * ch, ch, 1)
        self.proj = _c(ch, ch, 1)
        self.scale = nn.Parameter(torch.ones(1) * 0.1)


class _ChannelAttention(nn.Module):
    def __init__(self, ch, reduction=8):
        super().__init__()
        hidden = max(ch // reduction, 16)
        self.fc1 = _c(ch, hidden, 1, bias=True)
        self.fc2 = _c(hidden, ch, 1, bias=True)


class _Block(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.ms_spatial = _MultiScaleSpatial(ch)
        self.ssm = _FastConvSSM(ch)
        self.fuse = _c(2 * ch, ch, 1)
        self.ca = _ChannelAttention(ch)


class _Upsampler(nn.Module):
    def __init__(self, ch, scale):
        super().__init__()
        self.steps = [(0, 2), (3, 2)] if scale == 4 else [(0, scale)]
        self.up = slots({i: _c(ch, ch * r * r, 3, padding=1) for i, r in self.steps})


class get_model(LFNetBase):
    def __init__(self, args):
        super().__init__(args)
        if getattr(args, "use_macpi", False):
            # the reference's entry points never set it (option.py has no such flag; :39 defaults it to False)
            raise N.LfsrError("MyEfficientLFNetV4_5: use_macpi=True is outside the accelerated path")
        self.channels = ch = 64
        self.n_blocks = 8
        self.shallow = slots({0: _c(1, ch, 3, padding=1), 2: _LPE(ch)})
        self.blocks = nn.ModuleList([_Block(ch) for _ in range(self.n_blocks)])
        self.fuse_early = _c(4 * ch, ch, 1)
        self.fuse_late = _c(4 * ch, ch, 1)
        self.fuse_final = _c(2 * ch, ch, 1)
        self.refine = _c(ch, ch, 3, padding=1)
        self.upsampler = _Upsampler(ch, self.scale)
        self.output = _c(ch, 1, 3, bias=True, padding=1)

    # -- pack ------------------------------------------------------------------------------------------
    def _pack(self, device, ops):
        pc = lambda w, b=None, **kw: K.pack_conv(w, b, device=device, **kw)
        f64 = lambda t: t.detach().double().cpu()
        dev_f = lambda t: t.float().contiguous().to(device)

        def dw_taps(w):                      # torch depthwise [c,1,kh,kw] -> [kh*kw][c]
            c, _, kh, kw = w.shape
            return w.detach().float().reshape(c, kh * kw).t().contiguous().to(device)

        lpe = self.shallow["2"]
        s, t = bn_affine(lpe.bn)
        pk = {"stem": pc(self.shallow["0"].weight, pad=(1, 1)),
              "lpe_dw": dw_taps(lpe.dw.weight), "lpe_s": dev_f(s), "lpe_t": dev_f(t),
              "lpe_pw": pc(lpe.pw.weight, tc=True), "blocks": []}
        for blk in self.blocks:
            ms, ssm = blk.ms_spatial, blk.ssm
            c = ms.c
            ch = 4 * c
            ms_br = [dict(w=torch.ones(1, c, dtype=torch.float32, device=device), kh=1, kw=1, in_c0=0, out_c0=0, c=c)]
            for j, cv in enumerate((ms.conv3, ms.conv5, ms.conv7), 1):
                k = cv.kernel_size[0]
                ms_br.append(dict(w=dw_taps(cv.weight), kh=k, kw=k, in_c0=j * c, out_c0=j * c, c=c))
            pw = f64(ms.pw.weight)[:, :, 0, 0].clone()                            # [out, in]
            pw[:, :c] = pw[:, :c] @ f64(ms.conv1.weight)[:, :, 0, 0]
            s, t = bn_affine(ms.bn)
            pw = pw * f64(s)[:, None]
            sn, tn = bn_affine(ssm.norm)
            wg = f64(ssm.gate_conv.weight)[:, :, 0, 0]
            b = dict(
                ms_dw=ms_br,
                ms_pw=pc(pw.float()[:, :, None, None], t.detach().float(), tc=True),
                gate=pc((wg * f64(sn)[None, :]).float()[:, :, None, None], (wg @ f64(tn)).float(), tc=True),
                dws=[dict(w=dw_taps(getattr(ssm, f"conv{d}").weight), kh=3, kw=3, dil=(d, d), in_c0=ch, out_c0=k * ch, c=ch)
                     for k, d in enumerate(ssm.DILS)],
                ssm_fuse=pc(ssm.fuse.weight, tc=True), proj=pc(ssm.proj.weight, tc=True),
                ssm_scale=float(ssm.scale.detach().float().cpu()),
                fuse=pc(blk.fuse.weight, tc=True),
                fc1=pc(blk.ca.fc1.weight, blk.ca.fc1.bias), fc2=pc(blk.ca.fc2.weight, blk.ca.fc2.bias))
            pk["blocks"].append(b)
        pk["fuse_early"] = pc(self.fuse_early.weight, tc=True)
        pk["fuse_late"] = pc(self.fuse_late.weight, tc=True)
        pk["fuse_final"] = pc(self.fuse_final.weight, tc=True)
        pk["refine"] = pc(self.refine.weight, pad=(1, 1), tc=True)
        pk["up"] = [(pc(self.upsampler.up[str(i)].weight, pad=(1, 1), tc=True, tc_shuffle=(r, r, N.SHUF_CHANNEL_MAJOR)), r)
                    for i, r in self.upsampler.steps]
        pk["out"] = pc(self.output.weight, self.output.bias, pad=(1, 1))
        pk["tail_w"] = tail_table(self.output.weight, self.channels, device)
        return pk

    # -- run ---------------------------------------------------------------------------------------------
    def _run(self, ops, pk, x, out):
        A, s, C = self.angRes, self.scale, self.channels
        B, _, H, W = x.shape
        dev = x.device
        buf = lambda name, h, w, c: self._buf(name, B, h, w, c, dev)
        LR = N.ACT_LRELU
        xin = x.view(B, H, W, 1)
        Y = out.view(B, H * s, W * s, 1)
        ops.interp(x, out, B, H, W, s, N.INTERP_BICUBIC, H, W)

        f0, f1, shallow = buf("f0", H, W, C), buf("f1", H, W, C), buf("shallow", H, W, C)
        ops.conv(xin, pk["stem"], f0, act=LR, slope=0.1)
        ops.dwconv(f0, pk["lpe_dw"], f1, 3, 3, scale=pk["lpe_s"], shift=pk["lpe_t"], act=LR, slope=0.1)
        ops.conv(f1, pk["lpe_pw"], shallow, res=f0)

        nb = len(pk["blocks"])
        half = (nb + 1) // 2
        early, late = buf("early", H, W, half * C), buf("late", H, W, max(nb - half, 1) * C)
        ms, cat2, g, cat4 = buf("ms", H, W, C), buf("cat2", H, W, 2 * C), buf("g", H, W, 2 * C), buf("cat4", H, W, 4 * C)
        yf, fused = buf("yf", H, W, C), buf("fused", H, W, C)
        hA, wA = H // A, W // A
        vmean, gmean = buf("vmean", A, A, C), buf("gmean", 1, 1, C)
        hid = pk["blocks"][0]["fc1"].cout if nb else 1
        ca1, attn = buf("ca1", 1, 1, hid), buf("attn", 1, 1, C)
        feat = shallow
        for i, b in enumerate(pk["blocks"]):
            # multi-scale spatial -> f_local = cat2[..., :C]
            ops.dwconv_multi(feat, ms, b["ms_dw"])
            ops.conv(ms, b["ms_pw"], cat2[..., 0:C], act=LR, slope=0.1, res=feat)
            # gated multi-dilation branch -> f_global = cat2[..., C:]
            ops.conv(feat, b["gate"], g, act=N.ACT_GELU)
            ops.dwconv_multi(g, cat4, b["dws"])
            ops.conv(cat4, b["ssm_fuse"], yf, mul=g[..., 0:C], mul_act=N.ACT_SILU)
            ops.conv(yf, b["proj"], cat2[..., C:2 * C], alpha=b["ssm_scale"], res=feat)
            # fuse + channel attention + block residual
            ops.conv(cat2, b["fuse"], fused)
            ops.block_mean(fused, vmean, hA, wA)
            ops.block_mean(vmean, gmean, A, A)
            ops.conv(gmean, b["fc1"], ca1, act=N.ACT_RELU)
            ops.conv(ca1, b["fc2"], attn, act=N.ACT_SIGMOID)
            dst = early if i < half else late
            j = i if i < half else i - half
            nxt = dst[..., j * C:(j + 1) * C]
            ops.scale_add(fused, attn, feat, nxt)
            feat = nxt
        el = cat2
        ops.conv(early, pk["fuse_early"], el[..., 0:C])
        ops.conv(late, pk["fuse_late"], el[..., C:2 * C])
        ops.conv(el, pk["fuse_final"], f0, res=shallow)
        ops.conv(f0, pk["refine"], f1, act=LR, slope=0.1)
        upsample_tail(self, ops, pk, f1, H, W, Y, C, LR, N.SHUF_CHANNEL_MAJOR)


class get_loss(nn.Module):
    """L1 + 0.05 * L1 on rfft2 magnitudes (MyEfficientLFNetV4_5.py:326-337); training itself is out of scope."""

    def __init__(self, args=None):
        super().__init__()
        self.l1 = nn.L1Loss()
        self.fft_weight = 0.05

    def forward(self, SR, HR, criterion_data=None):
        spec = F.l1_loss(torch.abs(torch.fft.rfft2(SR)), torch.abs(torch.fft.rfft2(HR)))
        return self.l1(SR, HR) + self.fft_weight * spec


def weights_init(m):
    """MyEfficientLFNetV4_5.py:340-344."""
    if isinstance(m, (nn.Conv2d, nn.Linear)):
        nn.init.kaiming_normal_(m.weight, a=0.1, mode="fan_in", nonlinearity="leaky_relu")
        if m.bias is not None:
            nn.init.zeros_(m.bias)
