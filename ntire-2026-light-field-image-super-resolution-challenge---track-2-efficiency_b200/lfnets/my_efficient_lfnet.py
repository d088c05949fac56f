"""Track-2 model `MyEfficientLFNet` (v2.0) on liblfsr_b200 kernels.

Mirror of /root/reference/model/SR/MyEfficientLFNet.py: same `get_model(args)`, `get_loss(args)`,
`weights_init(m)` symbols and the same state_dict (279 entries, 547 540 parameters at x4), but the
eval forward is a fixed launch plan over NHWC fp32 buffers:

  bicubic(mosaic) -> Y                                       lfsr_interp      (:88-90)
  RepConv stem 1->54 (3 branches + BN folded to one conv)    lfsr_conv2d_*    (:40-43, :374-385)
  5 x ProgressiveDisentanglingStage                          (:183-208)
      spatial  RepConv(18) LReLU conv3x3 d5                  -> cat[0:18]
      angular  5x5/s5, gate chain, 1x1->450 + PixelShuffle(5) fused store, x+scale*out -> cat[18:36]
      EPI      dw 1x11 / 11x1 / 3x3 d5 + 1x1, fuse 1x1       -> cat[36:54]
      gates    per-view means -> global means -> block-diagonal FC+sigmoid -> folded into the
               fusion 1x1 as a per-sample input-channel scale
      fusion   1x1 LReLU, 3x3 d5 ; SAModulator tail + stage residual in one kernel
  global fusion 1x1 LReLU, RepConv(54) + shallow residual
  upsampler  [3x3 54->216 + PixelShuffle(2) + LReLU fused] x2 (:555-565)
  3x3 54->1 + bias + bicubic residual (in place on Y)        (:70-73, :105)

torch.cat / torch.split never materialise: branches read and write channel slices of one buffer.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native as N
from .. import kernels as K
from .common import LFNetBase, bn_affine, slots


def _conv(cin, cout, k, **kw):
    return nn.Conv2d(cin, cout, k, **kw)


class _RepConvParams(nn.Module):
    """parameter holder with RepConvBlock's names (MyEfficientLFNet.py:330-372)."""

    def __init__(self, cin, cout, dil):
        super().__init__()
        self.conv_3x3 = _conv(cin, cout, 3, padding=dil, dilation=dil, bias=False)
        self.bn_3x3 = nn.BatchNorm2d(cout)
        self.conv_1x1 = _conv(cin, cout, 1, bias=False)
        self.bn_1x1 = nn.BatchNorm2d(cout)
        self.bn_identity = nn.BatchNorm2d(cout) if cin == cout else None

    def merged(self):
        """One 3x3 kernel + bias equal to bn(conv3x3)+bn(conv1x1)[+bn(x)] in eval mode. (The
        reference's own switch_to_deploy is broken - SURVEY 7f - so this is derived, not copied.)"""
        s3, b3 = bn_affine(self.bn_3x3)
        s1, b1 = bn_affine(self.bn_1x1)
        w = self.conv_3x3.weight.detach() * s3.view(-1, 1, 1, 1)
        w = w.clone()
        w[:, :, 1, 1] += self.conv_1x1.weight.detach()[:, :, 0, 0] * s1.view(-1, 1)
        b = b3 + b1
        if self.bn_identity is not None:
            si, bi = bn_affine(self.bn_identity)
            idx = torch.arange(w.shape[0], device=w.device)
            w[idx, idx, 1, 1] += si
            b = b + bi
        return w, b


class _AngularParams(nn.Module):
    def __init__(self, c, A):
        super().__init__()
        hidden = max(c // 4, 16)
        self.scale = nn.Parameter(torch.ones(1) * 0.1)
        self.to_angular = _conv(c, c, A, stride=A, bias=False)
        self.attention = slots({0: _conv(c, hidden, 1, bias=False),
                                2: _conv(hidden, hidden, 3, padding=1, groups=hidden, bias=False),
                                4: _conv(hidden, c, 1, bias=False)})
        self.cross_view = slots({0: _conv(c, c, 3, padding=1, bias=False)})
        self.expand = slots({0: _conv(c, c * A * A, 1, bias=False)})


class _EpiParams(nn.Module):
    def __init__(self, c, A):
        super().__init__()
        k = 2 * A + 1
        self.epi_h = slots({0: _conv(c, c, (1, k), padding=(0, A), groups=c, bias=False), 1: _conv(c, c, 1, bias=False)})
        self.epi_v = slots({0: _conv(c, c, (k, 1), padding=(A, 0), groups=c, bias=False), 1: _conv(c, c, 1, bias=False)})
        self.epi_diag = slots({0: _conv(c, c, 3, padding=A, dilation=A, groups=c, bias=False),
                               1: _conv(c, c, 1, bias=False)})
        self.fuse = slots({0: _conv(3 * c, c, 1, bias=False)})


class _SAModParams(nn.Module):
    def __init__(self, c, A):
        super().__init__()
        self.combine = nn.Parameter(torch.ones(2) * 0.5)
        self.spatial_mod = slots({0: _conv(c, c, 3, padding=A, dilation=A, groups=c, bias=False), 1: nn.BatchNorm2d(c)})
        self.angular_conv = slots({0: _conv(c, c // 4, 1, bias=False), 2: _conv(c // 4, c, 1, bias=False)})


class _StageParams(nn.Module):
    def __init__(self, c, A):
        super().__init__()
        self.split = [c // 3, c // 3, c - 2 * (c // 3)]
        s0, s1, s2 = self.split
        self.spatial_branch = slots({0: _RepConvParams(s0, s0, A), 2: _conv(s0, s0, 3, padding=A, dilation=A, bias=False)})
        self.angular_branch = _AngularParams(s1, A)
        self.epi_branch = _EpiParams(s2, A)
        self.gate_spatial = slots({1: _conv(s0, s0, 1, bias=True)})
        self.gate_angular = slots({1: _conv(s1, s1, 1, bias=True)})
        self.gate_epi = slots({1: _conv(s2, s2, 1, bias=True)})
        self.fusion = slots({0: _conv(c, c, 1, bias=False), 2: _conv(c, c, 3, padding=A, dilation=A, bias=False)})
        self.sa_modulator = _SAModParams(c, A)


class _UpsamplerParams(nn.Module):
    def __init__(self, c, scale):
        super().__init__()
        if scale == 4:
            self.up = slots({0: _conv(c, 4 * c, 3, padding=1, bias=False), 3: _conv(c, 4 * c, 3, padding=1, bias=False)})
            self.steps = [(0, 2), (3, 2)]
        elif scale == 2:
            self.up = slots({0: _conv(c, 4 * c, 3, padding=1, bias=False)})
            self.steps = [(0, 2)]
        else:
            self.up = slots({0: _conv(c, c * scale * scale, 3, padding=1, bias=False)})
            self.steps = [(0, scale)]


def _dw_pack(w, device):
    """depthwise weight [C,1,kh,kw] -> [kh*kw, C]"""
    c, _, kh, kw = w.shape
    return w.detach().reshape(c, kh * kw).t().contiguous().to(device=device, dtype=torch.float32)


class get_model(LFNetBase):
    def __init__(self, args):
        super().__init__(args)
        A = self.angRes
        self.channels = 54
        self.n_stages = 5
        c = self.channels
        self.shallow_feat = _RepConvParams(1, c, A)
        self.stages = nn.ModuleList([_StageParams(c, A) for _ in range(self.n_stages)])
        self.global_fusion = slots({0: _conv(c, c, 1, bias=False), 2: _RepConvParams(c, c, A)})
        self.upsampler = _UpsamplerParams(c, self.scale)
        self.output_conv = _conv(c, 1, 3, stride=1, padding=1, bias=True)

    # -- pack ---------------------------------------------------------------------------------------
    def _pack(self, device, ops):
        A = self.angRes
        pc = lambda w, b=None, **kw: K.pack_conv(w, b, device=device, **kw)
        dil = dict(dil=(A, A), pad=(A, A))
        pk = {}
        w, b = self.shallow_feat.merged()
        pk["stem"] = pc(w, b, **dil)
        stages = []
        for st in self.stages:
            s = {}
            w, b = st.spatial_branch["0"].merged()
            s["spa0"] = pc(w, b, **dil)
            s["spa2"] = pc(st.spatial_branch["2"].weight, **dil)
            ab = st.angular_branch
            s["ang_to"] = pc(ab.to_angular.weight, stride=(A, A))
            s["ang_a0"] = pc(ab.attention["0"].weight)
            s["ang_a2"] = _dw_pack(ab.attention["2"].weight, device)
            s["ang_a4"] = pc(ab.attention["4"].weight)
            s["ang_cv"] = pc(ab.cross_view["0"].weight, pad=(1, 1))
            s["ang_ex"] = pc(ab.expand["0"].weight)
            s["ang_scale"] = float(ab.scale.detach().item())
            eb = st.epi_branch
            s["epi_h_dw"] = _dw_pack(eb.epi_h["0"].weight, device)
            s["epi_v_dw"] = _dw_pack(eb.epi_v["0"].weight, device)
            s["epi_d_dw"] = _dw_pack(eb.epi_diag["0"].weight, device)
            s["epi_h_pw"] = pc(eb.epi_h["1"].weight)
            s["epi_v_pw"] = pc(eb.epi_v["1"].weight)
            s["epi_d_pw"] = pc(eb.epi_diag["1"].weight)
            s["epi_fuse"] = pc(eb.fuse["0"].weight)
            # three gate FCs as one block-diagonal 1x1 (MyEfficientLFNet.py:159-173)
            sp = st.split
            C = sum(sp)
            gw = torch.zeros(C, C, 1, 1)
            gb = torch.zeros(C)
            o = 0
            for g, n in zip((st.gate_spatial, st.gate_angular, st.gate_epi), sp):
                gw[o:o + n, o:o + n] = g["1"].weight.detach().cpu()
                gb[o:o + n] = g["1"].bias.detach().cpu()
                o += n
            s["gate"] = pc(gw, gb)
            s["fus0"] = pc(st.fusion["0"].weight)
            s["fus2"] = pc(st.fusion["2"].weight, tc=True, **dil)
            sm = st.sa_modulator
            s["sa_dw"] = _dw_pack(sm.spatial_mod["0"].weight, device)
            sc, sh = bn_affine(sm.spatial_mod["1"])
            s["sa_bns"] = sc.to(device=device, dtype=torch.float32).contiguous()
            s["sa_bnb"] = sh.to(device=device, dtype=torch.float32).contiguous()
            s["sa_c0"] = pc(sm.angular_conv["0"].weight)
            s["sa_c2"] = pc(sm.angular_conv["2"].weight)
            wts = torch.softmax(sm.combine.detach().float(), dim=0)
            s["sa_w"] = (float(wts[0]), float(wts[1]))
            stages.append(s)
        pk["stages"] = stages
        pk["gf0"] = pc(self.global_fusion["0"].weight)
        w, b = self.global_fusion["2"].merged()
        pk["gf2"] = pc(w, b, tc=True, **dil)
        pk["up"] = [(pc(self.upsampler.up[str(i)].weight, pad=(1, 1), tc=True, tc_shuffle=(r, r, N.SHUF_CHANNEL_MAJOR)), r)
                    for i, r in self.upsampler.steps]
        pk["out"] = pc(self.output_conv.weight, self.output_conv.bias, pad=(1, 1), tc=True)
        return pk

    # -- run ------------------------------------------------------------------------------------------
    def _run(self, ops, pk, x, out):
        A, s, C = self.angRes, self.scale, self.channels
        B, _, H, W = x.shape
        dev = x.device
        buf = lambda name, h, w, c: self._buf(name, B, h, w, c, dev)
        LR = N.ACT_LRELU
        xin = x.view(B, H, W, 1)
        Y = out.view(B, H * s, W * s, 1)
        ops.interp(x, out, B, H, W, s, N.INTERP_BICUBIC, H, W)

        shallow = buf("shallow", H, W, C)
        ops.conv(xin, pk["stem"], shallow)
        feat = shallow
        pp = [buf("feat_a", H, W, C), buf("feat_b", H, W, C)]
        c0 = C // 3
        hA, wA = H // A, W // A
        cat = buf("cat", H, W, C)
        t18 = buf("t18", H, W, c0)
        ecat = buf("ecat", H, W, 3 * (C - 2 * c0))
        hid = pk["stages"][0]["ang_a0"].cout
        ang1, ang4, ang5 = buf("ang1", hA, wA, c0), buf("ang4", hA, wA, c0), buf("ang5", hA, wA, c0)
        ang2, ang3 = buf("ang2", hA, wA, hid), buf("ang3", hA, wA, hid)
        vmean, gmean, gate = buf("vmean", A, A, C), buf("gmean", 1, 1, C), buf("gate", 1, 1, C)
        fu1, fu2 = buf("fu1", H, W, C), buf("fu2", H, W, C)
        pm, am1, am = buf("pm", A, A, C), buf("am1", A, A, C // 4), buf("am", A, A, C)
        for i, sp in enumerate(pk["stages"]):
            xs, xa, xe = feat[..., 0:c0], feat[..., c0:2 * c0], feat[..., 2 * c0:C]
            ce = C - 2 * c0
            # spatial branch
            ops.conv(xs, sp["spa0"], t18, act=LR, slope=0.1)
            ops.conv(t18, sp["spa2"], cat[..., 0:c0])
            # angular branch
            ops.conv(xa, sp["ang_to"], ang1)
            ops.conv(ang1, sp["ang_a0"], ang2, act=N.ACT_RELU)
            ops.dwconv(ang2, sp["ang_a2"], ang3, 3, 3, act=N.ACT_RELU)
            ops.conv(ang3, sp["ang_a4"], ang4, act=N.ACT_SIGMOID, mul=ang1)
            ops.conv(ang4, sp["ang_cv"], ang5, act=LR, slope=0.1)
            ops.conv(ang5, sp["ang_ex"], cat[..., c0:2 * c0], act=LR, slope=0.1, alpha=sp["ang_scale"], res=xa,
                     shuffle=(A, A, N.SHUF_CHANNEL_MAJOR))
            # EPI branch
            te = t18[..., 0:ce] if ce == c0 else buf("te", H, W, ce)
            ops.dwconv(xe, sp["epi_h_dw"], te, 1, 2 * A + 1)
            ops.conv(te, sp["epi_h_pw"], ecat[..., 0:ce], act=LR, slope=0.1)
            ops.dwconv(xe, sp["epi_v_dw"], te, 2 * A + 1, 1)
            ops.conv(te, sp["epi_v_pw"], ecat[..., ce:2 * ce], act=LR, slope=0.1)
            ops.dwconv(xe, sp["epi_d_dw"], te, 3, 3, dil=(A, A))
            ops.conv(te, sp["epi_d_pw"], ecat[..., 2 * ce:3 * ce], act=LR, slope=0.1)
            ops.conv(ecat, sp["epi_fuse"], cat[..., 2 * c0:C], act=LR, slope=0.1)
            # gates -> per-sample channel scale of the fusion 1x1
            ops.block_mean(cat, vmean, hA, wA)
            ops.block_mean(vmean, gmean, A, A)
            ops.conv(gmean, sp["gate"], gate, act=N.ACT_SIGMOID)
            ops.conv(cat, sp["fus0"], fu1, act=LR, slope=0.1, in_scale=gate)
            ops.conv(fu1, sp["fus2"], fu2)
            # SA modulator + stage residual
            ops.block_mean(fu2, pm, hA, wA)
            ops.conv(pm, sp["sa_c0"], am1, act=N.ACT_RELU)
            ops.conv(am1, sp["sa_c2"], am, act=N.ACT_SIGMOID)
            nxt = pp[i & 1]
            ops.sa_modulate(fu2, sp["sa_dw"], sp["sa_bns"], sp["sa_bnb"], am, sp["sa_w"][0], sp["sa_w"][1], feat, nxt, A)
            feat = nxt
        ops.conv(feat, pk["gf0"], fu1, act=LR, slope=0.1)
        ops.conv(fu1, pk["gf2"], fu2, res=shallow)
        cur, ch, cw = fu2, H, W
        for j, (pcv, r) in enumerate(pk["up"]):
            nb = buf(f"up{j}", ch * r, cw * r, C)
            ops.conv(cur, pcv, nb, act=LR, slope=0.1, shuffle=(r, r, N.SHUF_CHANNEL_MAJOR))
            cur, ch, cw = nb, ch * r, cw * r
        ops.conv(cur, pk["out"], Y, res=Y)


    # -- measurement hook --------------------------------------------------------------------------
    def dominant_kernel(self, batch: int, h: int = 32):
        """(callable, info) for the layer that dominates HBM bytes and FLOPs: the last upsampler
        conv 3x3 C->4C + PixelShuffle(2) + LReLU (MyEfficientLFNet.py:560-565): 55 % of the MACs,
        reads C x (2H)^2, writes C x (4H)^2 per patch. Used by bench.py's roofline entry."""
        dev = next(self.parameters()).device
        ops = K.default_ops()
        pk = self._get_packed(dev, ops)
        A, C = self.angRes, self.channels
        pcv, r = pk["up"][-1]
        hin = A * h * (self.scale // r)
        src = self._buf(f"up{len(pk['up']) - 2}", batch, hin, hin, C, dev) if len(pk["up"]) > 1 else \
            self._buf("fu2", batch, hin, hin, C, dev)
        dst = self._buf(f"up{len(pk['up']) - 1}", batch, hin * r, hin * r, C, dev)
        info = {
            "name": "conv3x3 %d->%d + PixelShuffle(%d) + LReLU @%dx%d (upsampler.up.%s)" % (C, pcv.cout, r, hin, hin,
                                                                                        self.upsampler.steps[-1][0]),
            "bytes": batch * (hin * hin * C + hin * r * hin * r * C) * 4 + pcv.w_f32.numel() * 4,
            "flops": 2 * batch * hin * hin * pcv.kh * pcv.kw * pcv.cin * pcv.cout,
        }
        return (lambda: ops.conv(src, pcv, dst, act=N.ACT_LRELU, slope=0.1, shuffle=(r, r, N.SHUF_CHANNEL_MAJOR))), info


class get_loss(nn.Module):
    """L1 + 0.05 * L1 of rFFT magnitudes (MyEfficientLFNet.py:585-609); training only."""

    def __init__(self, args=None):
        super().__init__()
        self.l1_loss = nn.L1Loss()
        self.use_freq = True
        self.freq_weight = 0.05

    def forward(self, SR, HR, criterion_data=None):
        loss = self.l1_loss(SR, HR)
        if self.use_freq:
            loss = loss + self.freq_weight * nn.functional.l1_loss(torch.fft.rfft2(SR).abs(), torch.fft.rfft2(HR).abs())
        return loss


def weights_init(m):
    """Kaiming-normal (a=0.1, fan_in) on conv/linear, unit BatchNorm (MyEfficientLFNet.py:612-624)."""
    if isinstance(m, (nn.Conv2d, nn.Linear)):
        nn.init.kaiming_normal_(m.weight, a=0.1, mode="fan_in", nonlinearity="leaky_relu")
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)
