"""DistgSSR on liblfsr_b200 kernels - mirror of /root/reference/model/SR/DistgSSR.py.

Features live in HBM in the MacPI arrangement as NHWC fp32; SAI2MacPI (:145-155) is folded into
the stem conv's load addressing and MacPI2SAI (:134-142) into the upsampler's store addressing,
so the 1024-slice Python loops of the reference disappear. Per DisentgBlock (:73-111):

  Spa   2 x [3x3 d=A + LReLU]                                  -> cat[0:64]
  Ang   AxA/sA 64->16 + LReLU, 1x1 16->16A^2 + LReLU + PixelShuffle(A) fused store -> cat[64:80]
  EPI-H 1xA^2 / s(1,A) 64->32 + LReLU, 1x1 32->32A + LReLU + PixelShuffle1D(A)     -> cat[80:112]
  EPI-V the same weights as an A^2x1 / s(A,1) conv + vertical PixelShuffle1D (replaces the two
        explicit transposes of :107)                                               -> cat[112:144]
  fuse  1x1 144->64 + LReLU, 3x3 d=A, + x

The tail `1x1 64->64s^2 (+bias) -> PixelShuffle(s) -> 1x1 64->1` (:24-27) contains no
non-linearity, so it is composed at pack time into one 1x1 64->s^2 conv whose PixelShuffle store
lands directly on the bilinear residual (:30,35) - the 64 x (sH) x (sW) intermediate never exists.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native as N
from .. import kernels as K
from .common import LFNetBase, L1Loss, slots


def _c(cin, cout, k, **kw):
    return nn.Conv2d(cin, cout, k, bias=kw.pop("bias", False), **kw)


class _Block(nn.Module):
    def __init__(self, A, ch):
        super().__init__()
        spa, ang, epi = ch, ch // 4, ch // 2
        d = dict(dilation=A, padding=A)
        self.SpaConv = slots({0: _c(ch, spa, 3, **d), 2: _c(spa, spa, 3, **d)})
        self.AngConv = slots({0: _c(ch, ang, A, stride=A), 2: _c(ang, A * A * ang, 1)})
        self.EPIConv = slots({0: _c(ch, epi, (1, A * A), stride=(1, A), padding=(0, A * (A - 1) // 2)),
                              2: _c(epi, A * epi, 1)})
        self.fuse = slots({0: _c(spa + ang + 2 * epi, ch, 1), 2: _c(ch, ch, 3, **d)})


class _Group(nn.Module):
    def __init__(self, n_block, A, ch):
        super().__init__()
        self.Block = nn.ModuleList([_Block(A, ch) for _ in range(n_block)])
        self.conv = _c(ch, ch, 3, dilation=A, padding=A)


class _Cascade(nn.Module):
    def __init__(self, n_group, n_block, A, ch):
        super().__init__()
        self.Group = nn.ModuleList([_Group(n_block, A, ch) for _ in range(n_group)])
        self.conv = _c(ch, ch, 3, dilation=A, padding=A)


class get_model(LFNetBase):
    def __init__(self, args):
        super().__init__(args)
        A, ch = self.angRes, 64
        self.channels = ch
        self.factor = self.scale
        self.init_conv = _c(1, ch, 3, dilation=A, padding=A)
        self.disentg = _Cascade(4, 4, A, ch)
        self.upsample = slots({0: _c(ch, ch * self.scale ** 2, 1, bias=True), 2: _c(ch, 1, 1)})

    def _pack(self, device, ops):
        A, s = self.angRes, self.scale
        f16 = self._fp16(ops)
        pc = lambda w, b=None, **kw: K.pack_conv(w, b, device=device, **(dict(kw, tc16=True) if (f16 and kw.get("tc")) else kw))
        dil = dict(dil=(A, A), pad=(A, A))
        pk = {"stem": pc(self.init_conv.weight, **dil), "groups": [], "f16": f16}
        for g in self.disentg.Group:
            blocks = []
            for b in g.Block:
                we = b.EPIConv["0"].weight
                blocks.append(dict(
                    spa0=pc(b.SpaConv["0"].weight, tc=True, **dil), spa2=pc(b.SpaConv["2"].weight, tc=True, **dil),
                    ang0=pc(b.AngConv["0"].weight, stride=(A, A), tc=True),
                    ang2=pc(b.AngConv["2"].weight, tc=True, tc_shuffle=(A, A, N.SHUF_CHANNEL_MAJOR)),
                    epi_h=pc(we, stride=(1, A), pad=(0, A * (A - 1) // 2), tc=True),
                    epi_v=pc(we.permute(0, 1, 3, 2), stride=(A, 1), pad=(A * (A - 1) // 2, 0), tc=True),
                    epi2=pc(b.EPIConv["2"].weight, tc=True, tc_shuffle=(1, A, N.SHUF_FACTOR_MAJOR)),
                    fuse0=pc(b.fuse["0"].weight, tc=True), fuse2=pc(b.fuse["2"].weight, tc=True, **dil)))
            pk["groups"].append(dict(blocks=blocks, conv=pc(g.conv.weight, tc=True, **dil)))
        pk["cascade"] = pc(self.disentg.conv.weight, tc=True, **dil)
        # compose 1x1(64->64 s^2, bias) . PixelShuffle(s) . 1x1(64->1)  ==  1x1(64->s^2) . PixelShuffle(s)
        w1 = self.upsample["0"].weight.detach().double()[:, :, 0, 0]              # [64 s^2, 64]
        b1 = self.upsample["0"].bias.detach().double()
        w2 = self.upsample["2"].weight.detach().double()[0, :, 0, 0]              # [64]
        ch = w2.numel()
        w_eff = torch.einsum("c,cri->ri", w2, w1.view(ch, s * s, -1))             # [s^2, 64]
        b_eff = torch.einsum("c,cr->r", w2, b1.view(ch, s * s))
        pk["tail"] = pc(w_eff.float().view(s * s, -1, 1, 1), b_eff.float())
        return pk

    @staticmethod
    def _fp16(ops):
        from . import common
        return bool(common.USE_FP16_OPERANDS and getattr(ops, "fp16_operands", getattr(ops, "use_tc", False)) and hasattr(ops, "to_f16"))

    def _run(self, ops, pk, x, out):
        if pk.get("f16"):
            return self._run16(ops, pk, x, out)
        A, s, C = self.angRes, self.scale, self.channels
        B, _, H, W = x.shape
        dev = x.device
        buf = lambda name, h, w, c: self._buf(name, B, h, w, c, dev)
        LR = N.ACT_LRELU
        xin = x.view(B, H, W, 1)
        Y = out.view(B, H * s, W * s, 1)
        ops.interp(x, out, B, H, W, s, N.INTERP_BILINEAR, H, W)
        hA, wA = H // A, W // A
        ang_c, epi_c = C // 4, C // 2
        buf0 = buf("buf0", H, W, C)
        ops.conv(xin, pk["stem"], buf0, in_perm=N.PERM_MACPI_OVER_SAI, perm_a=A)
        cat = buf("cat", H, W, C + ang_c + 2 * epi_c)
        spa1, f1 = buf("spa1", H, W, C), buf("f1", H, W, C)
        ang1 = buf("ang1", hA, wA, ang_c)
        eh, ev = buf("eh", H, wA, epi_c), buf("ev", hA, W, epi_c)
        ring = [buf(f"ring{i}", H, W, C) for i in range(3)]
        cur, nring = buf0, 0
        o1, o2, o3 = C, C + ang_c, C + ang_c + epi_c
        for g in pk["groups"]:
            gin = cur
            for b in g["blocks"]:
                ops.conv(cur, b["spa0"], spa1, act=LR, slope=0.1)
                ops.conv(spa1, b["spa2"], cat[..., 0:o1], act=LR, slope=0.1)
                ops.conv(cur, b["ang0"], ang1, act=LR, slope=0.1)
                ops.conv(ang1, b["ang2"], cat[..., o1:o2], act=LR, slope=0.1, shuffle=(A, A, N.SHUF_CHANNEL_MAJOR))
                ops.conv(cur, b["epi_h"], eh, act=LR, slope=0.1)
                ops.conv(eh, b["epi2"], cat[..., o2:o3], act=LR, slope=0.1, shuffle=(1, A, N.SHUF_FACTOR_MAJOR))
                ops.conv(cur, b["epi_v"], ev, act=LR, slope=0.1)
                ops.conv(ev, b["epi2"], cat[..., o3:o3 + epi_c], act=LR, slope=0.1, shuffle=(A, 1, N.SHUF_FACTOR_MAJOR))
                ops.conv(cat, b["fuse0"], f1, act=LR, slope=0.1)
                nxt = ring[nring]
                nring = (nring + 1) % 3
                if nxt is gin:              # never overwrite the group input that is still needed
                    nxt = ring[nring]
                    nring = (nring + 1) % 3
                ops.conv(f1, b["fuse2"], nxt, res=cur)
                cur = nxt
            nxt = ring[nring]
            nring = (nring + 1) % 3
            if nxt is gin:
                nxt = ring[nring]
                nring = (nring + 1) % 3
            ops.conv(cur, g["conv"], nxt, res=gin)
            cur = nxt
        ops.conv(cur, pk["cascade"], spa1, res=buf0)
        self._tail(ops, pk, spa1, out, B, H, W)

    def _tail(self, ops, pk, last, out, B, H, W):
        """composed upsampler 1x1(64 -> s^2) on the tensor cores (MacPI arrangement), then MacPI->SAI + PixelShuffle(s) added
        onto the bilinear skip already in `out`"""
        A, s = self.angRes, self.scale
        # fp32 1x1 to s^2 channels in the MacPI arrangement (lfsr_conv1x1_few), then the un-shuffle kernel adds it onto the skip:
        # 0.6 -> ~0.2 ms against the general fp32 conv with the MacPI -> SAI + PixelShuffle store fused (no rounding involved)
        if hasattr(ops, "macpi_unshuffle") and s in (2, 4) and pk["tail"].w_tc is None:
            rec = self._buf("recon", B, H, W, s * s, last.device)
            ops.conv(last, pk["tail"], rec)
            ops.macpi_unshuffle(rec, out, A, s, True)
        else:
            Y = out.view(B, H * s, W * s, 1)
            ops.conv(last, pk["tail"], Y, res=Y, out_perm=N.PERM_MACPI_OVER_SAI, perm_a=A, shuffle=(s, s, N.SHUF_CHANNEL_MAJOR))


    def _run16(self, ops, pk, x, out):
        """the same launch plan with fp16 activations between the tensor-core layers: every intermediate of a DisentgBlock
        (spa1, the 144-channel concat, ang1, eh / ev, f1) only ever feeds another convolution and exists in fp16 alone; the
        residual trunk (`cur`) stays fp32 and carries an fp16 copy for the four convolutions that read it."""
        A, s, C = self.angRes, self.scale, self.channels
        B, _, H, W = x.shape
        dev = x.device
        buf = lambda name, h, w, c: self._buf(name, B, h, w, c, dev)
        b16 = lambda name, h, w, c: self._buf16(name, B, h, w, c, dev)
        LR = N.ACT_LRELU
        xin = x.view(B, H, W, 1)
        Y = out.view(B, H * s, W * s, 1)
        ops.interp(x, out, B, H, W, s, N.INTERP_BILINEAR, H, W)
        hA, wA = H // A, W // A
        ang_c, epi_c = C // 4, C // 2
        buf0, buf0h = buf("buf0", H, W, C), b16("buf0", H, W, C)
        ops.conv(xin, pk["stem"], buf0, in_perm=N.PERM_MACPI_OVER_SAI, perm_a=A)
        ops.to_f16(buf0, buf0h)
        cat = b16("cat", H, W, C + ang_c + 2 * epi_c)
        spa1, f1 = b16("spa1", H, W, C), b16("f1", H, W, C)
        ang1 = b16("ang1", hA, wA, ang_c)
        eh, ev = b16("eh", H, wA, epi_c), b16("ev", hA, W, epi_c)
        ring = [(buf(f"ring{i}", H, W, C), b16(f"ring{i}", H, W, C)) for i in range(3)]
        cur, curh, nring = buf0, buf0h, 0
        o1, o2, o3 = C, C + ang_c, C + ang_c + epi_c
        for g in pk["groups"]:
            gin = cur
            for b in g["blocks"]:
                ops.conv(curh, b["spa0"], None, out16=spa1, act=LR, slope=0.1)
                ops.conv(spa1, b["spa2"], None, out16=cat[..., 0:o1], act=LR, slope=0.1)
                ops.conv(curh, b["ang0"], None, out16=ang1, act=LR, slope=0.1)
                ops.conv(ang1, b["ang2"], None, out16=cat[..., o1:o2], act=LR, slope=0.1, shuffle=(A, A, N.SHUF_CHANNEL_MAJOR))
                ops.conv(curh, b["epi_h"], None, out16=eh, act=LR, slope=0.1)
                ops.conv(eh, b["epi2"], None, out16=cat[..., o2:o3], act=LR, slope=0.1, shuffle=(1, A, N.SHUF_FACTOR_MAJOR))
                ops.conv(curh, b["epi_v"], None, out16=ev, act=LR, slope=0.1)
                ops.conv(ev, b["epi2"], None, out16=cat[..., o3:o3 + epi_c], act=LR, slope=0.1, shuffle=(A, 1, N.SHUF_FACTOR_MAJOR))
                ops.conv(cat, b["fuse0"], None, out16=f1, act=LR, slope=0.1)
                nxt = ring[nring]
                nring = (nring + 1) % 3
                if nxt[0] is gin:              # never overwrite the group input that is still needed
                    nxt = ring[nring]
                    nring = (nring + 1) % 3
                ops.conv(f1, b["fuse2"], nxt[0], out16=nxt[1], res=cur)
                cur, curh = nxt
            nxt = ring[nring]
            nring = (nring + 1) % 3
            if nxt[0] is gin:
                nxt = ring[nring]
                nring = (nring + 1) % 3
            ops.conv(curh, g["conv"], nxt[0], out16=nxt[1], res=gin)
            cur, curh = nxt
        last = buf("spa1_f32", H, W, C)
        ops.conv(curh, pk["cascade"], last, res=buf0)
        self._tail(ops, pk, last, out, B, H, W)


get_loss = L1Loss


def weights_init(m):
    """DistgSSR.py:169-170: a no-op."""
    pass
