"""LF-InterNet on liblfsr_b200 kernels - mirror of /root/reference/model/SR/LF_InterNet.py.

Two feature streams in NHWC fp32: angular `xa` [B, h, w, 64] and spatial `xs` [B, A*h, A*w, 64]
(MacPI arrangement; SAI2MacPI :155-165 is folded into the two stem convs' load addressing).
Every torch.cat of the reference (:59-60, :103, :121) is a channel window of a wider buffer:
the four block outputs are written straight into the 256(+64)-channel bottleneck inputs, and the
first chain of block b uses the (not yet written) slot of block b as the scratch half of its
128-channel concat, so nothing is ever copied.

ReconBlock (:127-141) `3x3 d=A 64->64 s^2 -> MacPI2SAI -> PixelShuffle(s) -> 1x1 64->1` is linear,
so it is composed at pack time into one 3x3 d=A conv 64->s^2 with a MacPI->SAI PixelShuffle store:
15.1 of the model's 53.7 GMAC/patch and its 1024-channel intermediate vanish, results are equal up
to fp32 summation order.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native as N
from .. import kernels as K
from .common import LFNetBase, L1Loss, slots


def _c(cin, cout, k, **kw):
    return nn.Conv2d(cin, cout, k, bias=False, **kw)


class _Chain(nn.Module):
    def __init__(self, A, ch):
        super().__init__()
        self.Spa2Ang = _c(ch, ch, A, stride=A)
        self.Ang2Spa = slots({0: _c(ch, A * A * ch, 1)})
        self.AngConvSq = _c(2 * ch, ch, 1)
        self.SpaConvSq = _c(2 * ch, ch, 3, dilation=A, padding=A)


class _InterBlock(nn.Module):
    def __init__(self, A, n_layers, ch):
        super().__init__()
        self.chained_layers = nn.ModuleList([_Chain(A, ch) for _ in range(n_layers)])


class _Cascade(nn.Module):
    def __init__(self, A, n_blocks, n_layers, ch):
        super().__init__()
        self.body = nn.ModuleList([_InterBlock(A, n_layers, ch) for _ in range(n_blocks)])


class _BottleNeck(nn.Module):
    def __init__(self, A, n_blocks, ch):
        super().__init__()
        self.AngBottle = _c(n_blocks * ch, ch, 1)
        self.Ang2Spa = slots({0: _c(ch, A * A * ch, 1)})
        self.SpaBottle = _c((n_blocks + 1) * ch, ch, 3, dilation=A, padding=A)


class _Recon(nn.Module):
    def __init__(self, A, ch, s):
        super().__init__()
        self.PreConv = _c(ch, ch * s * s, 3, dilation=A, padding=A)
        self.FinalConv = _c(ch, 1, 1)


class get_model(LFNetBase):
    def __init__(self, args):
        super().__init__(args)
        A, ch = self.angRes, 64
        self.channels = ch
        self.factor = self.scale
        self.n_groups, self.n_blocks = 4, 4
        self.AngFE = slots({0: _c(1, ch, A, stride=A)})
        self.SpaFE = slots({0: _c(1, ch, 3, dilation=A, padding=A)})
        self.CascadeInterBlock = _Cascade(A, self.n_groups, self.n_blocks, ch)
        self.BottleNeck = _BottleNeck(A, self.n_blocks, ch)
        self.ReconBlock = _Recon(A, ch, self.scale)

    def _pack(self, device, ops):
        A, s = self.angRes, self.scale
        f16 = self._fp16(ops)
        pc = lambda w, b=None, **kw: K.pack_conv(w, b, device=device, **(dict(kw, tc16=True) if (f16 and kw.get("tc")) else kw))
        dil = dict(dil=(A, A), pad=(A, A))
        pk = {"f16": f16, "ang_fe": pc(self.AngFE["0"].weight, stride=(A, A)), "spa_fe": pc(self.SpaFE["0"].weight, **dil), "chains": []}
        for blk in self.CascadeInterBlock.body:
            pk["chains"].append([dict(s2a=pc(c.Spa2Ang.weight, stride=(A, A), tc=True), a2s=pc(c.Ang2Spa["0"].weight, tc=True, tc_shuffle=(A, A, N.SHUF_CHANNEL_MAJOR)),
                                      asq=pc(c.AngConvSq.weight, tc=True), ssq=pc(c.SpaConvSq.weight, tc=True, **dil))
                                 for c in blk.chained_layers])
        bn = self.BottleNeck
        pk["bn_ang"] = pc(bn.AngBottle.weight, tc=True)
        pk["bn_a2s"] = pc(bn.Ang2Spa["0"].weight, tc=True, tc_shuffle=(A, A, N.SHUF_CHANNEL_MAJOR))
        pk["bn_spa"] = pc(bn.SpaBottle.weight, tc=True, **dil)
        wp = self.ReconBlock.PreConv.weight.detach().double()                     # [64 s^2, 64, 3, 3]
        wf = self.ReconBlock.FinalConv.weight.detach().double()[0, :, 0, 0]       # [64]
        ch = wf.numel()
        w_eff = torch.einsum("c,crikl->rikl", wf, wp.view(ch, s * s, *wp.shape[1:]))
        # the image-producing conv: TF32 tensor cores with hi / lo splits of weights and activations (3 passes ~ fp32 accuracy)
        w_eff = w_eff.float()
        w_hi = (w_eff.contiguous().view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)
        pk["recon"] = pc(w_eff, **dil)
        pk["recon_hi"] = pc(w_hi, tc=True, **dil)
        pk["recon_lo"] = pc(w_eff - w_hi, tc=True, **dil)
        return pk

    @staticmethod
    def _fp16(ops):
        from . import common
        return bool(common.USE_FP16_OPERANDS and getattr(ops, "fp16_operands", getattr(ops, "use_tc", False)) and hasattr(ops, "to_f16"))

    def _run16(self, ops, pk, x, out):
        """the launch plan of _run with fp16 activations between the tensor-core layers: the two streams (residual trunks)
        stay fp32 and carry fp16 copies inside fp16 [stream | scratch] windows; the scratch halves (Spa2Ang / Ang2Spa
        outputs), the BottleNeck intermediates and the collection buffers the BottleNeck reads exist in fp16 only."""
        A, s, C = self.angRes, self.scale, self.channels
        B, _, H, W = x.shape
        dev = x.device
        buf = lambda name, h, w, c: self._buf(name, B, h, w, c, dev)
        b16 = lambda name, h, w, c: self._buf16(name, B, h, w, c, dev)
        RL = N.ACT_RELU
        xin = x.view(B, H, W, 1)
        Y = out.view(B, H * s, W * s, 1)
        h, w = H // A, W // A
        nb = len(pk["chains"])
        CA, CS = buf("ca", h, w, (nb + 1) * C), buf("cs", H, W, (nb + 1) * C)
        CAh, CSh = b16("ca", h, w, (nb + 1) * C), b16("cs", H, W, (nb + 1) * C)
        xs0 = buf("xs0", H, W, C)
        pa = [(buf("pa0", h, w, 2 * C), b16("pa0", h, w, 2 * C)), (buf("pa1", h, w, 2 * C), b16("pa1", h, w, 2 * C))]
        ps = [(buf("ps0", H, W, 2 * C), b16("ps0", H, W, 2 * C)), (buf("ps1", H, W, 2 * C), b16("ps1", H, W, 2 * C))]
        ops.conv(xin, pk["spa_fe"], xs0, in_perm=N.PERM_MACPI_OVER_SAI, perm_a=A)
        ops.conv(xin, pk["ang_fe"], pa[0][0][..., 0:C], in_perm=N.PERM_MACPI_OVER_SAI, perm_a=A)
        ops.to_f16(pa[0][0][..., 0:C], pa[0][1][..., 0:C])
        wa, ws = pa[0], None          # current windows as (fp32, fp16) pairs: [..., :C] = stream, [..., C:] = scratch
        pp = 0
        for bi, chains in enumerate(pk["chains"]):
            for li, c in enumerate(chains):
                if bi == 0 and li == 0:
                    ws = ps[0]
                    ops.conv(xin, pk["spa_fe"], ws[0][..., 0:C], in_perm=N.PERM_MACPI_OVER_SAI, perm_a=A)
                    ops.to_f16(ws[0][..., 0:C], ws[1][..., 0:C])
                ops.conv(ws[1][..., 0:C], c["s2a"], None, out16=wa[1][..., C:2 * C], act=RL)                                 # Spa2Ang + ReLU
                ops.conv(wa[1][..., 0:C], c["a2s"], None, out16=ws[1][..., C:2 * C], shuffle=(A, A, N.SHUF_CHANNEL_MAJOR))    # Ang2Spa
                if li == len(chains) - 1:          # block output -> its slot in the collection buffers
                    na = (CA[..., bi * C:(bi + 2) * C], CAh[..., bi * C:(bi + 2) * C])
                    ns = (CS[..., bi * C:(bi + 2) * C], CSh[..., bi * C:(bi + 2) * C])
                else:
                    pp ^= 1
                    na, ns = pa[pp], ps[pp]
                    if na[0] is wa[0]:
                        pp ^= 1
                        na, ns = pa[pp], ps[pp]
                ops.conv(wa[1], c["asq"], na[0][..., 0:C], out16=na[1][..., 0:C], act=RL, res=wa[0][..., 0:C])
                ops.conv(ws[1], c["ssq"], ns[0][..., 0:C], out16=ns[1][..., 0:C], act=RL, res=ws[0][..., 0:C])
                wa, ws = na, ns
        ab = b16("ab", h, w, C)
        ops.conv(CAh[..., 0:nb * C], pk["bn_ang"], None, out16=ab, act=RL)
        ops.conv(ab, pk["bn_a2s"], None, out16=CSh[..., nb * C:(nb + 1) * C], shuffle=(A, A, N.SHUF_CHANNEL_MAJOR))
        fin = buf("fin", H, W, C)
        ops.conv(CSh, pk["bn_spa"], fin, act=RL, res=xs0)
        self._recon(ops, pk, fin, out, B, H, W)

    def _run(self, ops, pk, x, out):
        if pk.get("f16"):
            return self._run16(ops, pk, x, out)
        A, s, C = self.angRes, self.scale, self.channels
        B, _, H, W = x.shape
        dev = x.device
        buf = lambda name, h, w, c: self._buf(name, B, h, w, c, dev)
        RL = N.ACT_RELU
        xin = x.view(B, H, W, 1)
        Y = out.view(B, H * s, W * s, 1)
        h, w = H // A, W // A
        nb = len(pk["chains"])
        # collection buffers double as concat windows; +C scratch so the last block also has a window
        CA = buf("ca", h, w, (nb + 1) * C)
        CS = buf("cs", H, W, (nb + 1) * C)
        xs0 = buf("xs0", H, W, C)
        pa = [buf("pa0", h, w, 2 * C), buf("pa1", h, w, 2 * C)]
        ps = [buf("ps0", H, W, 2 * C), buf("ps1", H, W, 2 * C)]
        ops.conv(xin, pk["spa_fe"], xs0, in_perm=N.PERM_MACPI_OVER_SAI, perm_a=A)
        # first concat window of block 0 lives in the ping-pong buffers
        ops.conv(xin, pk["ang_fe"], pa[0][..., 0:C], in_perm=N.PERM_MACPI_OVER_SAI, perm_a=A)
        wa, ws = pa[0], None          # current 2C-channel windows: [..., :C] = stream, [..., C:] = scratch
        pp = 0
        for bi, chains in enumerate(pk["chains"]):
            for li, c in enumerate(chains):
                first = bi == 0 and li == 0
                if first:
                    # spatial window: copy-free trick needs the stream inside a 2C window; xs0 is
                    # also the final residual, so give the first chain its own window
                    ws = ps[0]
                    ops.conv(xin, pk["spa_fe"], ws[..., 0:C], in_perm=N.PERM_MACPI_OVER_SAI, perm_a=A)
                a_str, s_str = wa[..., 0:C], ws[..., 0:C]
                ops.conv(s_str, c["s2a"], wa[..., C:2 * C], act=RL)                                   # Spa2Ang + ReLU
                ops.conv(a_str, c["a2s"], ws[..., C:2 * C], shuffle=(A, A, N.SHUF_CHANNEL_MAJOR))     # Ang2Spa
                last = li == len(chains) - 1
                if last:          # block output -> its slot in the collection buffers
                    na, ns = CA[..., bi * C:(bi + 2) * C], CS[..., bi * C:(bi + 2) * C]
                else:
                    pp ^= 1
                    na, ns = pa[pp], ps[pp]
                    if na is wa:
                        pp ^= 1
                        na, ns = pa[pp], ps[pp]
                ops.conv(wa, c["asq"], na[..., 0:C], act=RL, res=a_str)
                ops.conv(ws, c["ssq"], ns[..., 0:C], act=RL, res=s_str)
                wa, ws = na, ns
        # BottleNeck (:117-124)
        ab = buf("ab", h, w, C)
        ops.conv(CA[..., 0:nb * C], pk["bn_ang"], ab, act=RL)
        ops.conv(ab, pk["bn_a2s"], CS[..., nb * C:(nb + 1) * C], shuffle=(A, A, N.SHUF_CHANNEL_MAJOR))
        fin = buf("fin", H, W, C)
        ops.conv(CS, pk["bn_spa"], fin, act=RL, res=xs0)
        self._recon(ops, pk, fin, out, B, H, W)

    def _recon(self, ops, pk, fin, out, B, H, W):
        """composed ReconBlock: 3x3 d=A 64->s^2 (tensor cores, MacPI arrangement), then MacPI->SAI + PixelShuffle(s) as one
        addressing pass (the tensor-core epilogue does not address permuted outputs)"""
        A, s = self.angRes, self.scale
        if hasattr(ops, "macpi_unshuffle") and hasattr(ops, "split_tf32") and s in (2, 4) and getattr(ops, "use_tc", True):
            C = self.channels
            rec = self._buf("recon", B, H, W, s * s, fin.device)
            hi, lo = self._buf("fin_hi", B, H, W, C, fin.device), self._buf("fin_lo", B, H, W, C, fin.device)
            ops.split_tf32(fin, hi, lo)
            ops.conv(hi, pk["recon_hi"], rec)
            ops.conv(lo, pk["recon_hi"], rec, res=rec)
            ops.conv(hi, pk["recon_lo"], rec, res=rec)
            ops.macpi_unshuffle(rec, out, A, s, False)
        else:
            Y = out.view(B, H * s, W * s, 1)
            ops.conv(fin, pk["recon"], Y, out_perm=N.PERM_MACPI_OVER_SAI, perm_a=A, shuffle=(s, s, N.SHUF_CHANNEL_MAJOR))


get_loss = L1Loss


def weights_init(m):
    """LF_InterNet.py:168-170: a no-op."""
    pass
