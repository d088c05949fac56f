"""Plain-torch fp32 statement of every kernel contract in include/lfsr.h (test infrastructure).

`RefOps` has the same methods as lfsr_b200.kernels.CudaOps. It is used (a) on the GPU as the
per-kernel reference each CUDA kernel is compared with, and (b) on CPU, injected through
`net.set_backend(RefOps())`, to check the host-side launch plans (weight folding, buffer slicing,
fused epilogue wiring) against the oracle without a GPU. It is never reachable from the product
path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from einops import rearrange

from oracle import lf_oracle


def _act(y, act, slope):
    if act == 1:
        return F.relu(y)
    if act == 2:
        return F.leaky_relu(y, slope)
    if act == 3:
        return torch.sigmoid(y)
    if act == 4:
        return F.gelu(y)
    if act == 5:
        return F.silu(y)
    return y


def _nchw(t):
    return t.permute(0, 3, 1, 2)


def basictrans_ref(x, W, ln1, ln2, A, S, hw, nb, np_, nq, sa, ss, sb, sp, sq):
    """plain torch fp32 BasicTrans (EPIT.py:110-128) on the tokens addressed by the stride set; returns y like x"""
    import torch.nn.functional as F
    T = x.numel() // 64
    xt = x.reshape(T, 64)
    tok = xt.as_strided((nb, np_, nq, A, S, 64), (sb * 64, sp * 64, sq * 64, sa * 64, ss * 64, 1)).reshape(-1, A * S, 64)
    X = tok @ W["in"].t()
    Nn = F.layer_norm(X, (128,), ln1[0], ln1[1], ln1[2])
    q, k, v = Nn @ W["qkv"][:128].t(), Nn @ W["qkv"][128:256].t(), X @ W["qkv"][256:].t()
    hs = lambda t: t.reshape(t.shape[0], A * S, 8, 16).transpose(1, 2)
    sc = (hs(q) @ hs(k).transpose(-1, -2)) * 0.25
    spos = torch.arange(A * S, device=x.device) % S                          # token index = a*S + s
    mask = (spos[:, None] - spos[None, :]).abs() > hw
    sc = sc.masked_fill(mask, float("-inf"))
    o = (torch.softmax(sc, -1) @ hs(v)).transpose(1, 2).reshape(-1, A * S, 128)
    X2 = o @ W["o"].t() + X
    N2 = F.layer_norm(X2, (128,), ln2[0], ln2[1], ln2[2])
    X3 = torch.relu(N2 @ W["ff1"].t()) @ W["ff2"].t() + X2
    yt = X3 @ W["out"].t()
    y = torch.zeros_like(xt)
    y.as_strided((nb, np_, nq, A, S, 64), (sb * 64, sp * 64, sq * 64, sa * 64, ss * 64, 1)).copy_(
        yt.reshape(nb, np_, nq, A, S, 64))
    return y.reshape(x.shape)



class RefOps:
    name = "torch-ref"

    def __init__(self, fp16_operands: bool = False):
        #: True: the networks run the launch plans that exchange fp16 activations between tensor-core layers
        self.fp16_operands = fp16_operands

    # -- patch pipeline ---------------------------------------------------------------------------
    def divide_rows(self, scene, patches, ang, h0, w0, patch, stride, u0, u1):
        sub = lf_oracle.lfdivide(scene.detach().cpu().numpy().reshape(ang * h0, ang * w0), ang, patch, stride)
        sel = torch.from_numpy(sub[u0:u1].reshape(-1)).to(patches.device)
        patches.view(-1)[: sel.numel()].copy_(sel)

    def integrate_rows(self, patches, out, ang, pz, stride, h, w, num_u, num_v, u0, u1):
        sub = patches.detach().cpu().numpy().reshape(-1)[: (u1 - u0) * num_v * (ang * pz) ** 2]
        sub = sub.reshape(u1 - u0, num_v, ang * pz, ang * pz)
        y0, y1 = u0 * stride, min(u1 * stride, h)
        if y1 <= y0:
            return
        part = lf_oracle.lfintegrate(sub, ang, pz, stride, y1 - y0, w)         # [a1,a2,rows,w]
        o = out.view(ang, h, ang, w)
        o[:, y0:y1, :, :] = torch.from_numpy(part).permute(0, 2, 1, 3).to(out.device)

    def interp(self, x, out, n, h, w, scale, mode, block_h, block_w):
        m = "bicubic" if mode == 0 else "bilinear"
        xi = x.reshape(n, 1, h // block_h, block_h, w // block_w, block_w)
        xi = rearrange(xi, "n c a h b w -> (n a b) c h w")
        y = F.interpolate(xi, scale_factor=scale, mode=m, align_corners=False)
        y = rearrange(y, "(n a b) c h w -> n c (a h) (b w)", n=n, a=h // block_h, b=w // block_w)
        out.view(n, 1, h * scale, w * scale).copy_(y)

    # -- convolutions ------------------------------------------------------------------------------
    def conv(self, x, pc, out, act=0, slope=0.0, alpha=1.0, mul=None, mul_act=0, res=None, in_scale=None, in_perm=0, out_perm=0,
             perm_a=0, shuffle=(1, 1, 0), block=(0, 0), tail=None, out16=None):
        n = x.shape[0]
        xi = _nchw(x.float())
        if in_perm:
            xi = rearrange(xi, "b c (u h) (v w) -> b c (h u) (w v)", u=perm_a, v=perm_a)
        if in_scale is not None:
            xi = xi * in_scale.reshape(n, pc.cin, 1, 1)
        w = pc.w_f32.view(pc.kh, pc.kw, pc.cin, pc.cout).permute(3, 2, 0, 1)
        if block[0] > 0 or block[1] > 0:
            bh = block[0] if block[0] > 0 else xi.shape[2]
            bw = block[1] if block[1] > 0 else xi.shape[3]
            a, b = xi.shape[2] // bh, xi.shape[3] // bw
            xb = rearrange(xi, "n c (a h) (b w) -> (n a b) c h w", a=a, b=b)
            y = F.conv2d(xb, w, pc.bias, pc.stride, pc.pad, pc.dil)
            y = rearrange(y, "(n a b) c h w -> n c (a h) (b w)", n=n, a=a, b=b)
        else:
            y = F.conv2d(xi, w, pc.bias, pc.stride, pc.pad, pc.dil)
        y = _act(y, act, slope)
        if mul is not None:
            y = y * _act(_nchw(mul), mul_act, 0.0)
        y = y * alpha
        if out_perm:
            y = rearrange(y, "b c (h u) (w v) -> b c (u h) (v w)", u=perm_a, v=perm_a)
        ry, rx, mode = shuffle
        if ry * rx > 1:
            b_, c_, h_, w_ = y.shape
            cq = c_ // (ry * rx)
            if mode == 0:
                y = y.reshape(b_, cq, ry, rx, h_, w_).permute(0, 1, 4, 2, 5, 3)
            else:
                y = y.reshape(b_, ry, rx, cq, h_, w_).permute(0, 3, 4, 1, 5, 2)
            y = y.reshape(b_, cq, h_ * ry, w_ * rx)
        if res is not None:
            y = y + _nchw(res)
        if tail is not None:
            tw, taps, c = tail
            y = torch.einsum("bchw,ct->bthw", y, tw[:c, :taps])
        if out is not None:
            out.copy_(y.permute(0, 2, 3, 1))
        if out16 is not None:          # fp16 copy of the fp32 result (the product path rounds the same way)
            out16.copy_(y.permute(0, 2, 3, 1))

    def split_tf32(self, x, hi, lo):
        h = (x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF          # round the magnitude to 10 mantissa bits (ties away)
        hi.copy_(h.view(torch.float32).view_as(x))
        lo.copy_(x - hi)

    def macpi_unshuffle(self, x, out, ang, r, accumulate):
        n, H, W, _ = x.shape
        y = rearrange(_nchw(x), "b c (h u) (w v) -> b c (u h) (v w)", u=ang, v=ang)            # MacPI -> SAI
        y = F.pixel_shuffle(y, r).reshape(n, 1, H * r, W * r)
        o = out.view(n, 1, H * r, W * r)
        o.copy_(o + y if accumulate else y)

    def to_f16(self, x, out16):
        out16.copy_(x)

    # -- fused BasicTrans (fp16 in / out) ----------------------------------------------------------------------------
    def pack_basictrans(self, w_in, w_qkv, w_o, w_ff1, w_ff2, w_out, ln1, ln2, heads, device):
        f = lambda t: t.detach().float()
        W = {"in": f(w_in), "qkv": f(w_qkv), "o": f(w_o), "ff1": f(w_ff1), "ff2": f(w_ff2), "out": f(w_out)}
        return W, ((f(ln1[0]), f(ln1[1]), ln1[2]), (f(ln2[0]), f(ln2[1]), ln2[2]))

    def basictrans(self, x, packed, desc, y, A, S, half_window, nb, np_, nq, stride_a, stride_s, stride_b, stride_p, stride_q,
                   check_only=False):
        if not check_only:
            y.copy_(basictrans_ref(x.float().contiguous(), packed, desc[0], desc[1], A, S, half_window, nb, np_, nq, stride_a,
                                   stride_s, stride_b, stride_p, stride_q))
        return True

    def tail_supported(self, pc, cq, shuffle):
        return shuffle[0] * shuffle[1] > 1 and cq % 4 == 0

    def tap_gather(self, taps, kh, kw, bias, res, out):
        t = _nchw(taps)
        n, _, h, w = t.shape
        tp = F.pad(t, (kw // 2, kw // 2, kh // 2, kh // 2))
        y = sum(tp[:, ky * kw + kx, ky:ky + h, kx:kx + w] for ky in range(kh) for kx in range(kw)).unsqueeze(1)
        if bias is not None:
            y = y + bias.view(1, 1, 1, 1)
        if res is not None:
            y = y + _nchw(res)
        out.copy_(y.permute(0, 2, 3, 1))

    def dwconv(self, x, w, out, kh, kw, dil=(1, 1), scale=None, shift=None, act=0, slope=0.0):
        c = x.shape[3]
        wt = w.t().reshape(c, 1, kh, kw)
        y = F.conv2d(_nchw(x), wt, None, 1, ((kh // 2) * dil[0], (kw // 2) * dil[1]), dil, c)
        if scale is not None:
            y = y * scale.view(1, c, 1, 1) + shift.view(1, c, 1, 1)
        out.copy_(_act(y, act, slope).permute(0, 2, 3, 1))

    def dwconv_multi(self, x, out, branches):
        for b in branches:
            i0, o0, c = b.get("in_c0", 0), b.get("out_c0", 0), b["c"]
            self.dwconv(x[..., i0:i0 + c], b["w"], out[..., o0:o0 + c], b["kh"], b["kw"], b.get("dil", (1, 1)),
                        b.get("scale"), b.get("shift"), b.get("act", 0), b.get("slope", 0.0))

    def mel_epi_branch(self, x, w_packed, out, klen, dil, slope, tc=False):
        c = x.shape[3]
        o = 0
        def take(n, shape):
            nonlocal o
            t = w_packed[o:o + n].view(*shape)
            o += n
            return t
        dwh, dwv, dwd = take(klen * c, (klen, c)), take(klen * c, (klen, c)), take(9 * c, (9, c))
        pws = [take(c * c, (c, c)) for _ in range(3)]
        fu = take(3 * c * c, (3 * c, c))
        xi = _nchw(x)
        h = F.conv2d(xi, dwh.t().reshape(c, 1, 1, klen), None, 1, (0, klen // 2), 1, c)
        v = F.conv2d(xi, dwv.t().reshape(c, 1, klen, 1), None, 1, (klen // 2, 0), 1, c)
        d = F.conv2d(xi, dwd.t().reshape(c, 1, 3, 3), None, 1, dil, dil, c)
        outs = [F.leaky_relu(F.conv2d(t, pw.t().reshape(c, c, 1, 1)), slope) for t, pw in zip((h, v, d), pws)]
        y = F.leaky_relu(F.conv2d(torch.cat(outs, 1), fu.t().reshape(c, 3 * c, 1, 1)), slope)
        out.copy_(y.permute(0, 2, 3, 1))

    def ang_expand(self, x, w, res, out, A, act=0, slope=0.0, alpha=1.0):
        n, h, wd, c = x.shape
        y = torch.einsum("nyxk,ijkc->nyixjc", x.float(), w)                  # [n, h, A, w, A, cout]
        y = _act(y, act, slope) * alpha
        y = y.reshape(n, h * A, wd * A, w.shape[3])
        if res is not None:
            y = y + res
        out.copy_(y)

    def pooled_mlp(self, x, out, pc1, act1, pc2=None, act2=0, pool=False):
        n, h, w, c = x.shape
        t = x.float().reshape(n, h * w, c)
        if pool:
            t = t.mean(1, keepdim=True)
        t = t @ pc1.w_f32.view(pc1.cin, pc1.cout)
        if pc1.bias is not None:
            t = t + pc1.bias
        t = _act(t, act1, 0.0)
        if pc2 is not None:
            t = t @ pc2.w_f32.view(pc2.cin, pc2.cout)
            if pc2.bias is not None:
                t = t + pc2.bias
            t = _act(t, act2, 0.0)
        out.copy_(t.reshape(out.shape))

    # -- reductions / gates --------------------------------------------------------------------------
    def block_mean(self, x, out, bh, bw):
        n, h, w, c = x.shape
        out.copy_(x.reshape(n, h // bh, bh, w // bw, bw, c).mean(dim=(2, 4)))

    def sa_modulate(self, x, dw_w, bn_scale, bn_shift, amod, w0, w1, res, out, dil, out16=None, skip16=(0, 0)):
        n, h, w, c = x.shape
        xi = _nchw(x)
        s = F.conv2d(xi, dw_w.t().reshape(c, 1, 3, 3), None, 1, dil, dil, c)
        s = torch.sigmoid(s * bn_scale.view(1, c, 1, 1) + bn_shift.view(1, c, 1, 1))
        a = F.interpolate(_nchw(amod), size=(h, w), mode="nearest")
        y = xi * (w0 * s + w1 * a)
        if res is not None:
            y = y + _nchw(res)
        out.copy_(y.permute(0, 2, 3, 1))
        if out16 is not None:
            c16 = out16.shape[3]
            lo, hi = min(skip16[0], c16), min(skip16[1], c16)
            out16[..., :lo].copy_(out[..., :lo])
            out16[..., hi:].copy_(out[..., hi:c16])

    def scale_add(self, x, scale, res, out, out16=None):
        y = x * scale
        if res is not None:
            y = y + res
        (out16 if out16 is not None else out).copy_(y)

    # -- EPIT token ops ---------------------------------------------------------------------------------
    def layernorm(self, x, gamma, beta, eps, out):
        out.copy_(F.layer_norm(x, (x.shape[3],), gamma, beta, eps))

    def epi_attention(self, qk, v, out, heads, head_dim, A, S, half_window, nb, np_, nq, stride_a, stride_s, stride_b,
                      stride_p, stride_q):
        E = heads * head_dim
        dev = qk.device
        b = torch.arange(nb, device=dev).view(nb, 1, 1, 1, 1) * stride_b
        p = torch.arange(np_, device=dev).view(1, np_, 1, 1, 1) * stride_p
        q = torch.arange(nq, device=dev).view(1, 1, nq, 1, 1) * stride_q
        a = torch.arange(A, device=dev).view(1, 1, 1, A, 1) * stride_a
        s = torch.arange(S, device=dev).view(1, 1, 1, 1, S) * stride_s
        idx = (b + p + q + a + s).reshape(-1, A * S)                     # [nseq, L]
        qkf = qk.reshape(-1, 2 * E)
        vf = v.reshape(-1, E)
        Q = qkf[idx][..., :E].reshape(-1, A * S, heads, head_dim).transpose(1, 2)
        Kt = qkf[idx][..., E:].reshape(-1, A * S, heads, head_dim).transpose(1, 2)
        V = vf[idx].reshape(-1, A * S, heads, head_dim).transpose(1, 2)
        spos = torch.arange(S, device=dev).repeat(A)
        allowed = (spos[:, None] - spos[None, :]).abs() <= half_window
        scores = (Q * head_dim ** -0.5) @ Kt.transpose(-1, -2)
        scores = scores.masked_fill(~allowed, float("-inf"))
        o = torch.softmax(scores, dim=-1) @ V                              # [nseq, heads, L, d]
        o = o.transpose(1, 2).reshape(-1, A * S, E)
        out.reshape(-1, E)[idx.reshape(-1)] = o.reshape(-1, E)

    # -- metrics ---------------------------------------------------------------------------------------
    def metric_sums(self, label, out, ang, h, w, acc):
        la = label.detach().cpu().numpy().reshape(ang, h, ang, w)
        ou = out.detach().cpu().numpy().reshape(ang, h, ang, w)
        res = np.zeros((ang * ang, 2))
        for u in range(ang):
            for v in range(ang):
                a, b = la[u, :, v, :], ou[u, :, v, :]
                res[u * ang + v, 0] = ((a.astype(np.float64) - b.astype(np.float64)) ** 2).sum()
                res[u * ang + v, 1] = lf_oracle.ssim_view(a, b) * (h - 10) * (w - 10)
        acc.copy_(acc + torch.from_numpy(res.reshape(-1)).to(acc.device))
