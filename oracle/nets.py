"""ORACLE (test infrastructure, not product code): plain-PyTorch fp32 CPU restatement of the
reference networks' eval forward, written functionally over a reference-format state_dict.

Each function cites the reference lines it follows (paths relative to /root/reference/model/SR).
Pinned against the unmodified reference modules imported in the build container on seeded inputs
and weights (oracle/make_golden.py; tests/test_oracle_pinned.py). Used as: the GPU parity checker,
bench.py's cpu_baseline, and the `--impl reference` arm on the GPU box (the Python reference
itself cannot travel there).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from einops import rearrange


def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, eps)


def _lrelu(x, s):
    return F.leaky_relu(x, s)


# ---------------------------------------------------------------------------------------------
# MyEfficientLFNet (v2.0)  — MyEfficientLFNet.py
# ---------------------------------------------------------------------------------------------
def _repconv(x, sd, p, dil):
    """RepConvBlock.forward, un-fused branches (MyEfficientLFNet.py:374-385)."""
    out = _bn(F.conv2d(x, sd[p + ".conv_3x3.weight"], None, 1, dil, dil), sd, p + ".bn_3x3")
    out = out + _bn(F.conv2d(x, sd[p + ".conv_1x1.weight"]), sd, p + ".bn_1x1")
    if (p + ".bn_identity.weight") in sd:
        out = out + _bn(x, sd, p + ".bn_identity")
    return out


def _mel_stage(x, sd, p, A):
    """ProgressiveDisentanglingStage.forward (MyEfficientLFNet.py:183-208)."""
    C = x.shape[1]
    sp = [C // 3, C // 3, C - 2 * (C // 3)]
    xs, xa, xe = torch.split(x, sp, dim=1)
    # spatial branch (:134-141)
    fs = _repconv(xs, sd, p + ".spatial_branch.0", A)
    fs = F.conv2d(_lrelu(fs, 0.1), sd[p + ".spatial_branch.2.weight"], None, 1, A, A)
    # LightweightAngularAttention (:254-275)
    q = p + ".angular_branch"
    ang = F.conv2d(xa, sd[q + ".to_angular.weight"], None, A)
    att = F.relu(F.conv2d(ang, sd[q + ".attention.0.weight"]))
    att = F.relu(F.conv2d(att, sd[q + ".attention.2.weight"], None, 1, 1, 1, att.shape[1]))
    att = torch.sigmoid(F.conv2d(att, sd[q + ".attention.4.weight"]))
    ang = ang * att
    ang = _lrelu(F.conv2d(ang, sd[q + ".cross_view.0.weight"], None, 1, 1), 0.1)
    ex = _lrelu(F.pixel_shuffle(F.conv2d(ang, sd[q + ".expand.0.weight"]), A), 0.1)
    fa = xa + sd[q + ".scale"] * ex
    # MultiScaleEPIBlock (:323-327)
    q = p + ".epi_branch"
    ce = xe.shape[1]
    h = _lrelu(F.conv2d(F.conv2d(xe, sd[q + ".epi_h.0.weight"], None, 1, (0, A), 1, ce), sd[q + ".epi_h.1.weight"]), 0.1)
    v = _lrelu(F.conv2d(F.conv2d(xe, sd[q + ".epi_v.0.weight"], None, 1, (A, 0), 1, ce), sd[q + ".epi_v.1.weight"]), 0.1)
    d = _lrelu(F.conv2d(F.conv2d(xe, sd[q + ".epi_diag.0.weight"], None, 1, A, A, ce), sd[q + ".epi_diag.1.weight"]), 0.1)
    fe = _lrelu(F.conv2d(torch.cat([h, v, d], 1), sd[q + ".fuse.0.weight"]), 0.1)
    # gates (:159-173, :193-196)
    def gate(f, name):
        g = F.adaptive_avg_pool2d(f, 1)
        return f * torch.sigmoid(F.conv2d(g, sd[p + f".{name}.1.weight"], sd[p + f".{name}.1.bias"]))
    fs, fa, fe = gate(fs, "gate_spatial"), gate(fa, "gate_angular"), gate(fe, "gate_epi")
    fused = torch.cat([fs, fa, fe], 1)
    fused = _lrelu(F.conv2d(fused, sd[p + ".fusion.0.weight"]), 0.1)
    fused = F.conv2d(fused, sd[p + ".fusion.2.weight"], None, 1, A, A)
    # SAModulator (:495-515)
    q = p + ".sa_modulator"
    s_mod = torch.sigmoid(_bn(F.conv2d(fused, sd[q + ".spatial_mod.0.weight"], None, 1, A, A, C), sd, q + ".spatial_mod.1"))
    a_mod = F.adaptive_avg_pool2d(fused, A)
    a_mod = torch.sigmoid(F.conv2d(F.relu(F.conv2d(a_mod, sd[q + ".angular_conv.0.weight"])), sd[q + ".angular_conv.2.weight"]))
    a_mod = F.interpolate(a_mod, size=fused.shape[2:], mode="nearest")
    wts = F.softmax(sd[q + ".combine"], dim=0)
    return fused * (wts[0] * s_mod + wts[1] * a_mod) + x


def my_efficient_lfnet(x, sd, ang=5, scale=4):
    """get_model.forward (MyEfficientLFNet.py:76-109)."""
    x_up = F.interpolate(x, scale_factor=scale, mode="bicubic", align_corners=False)
    feat = _repconv(x, sd, "shallow_feat", ang)
    shallow = feat
    n_stages = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("stages."))
    for i in range(n_stages):
        feat = _mel_stage(feat, sd, f"stages.{i}", ang)
    g = _lrelu(F.conv2d(feat, sd["global_fusion.0.weight"]), 0.1)
    feat = _repconv(g, sd, "global_fusion.2", ang) + shallow
    # PixelShuffleUpsampler (:548-582)
    if scale == 4:
        feat = _lrelu(F.pixel_shuffle(F.conv2d(feat, sd["upsampler.up.0.weight"], None, 1, 1), 2), 0.1)
        feat = _lrelu(F.pixel_shuffle(F.conv2d(feat, sd["upsampler.up.3.weight"], None, 1, 1), 2), 0.1)
    else:
        feat = _lrelu(F.pixel_shuffle(F.conv2d(feat, sd["upsampler.up.0.weight"], None, 1, 1), scale), 0.1)
    return F.conv2d(feat, sd["output_conv.weight"], sd["output_conv.bias"], 1, 1) + x_up


# ---------------------------------------------------------------------------------------------
# MyEfficientLFNetV4_5 with its in-repo FastConvSSM branch (mamba_ssm absent) — MyEfficientLFNetV4_5.py
# ---------------------------------------------------------------------------------------------
def _v45_block(x, sd, p):
    """MambaLFBlock.forward (MyEfficientLFNetV4_5.py:143-148)."""
    C = x.shape[1]
    c = C // 4
    # MultiScaleSpatial (:262-282)
    q = p + ".ms_spatial"
    y = torch.cat([F.conv2d(x[:, :c], sd[q + ".conv1.weight"]),
                   F.conv2d(x[:, c:2 * c], sd[q + ".conv3.weight"], None, 1, 1, 1, c),
                   F.conv2d(x[:, 2 * c:3 * c], sd[q + ".conv5.weight"], None, 1, 2, 1, c),
                   F.conv2d(x[:, 3 * c:], sd[q + ".conv7.weight"], None, 1, 3, 1, c)], 1)
    f_local = _lrelu(_bn(F.conv2d(y, sd[q + ".pw.weight"]), sd, q + ".bn"), 0.1) + x
    # FastConvSSM (:208-244)
    q = p + ".ssm"
    yn = _bn(x, sd, q + ".norm")
    g = F.gelu(F.conv2d(yn, sd[q + ".gate_conv.weight"]))
    gate, yv = g.chunk(2, dim=1)
    fs = [F.conv2d(yv, sd[q + f".conv{d}.weight"], None, 1, d, d, C) for d in (1, 2, 4, 8)]
    yv = F.conv2d(torch.cat(fs, 1), sd[q + ".fuse.weight"]) * F.silu(gate)
    f_global = x + sd[q + ".scale"] * F.conv2d(yv, sd[q + ".proj.weight"])
    fused = F.conv2d(torch.cat([f_local, f_global], 1), sd[p + ".fuse.weight"])
    # ChannelAttention (:285-299)
    a = F.adaptive_avg_pool2d(fused, 1)
    a = torch.sigmoid(F.conv2d(F.relu(F.conv2d(a, sd[p + ".ca.fc1.weight"], sd[p + ".ca.fc1.bias"])),
                               sd[p + ".ca.fc2.weight"], sd[p + ".ca.fc2.bias"]))
    return fused * a + x


def my_efficient_lfnet_v4_5(x, sd, ang=5, scale=4):
    """get_model.forward (MyEfficientLFNetV4_5.py:65-109), use_macpi=False (its default, :39)."""
    x_up = F.interpolate(x, scale_factor=scale, mode="bicubic", align_corners=False)
    feat = _lrelu(F.conv2d(x, sd["shallow.0.weight"], None, 1, 1), 0.1)
    C = feat.shape[1]
    t = _lrelu(_bn(F.conv2d(feat, sd["shallow.2.dw.weight"], None, 1, 1, 1, C), sd, "shallow.2.bn"), 0.1)
    feat = feat + F.conv2d(t, sd["shallow.2.pw.weight"])                         # LocalPixelEnhancement (:285-294)
    shallow = feat
    n_blocks = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    outs = []
    for i in range(n_blocks):
        feat = _v45_block(feat, sd, f"blocks.{i}")
        outs.append(feat)
    early = F.conv2d(torch.cat(outs[:4], 1), sd["fuse_early.weight"])
    late = F.conv2d(torch.cat(outs[4:], 1), sd["fuse_late.weight"])
    feat = F.conv2d(torch.cat([early, late], 1), sd["fuse_final.weight"]) + shallow
    feat = _lrelu(F.conv2d(feat, sd["refine.weight"], None, 1, 1), 0.1)
    if scale == 4:
        feat = _lrelu(F.pixel_shuffle(F.conv2d(feat, sd["upsampler.up.0.weight"], None, 1, 1), 2), 0.1)
        feat = _lrelu(F.pixel_shuffle(F.conv2d(feat, sd["upsampler.up.3.weight"], None, 1, 1), 2), 0.1)
    else:
        feat = _lrelu(F.pixel_shuffle(F.conv2d(feat, sd["upsampler.up.0.weight"], None, 1, 1), scale), 0.1)
    return F.conv2d(feat, sd["output.weight"], sd["output.bias"], 1, 1) + x_up


# ---------------------------------------------------------------------------------------------
# EPIT — EPIT.py
# ---------------------------------------------------------------------------------------------
def _epit_mask(v, w, k_h, k_w):
    """BasicTrans.gen_mask (EPIT.py:93-108), vectorised: 0 where allowed, -inf elsewhere."""
    ii, jj = torch.meshgrid(torch.arange(v), torch.arange(w), indexing="ij")
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    hl, wl = k_h // 2, k_w // 2
    hr, wr = k_h - hl, k_w - wl
    ok = ((ii[None, :] >= (ii[:, None] - hl)) & (ii[None, :] < (ii[:, None] + hr)) &
          (jj[None, :] >= (jj[:, None] - wl)) & (jj[None, :] < (jj[:, None] + wr)))
    m = torch.zeros(v * w, v * w)
    return m.masked_fill(~ok, float("-inf"))


def _basic_trans(buf, sd, p, mask_field, heads=8):
    """BasicTrans.forward (EPIT.py:110-128); MultiheadAttention without biases, q=k=LN(x), v=x."""
    _, _, n, v, w = buf.shape
    mask = _epit_mask(v, w, mask_field[0], mask_field[1]).to(buf.device)
    tok = rearrange(buf, "b c n v w -> (v w) (b n) c")
    tok = F.linear(tok, sd[p + ".linear_in.weight"])
    E = tok.shape[-1]
    tn = F.layer_norm(tok, (E,), sd[p + ".norm.weight"], sd[p + ".norm.bias"])
    att, _ = F.multi_head_attention_forward(
        tn, tn, tok, E, heads, sd[p + ".attention.in_proj_weight"], None, None, None, False, 0.0,
        sd[p + ".attention.out_proj.weight"], None, training=False, need_weights=False, attn_mask=mask)
    tok = att + tok
    ff = F.layer_norm(tok, (E,), sd[p + ".feed_forward.0.weight"], sd[p + ".feed_forward.0.bias"])
    ff = F.linear(F.relu(F.linear(ff, sd[p + ".feed_forward.1.weight"])), sd[p + ".feed_forward.4.weight"])
    tok = ff + tok
    tok = F.linear(tok, sd[p + ".linear_out.weight"])
    return rearrange(tok, "(v w) (b n) c -> b c n v w", v=v, w=w, n=n)


def _epit_conv3(x, sd, p):
    y = _lrelu(F.conv3d(x, sd[p + ".0.weight"], None, 1, (0, 1, 1)), 0.2)
    y = _lrelu(F.conv3d(y, sd[p + ".2.weight"], None, 1, (0, 1, 1)), 0.2)
    return F.conv3d(y, sd[p + ".4.weight"], None, 1, (0, 1, 1))


def _alt_filter(buf, sd, p, A):
    """AltFilter.forward (EPIT.py:144-161)."""
    short = buf
    h, w = buf.shape[-2:]
    mf = [A * 2, 11]
    b = rearrange(buf, "b c (u v) h w -> b c (v w) u h", u=A, v=A)
    b = _basic_trans(b, sd, p + ".epi_trans", mf)
    b = rearrange(b, "b c (v w) u h -> b c (u v) h w", u=A, v=A, h=h, w=w)
    b = _epit_conv3(b, sd, p + ".conv") + short
    b = rearrange(b, "b c (u v) h w -> b c (u h) v w", u=A, v=A)
    b = _basic_trans(b, sd, p + ".epi_trans", mf)
    b = rearrange(b, "b c (u h) v w -> b c (u v) h w", u=A, v=A, h=h, w=w)
    return _epit_conv3(b, sd, p + ".conv") + short


def epit(lr, sd, ang=5, scale=4):
    """get_model.forward (EPIT.py:51-71)."""
    lr6 = rearrange(lr, "b c (u h) (v w) -> b c u v h w", u=ang, v=ang)
    b, c, u, v, h, w = lr6.shape
    up = F.interpolate(rearrange(lr6, "b c u v h w -> (b u v) c h w"), scale_factor=scale, mode="bicubic",
                       align_corners=False)
    sr_y = rearrange(up, "(b u v) c h w -> b c (u h) (v w)", u=u, v=v)
    x = rearrange(lr6, "b c u v h w -> b c (u v) h w")
    buf = F.conv3d(x, sd["conv_init0.0.weight"], None, 1, (0, 1, 1))
    y = buf
    for i in (0, 2, 4):
        y = _lrelu(F.conv3d(y, sd[f"conv_init.{i}.weight"], None, 1, (0, 1, 1)), 0.2)
    buf = y + buf
    y = buf
    n_alt = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("altblock."))
    for i in range(n_alt):
        y = _alt_filter(y, sd, f"altblock.{i}", ang)
    buf = y + buf
    buf = rearrange(buf, "b c (u v) h w -> b c (u h) (v w)", u=u, v=v)
    y = F.conv2d(buf, sd["upsampling.0.weight"])
    y = _lrelu(F.pixel_shuffle(y, scale), 0.2)
    return F.conv2d(y, sd["upsampling.3.weight"], None, 1, 1) + sr_y


# ---------------------------------------------------------------------------------------------
# DistgSSR — DistgSSR.py ; LF_InterNet — LF_InterNet.py
# ---------------------------------------------------------------------------------------------
def sai2macpi(x, A):
    """SAI2MacPI (DistgSSR.py:145-155): mac[i*A+u, j*A+v] = sai[u*h+i, v*w+j]."""
    return rearrange(x, "b c (u h) (v w) -> b c (h u) (w v)", u=A, v=A)


def macpi2sai(x, A):
    """MacPI2SAI (DistgSSR.py:134-142)."""
    return rearrange(x, "b c (h u) (w v) -> b c (u h) (v w)", u=A, v=A)


def _ps1d(x, f):
    """PixelShuffle1D (DistgSSR.py:114-131): [b, f*c, h, w] -> [b, c, h, w*f], factor-major channels."""
    b, fc, h, w = x.shape
    c = fc // f
    return x.contiguous().view(b, f, c, h, w).permute(0, 2, 3, 4, 1).contiguous().view(b, c, h, w * f)


def _disentg_block(x, sd, p, A):
    """DisentgBlock.forward (DistgSSR.py:103-111)."""
    spa = _lrelu(F.conv2d(x, sd[p + ".SpaConv.0.weight"], None, 1, A, A), 0.1)
    spa = _lrelu(F.conv2d(spa, sd[p + ".SpaConv.2.weight"], None, 1, A, A), 0.1)
    ang = _lrelu(F.conv2d(x, sd[p + ".AngConv.0.weight"], None, A), 0.1)
    ang = F.pixel_shuffle(_lrelu(F.conv2d(ang, sd[p + ".AngConv.2.weight"]), 0.1), A)

    def epi(z):
        e = _lrelu(F.conv2d(z, sd[p + ".EPIConv.0.weight"], None, (1, A), (0, A * (A - 1) // 2)), 0.1)
        return _ps1d(_lrelu(F.conv2d(e, sd[p + ".EPIConv.2.weight"]), 0.1), A)
    eh = epi(x)
    ev = epi(x.permute(0, 1, 3, 2).contiguous()).permute(0, 1, 3, 2)
    buf = torch.cat((spa, ang, eh, ev), 1)
    buf = _lrelu(F.conv2d(buf, sd[p + ".fuse.0.weight"]), 0.1)
    return F.conv2d(buf, sd[p + ".fuse.2.weight"], None, 1, A, A) + x


def distgssr(x, sd, ang=5, scale=4):
    """get_model.forward (DistgSSR.py:29-36) with CascadeDisentgGroup / DisentgGroup (:39-70)."""
    x_up = F.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=False)
    buf0 = F.conv2d(sai2macpi(x, ang), sd["init_conv.weight"], None, 1, ang, ang)
    n_group = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("disentg.Group."))
    n_block = 1 + max(int(k.split(".")[4]) for k in sd if k.startswith("disentg.Group.0.Block."))
    buf = buf0
    for g in range(n_group):
        gin = buf
        for bi in range(n_block):
            buf = _disentg_block(buf, sd, f"disentg.Group.{g}.Block.{bi}", ang)
        buf = F.conv2d(buf, sd[f"disentg.Group.{g}.conv.weight"], None, 1, ang, ang) + gin
    buf = F.conv2d(buf, sd["disentg.conv.weight"], None, 1, ang, ang) + buf0
    sai = macpi2sai(buf, ang)
    y = F.pixel_shuffle(F.conv2d(sai, sd["upsample.0.weight"], sd["upsample.0.bias"]), scale)
    return F.conv2d(y, sd["upsample.2.weight"]) + x_up


def lf_internet(x, sd, ang=5, scale=4):
    """get_model.forward (LF_InterNet.py:34-41), make_chains (:44-67), BottleNeck (:107-124),
    ReconBlock (:127-141)."""
    xm = sai2macpi(x, ang)
    xa = F.conv2d(xm, sd["AngFE.0.weight"], None, ang)
    xs = F.conv2d(xm, sd["SpaFE.0.weight"], None, 1, ang, ang)
    n_blocks = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("CascadeInterBlock.body."))
    n_layers = 1 + max(int(k.split(".")[4]) for k in sd if k.startswith("CascadeInterBlock.body.0.chained_layers."))
    ba, bs = xa, xs
    outs_a, outs_s = [], []
    for i in range(n_blocks):
        for j in range(n_layers):
            p = f"CascadeInterBlock.body.{i}.chained_layers.{j}"
            a2 = F.relu(F.conv2d(bs, sd[p + ".Spa2Ang.weight"], None, ang))
            s2 = F.pixel_shuffle(F.conv2d(ba, sd[p + ".Ang2Spa.0.weight"]), ang)
            na = F.relu(F.conv2d(torch.cat((ba, a2), 1), sd[p + ".AngConvSq.weight"])) + ba
            ns = F.relu(F.conv2d(torch.cat((bs, s2), 1), sd[p + ".SpaConvSq.weight"], None, 1, ang, ang)) + bs
            ba, bs = na, ns
        outs_a.append(ba)
        outs_s.append(bs)
    ca, cs = torch.cat(outs_a, 1), torch.cat(outs_s, 1)
    a = F.relu(F.conv2d(ca, sd["BottleNeck.AngBottle.weight"]))
    s = torch.cat((cs, F.pixel_shuffle(F.conv2d(a, sd["BottleNeck.Ang2Spa.0.weight"]), ang)), 1)
    out = F.relu(F.conv2d(s, sd["BottleNeck.SpaBottle.weight"], None, 1, ang, ang)) + xs
    buf = F.conv2d(out, sd["ReconBlock.PreConv.weight"], None, 1, ang, ang)
    hr = F.pixel_shuffle(macpi2sai(buf, ang), scale)
    return F.conv2d(hr, sd["ReconBlock.FinalConv.weight"])


FORWARD = {
    "MyEfficientLFNet": my_efficient_lfnet,
    "MyEfficientLFNetV4_5": my_efficient_lfnet_v4_5,
    "EPIT": epit,
    "DistgSSR": distgssr,
    "LF_InterNet": lf_internet,
}


def forward(model_name: str, x: torch.Tensor, sd: dict, ang: int = 5, scale: int = 4) -> torch.Tensor:
    with torch.no_grad():
        return FORWARD[model_name](x, sd, ang, scale)
