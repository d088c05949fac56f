"""On-device scene driver: the hot loop of train.test() (train.py:286-322) without its per-patch
host round trips. LFdivide -> batched forward -> LFintegrate (-> PSNR/SSIM) all stay in HBM; with
torch.distributed initialised, one scene is split by patch-grid rows across the ranks and the
stitched stripes are all-gathered over NCCL - there is no other exchange step.
"""
from __future__ import annotations

import torch

from . import kernels as K
from . import lfutils as U


def shard_rows(num_u: int, world: int, rank: int):
    """contiguous, balanced split of the patch-grid rows: rank r owns [lo, hi)."""
    base, extra = divmod(num_u, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def super_resolve_rows(net, lr_sai: torch.Tensor, ang: int, scale: int, patch: int = 32, stride: int = 16,
                       minibatch: int = 64, rows=None, out_mosaic: torch.Tensor = None, ops=None) -> torch.Tensor:
    """SR of the patch-grid rows `rows` of one scene; returns the full-size SAI mosaic
    [(a1 h s), (a2 w s)] with only the owned stripes written."""
    ops = ops or getattr(net, "_ops", None) or K.default_ops()
    H, W = lr_sai.shape
    h0, w0 = H // ang, W // ang
    _, num_u, num_v = U.divide_geometry(h0, w0, patch, stride)
    u0, u1 = (0, num_u) if rows is None else rows
    dev = lr_sai.device
    if out_mosaic is None:
        out_mosaic = torch.zeros((ang * h0 * scale, ang * w0 * scale), dtype=torch.float32, device=dev)
    if u1 <= u0:
        return out_mosaic
    sub = U.LFdivide(lr_sai, ang, patch, stride, rows=(u0, u1), ops=ops)          # [rows, numV, A*P, A*P]
    n = (u1 - u0) * num_v
    sub = sub.view(n, 1, ang * patch, ang * patch)
    pz = patch * scale
    sr = torch.empty((n, 1, ang * pz, ang * pz), dtype=torch.float32, device=dev)
    for i in range(0, n, minibatch):
        sr[i:i + minibatch] = net(sub[i:i + minibatch], [ang, ang])
    ops.integrate_rows(sr, out_mosaic, ang, pz, stride * scale, h0 * scale, w0 * scale, num_u, num_v, u0, u1)
    return out_mosaic


def super_resolve_scene(net, lr_sai: torch.Tensor, ang: int, scale: int, patch: int = 32, stride: int = 16,
                        minibatch: int = 64, group=None, ops=None) -> torch.Tensor:
    """Full scene; when a process group is given (or torch.distributed is initialised with
    world_size > 1) each rank computes its row band and the bands are all-gathered."""
    import torch.distributed as dist
    world, rank = 1, 0
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    H, W = lr_sai.shape
    h0, w0 = H // ang, W // ang
    _, num_u, _ = U.divide_geometry(h0, w0, patch, stride)
    if world == 1:
        return super_resolve_rows(net, lr_sai, ang, scale, patch, stride, minibatch, None, None, ops)
    lo, hi = shard_rows(num_u, world, rank)
    mosaic = super_resolve_rows(net, lr_sai, ang, scale, patch, stride, minibatch, (lo, hi), None, ops)
    return gather_stripes(mosaic, ang, h0 * scale, w0 * scale, stride * scale, num_u, world, group)


def gather_stripes(mosaic: torch.Tensor, ang: int, h: int, w: int, ss: int, num_u: int, world: int, group=None):
    """all-gather of the per-rank stripes: rank r owns view rows [lo_r*ss, min(hi_r*ss, h)) of every
    view. Stripes are packed [A, rows, A*w] so one all_gather per scene moves each byte once."""
    import torch.distributed as dist
    view = mosaic.view(ang, h, ang * w)
    bounds = [shard_rows(num_u, world, r) for r in range(world)]
    spans = [(min(lo * ss, h), min(hi * ss, h)) for lo, hi in bounds]
    max_rows = max(b - a for a, b in spans)
    rank = dist.get_rank(group)
    send = torch.zeros((ang, max_rows, ang * w), dtype=mosaic.dtype, device=mosaic.device)
    a, b = spans[rank]
    send[:, : b - a] = view[:, a:b]
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    for r, (a, b) in enumerate(spans):
        if r != rank and b > a:
            view[:, a:b] = recv[r][:, : b - a]
    return mosaic


def test_scene(net, lr_sai, hr_sai, ang: int, scale: int, patch: int = 32, stride: int = 16, minibatch: int = 64,
               ops=None):
    """(psnr, ssim, sr_mosaic) for one scene - the body of train.test()'s loop on the device."""
    class _A:
        angRes_in = ang
        task = "SR"
    sr = super_resolve_scene(net, lr_sai, ang, scale, patch, stride, minibatch, ops=ops)
    psnr, ssim = U.cal_metrics(_A, hr_sai.reshape(1, 1, *hr_sai.shape[-2:]), sr.reshape(1, 1, *sr.shape), ops=ops)
    return psnr, ssim, sr
