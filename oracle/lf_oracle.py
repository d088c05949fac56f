"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's patch pipeline.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. Every function cites the reference lines it restates (paths relative to
/root/reference). Pinned against the *unmodified reference imported in the build container*
(oracle/make_golden.py -> tests/golden/*.npz, tests/test_oracle_pinned.py): the reference ships no
golden vectors or tests of its own for this path.

skimage is not installed anywhere in this environment; `psnr_view` / `ssim_view` restate
skimage.metrics.peak_signal_noise_ratio / structural_similarity (scikit-image >= 0.19,
requirements.txt:25, unpinned) on scipy.ndimage - parity for that third-party arithmetic is
therefore "unpinned" beyond the published algorithm.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import gaussian_filter


# ---------------------------------------------------------------------------------------------
# LFdivide / LFintegrate : utils/utils.py:137-178
# ---------------------------------------------------------------------------------------------
def _mirror(i: np.ndarray, n: int) -> np.ndarray:
    """ImageExtend (utils/utils.py:137-149): flip+cat = symmetric mirror that repeats the edge sample."""
    i = np.asarray(i)
    return np.where(i < 0, -i - 1, np.where(i >= n, 2 * n - 1 - i, i))


def divide_geometry(h0: int, w0: int, patch: int, stride: int):
    bdr = (patch - stride) // 2
    return bdr, (h0 + bdr * 2 - 1) // stride, (w0 + bdr * 2 - 1) // stride      # utils.py:156-158


def divide_supported(h0: int, w0: int, patch: int, stride: int) -> bool:
    """does the reference's LFdivide run on this geometry? ImageExtend (utils.py:137-149) can only pad out of one mirrored
    copy (<= n rows per side) and F.unfold (utils.py:160) must return exactly the numU x numV windows that the rearrange
    of utils.py:161-164 expects, else einops raises (checked against the reference in oracle/make_golden.py)."""
    bdr = (patch - stride) // 2
    for n in (h0, w0):
        ext = n + bdr + min(bdr + stride - 1, n)
        windows = (ext - patch) // stride + 1 if ext >= patch else 0
        if n <= 0 or bdr > n or windows < 1 or windows != (n + 2 * bdr - 1) // stride:
            return False
    return True


def lfdivide(scene: np.ndarray, ang: int, patch: int, stride: int) -> np.ndarray:
    """scene [(a1 h0), (a2 w0)] -> [numU, numV, (a1 P), (a2 P)]  (utils/utils.py:152-166)."""
    H, W = scene.shape
    h0, w0 = H // ang, W // ang
    if not divide_supported(h0, w0, patch, stride):
        raise ValueError(f"lfdivide: the reference cannot tile {h0}x{w0} views with patch {patch} / stride {stride}")
    bdr, num_u, num_v = divide_geometry(h0, w0, patch, stride)
    views = scene.reshape(ang, h0, ang, w0)
    ys = _mirror(np.arange(num_u)[:, None] * stride + np.arange(patch)[None, :] - bdr, h0)  # [numU, P]
    xs = _mirror(np.arange(num_v)[:, None] * stride + np.arange(patch)[None, :] - bdr, w0)  # [numV, P]
    # out[n1, n2, a1, y, a2, x] = views[a1, ys[n1, y], a2, xs[n2, x]]
    out = views[:, ys][:, :, :, :, xs]            # [a1, n1, y, a2, n2, x]
    out = out.transpose(1, 4, 0, 2, 3, 5)         # [n1, n2, a1, y, a2, x]
    return np.ascontiguousarray(out).reshape(num_u, num_v, ang * patch, ang * patch)


def lfintegrate(sub: np.ndarray, ang: int, pz: int, stride: int, h: int, w: int) -> np.ndarray:
    """sub [n1, n2, (a1 pz), (a2 pz)] -> [a1, a2, h, w]  (utils/utils.py:169-178)."""
    n1, n2 = sub.shape[:2]
    s = sub.reshape(n1, n2, ang, pz, ang, pz)
    bdr = (pz - stride) // 2
    s = s[:, :, :, bdr:bdr + stride, :, bdr:bdr + stride]     # n1 n2 a1 y a2 x
    s = s.transpose(2, 4, 0, 3, 1, 5).reshape(ang, ang, n1 * stride, n2 * stride)
    return np.ascontiguousarray(s[:, :, :h, :w])


def to_sai(lf4d: np.ndarray) -> np.ndarray:
    """'a1 a2 h w -> (a1 h) (a2 w)' (train.py:319)."""
    a1, a2, h, w = lf4d.shape
    return np.ascontiguousarray(lf4d.transpose(0, 2, 1, 3)).reshape(a1 * h, a2 * w)


# ---------------------------------------------------------------------------------------------
# cal_metrics : utils/utils.py:91-134 (SR task branch) on skimage.metrics restated
# ---------------------------------------------------------------------------------------------
def psnr_view(a: np.ndarray, b: np.ndarray, data_range: float = 1.0) -> float:
    """skimage.metrics.peak_signal_noise_ratio: float64 MSE, 10*log10(R^2/mse)."""
    err = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2, dtype=np.float64)
    return float(10.0 * np.log10((data_range ** 2) / err)) if err > 0 else float("inf")


def ssim_view(a: np.ndarray, b: np.ndarray, data_range: float = 1.0) -> float:
    """skimage.metrics.structural_similarity(gaussian_weights=True, data_range=1.0): sigma 1.5,
    truncate 3.5 -> 11x11 window, use_sample_covariance=True (default; utils.py:116-118), float32
    maps for float32 input, crop (win-1)//2 = 5, float64 mean."""
    a = a.astype(np.float32)
    b = b.astype(np.float32)
    K1, K2, sigma, truncate = 0.01, 0.03, 1.5, 3.5
    win = 2 * int(truncate * sigma + 0.5) + 1
    npx = win ** 2
    cov_norm = npx / (npx - 1)
    f = lambda z: gaussian_filter(z, sigma=sigma, truncate=truncate, mode="reflect")
    ux, uy = f(a), f(b)
    uxx, uyy, uxy = f(a * a), f(b * b), f(a * b)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean(dtype=np.float64))


def cal_metrics(label_sai: np.ndarray, out_sai: np.ndarray, ang: int):
    """label/out: SAI mosaics [(a1 h), (a2 w)] -> (PSNR_mean, SSIM_mean) averaged over views with
    value > 0 (utils.py:127-132). Returns also the per-view arrays."""
    H, W = label_sai.shape
    h, w = H // ang, W // ang
    la = label_sai.reshape(ang, h, ang, w)
    ou = out_sai.reshape(ang, h, ang, w)
    psnr = np.zeros((ang, ang), np.float32)
    ssim = np.zeros((ang, ang), np.float32)
    for u in range(ang):
        for v in range(ang):
            psnr[u, v] = psnr_view(la[u, :, v, :], ou[u, :, v, :])
            ssim[u, v] = ssim_view(la[u, :, v, :], ou[u, :, v, :])
    vp = np.sum(psnr > 0)
    vs = np.sum(ssim > 0)
    pm = psnr.sum() / vp if vp > 0 else 0.0
    sm = ssim.sum() / vs if vs > 0 else 0.0
    return float(pm), float(sm), psnr, ssim


# ---------------------------------------------------------------------------------------------
# MATLAB-style imresize (utils/imresize.py): separable, antialiased when shrinking, symmetric border
# ---------------------------------------------------------------------------------------------
def _cubic(x):
    """Keys kernel with A = -0.5 (utils/imresize.py:24-30)."""
    ax = np.abs(np.asarray(x, dtype=np.float64))
    ax2, ax3 = ax * ax, ax * ax * ax
    return (1.5 * ax3 - 2.5 * ax2 + 1) * (ax <= 1) + (-0.5 * ax3 + 2.5 * ax2 - 4 * ax + 2) * ((1 < ax) & (ax <= 2))


def _triangle(x):
    """utils/imresize.py:17-22."""
    x = np.asarray(x, dtype=np.float64)
    return (x + 1) * ((x >= -1) & (x < 0)) + (1 - x) * ((x <= 1) & (x >= 0))


def imresize_contributions(in_length: int, out_length: int, scale: float, method: str = "bicubic"):
    """(weights [out, P], indices [out, P] int32) of one dimension (utils/imresize.py:32-55): output sample i (1-based
    x = i + 1) sits at u = x/scale + 0.5 (1 - 1/scale); when shrinking the kernel is stretched by 1/scale (antialiasing);
    weights are normalised per output sample; indices are reflected about the borders (symmetric, edge repeated);
    columns that are zero for every output sample are dropped."""
    kernel = _cubic if method == "bicubic" else _triangle
    k_width = 4.0
    if scale < 1:
        h = lambda t: scale * kernel(scale * t)
        kernel_width = k_width / scale
    else:
        h, kernel_width = kernel, k_width
    x = np.arange(1, out_length + 1, dtype=np.float64)
    u = x / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - kernel_width / 2)
    P = int(np.ceil(kernel_width)) + 2
    ind = (left[:, None] + np.arange(P) - 1).astype(np.int32)
    w = h(u[:, None] - ind - 1)
    w = w / w.sum(axis=1, keepdims=True)
    aux = np.concatenate((np.arange(in_length), np.arange(in_length - 1, -1, -1))).astype(np.int32)
    ind = aux[np.mod(ind, aux.size)]
    keep = np.nonzero(np.any(w, axis=0))[0]
    return np.ascontiguousarray(w[:, keep]), np.ascontiguousarray(ind[:, keep])


def imresize(img: np.ndarray, scalar_scale=None, method: str = "bicubic", output_shape=None) -> np.ndarray:
    """utils/imresize.py:104-145 ("vec" mode): float64 arithmetic, the dimension with the smaller scale first, uint8 inputs
    are clipped to [0, 255] and rounded half-to-even after EACH pass (imresize.py:87-91)."""
    if scalar_scale is not None:
        scale = [float(scalar_scale)] * 2
        out_size = [int(np.ceil(scale[k] * img.shape[k])) for k in range(2)]
    else:
        scale = [1.0 * output_shape[k] / img.shape[k] for k in range(2)]
        out_size = list(output_shape)
    B = img.copy()
    two_d = B.ndim == 2
    if two_d:
        B = B[:, :, None]
    for dim in np.argsort(np.array(scale)):
        w, ind = imresize_contributions(img.shape[dim], out_size[dim], scale[dim], method)
        src = B.astype(np.float64)
        if dim == 0:
            out = np.einsum("op,opwc->owc", w, src[ind])
        else:
            out = np.einsum("op,hopc->hoc", w, src[:, ind])
        B = np.around(np.clip(out, 0, 255)).astype(np.uint8) if img.dtype == np.uint8 else out
    return B[:, :, 0] if two_d else B


# ---------------------------------------------------------------------------------------------
# colour tail of test() (train.py:329-341, utils/utils.py:191-204)
# ---------------------------------------------------------------------------------------------
def ycbcr2rgb(x: np.ndarray) -> np.ndarray:
    """utils/utils.py:191-204: inverse BT.601 in fp64, term by term, left to right."""
    mat = np.array([[65.481, 128.553, 24.966], [-37.797, -74.203, 112.0], [112.0, -93.786, -18.214]])
    mat_inv = np.linalg.inv(mat)
    offset = np.matmul(mat_inv, np.array([16, 128, 128]))
    mat_inv = mat_inv * 255
    y = np.zeros(x.shape, dtype="double")
    for k in range(3):
        y[:, :, k] = mat_inv[k, 0] * x[:, :, 0] + mat_inv[k, 1] * x[:, :, 1] + mat_inv[k, 2] * x[:, :, 2] - offset[k]
    return y


def sai_to_rgb8_views(sr_y: np.ndarray, sr_cbcr: np.ndarray, ang: int) -> np.ndarray:
    """train.py:332-335: cat(Y, CbCr) -> ycbcr2rgb -> clip(0,1)*255 -> uint8 (truncation) -> [a1, a2, h, w, 3]."""
    ycbcr = np.concatenate([sr_y[None], sr_cbcr], 0).transpose(1, 2, 0)           # float32 [H, W, 3]
    rgb = (ycbcr2rgb(ycbcr).clip(0, 1) * 255).astype("uint8")
    H, W, _ = rgb.shape
    return np.ascontiguousarray(rgb.reshape(ang, H // ang, ang, W // ang, 3).transpose(0, 2, 1, 3, 4))
