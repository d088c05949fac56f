"""Drop-in boundary: plugin discovery, option flags, utils exports, the C-ABI export
list, and (GPU) the test.py entry point end to end on a synthetic scene."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code):
    return subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_cabi_exports_every_declared_symbol():
    import lfsr_b200
    hdr = open(os.path.join(ROOT, "include", "lfsr.h")).read()
    declared = set(re.findall(r"\b(lfsr_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"lfsr_tensor", "lfsr_conv_desc", "lfsr_status", "lfsr_epi_attn_desc"}
    assert declared == set(lfsr_b200._native.SIGNATURES), declared ^ set(lfsr_b200._native.SIGNATURES)
    lib = ctypes.CDLL(lfsr_b200._native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib2 = lfsr_b200._native.load()
    assert lib2.lfsr_abi_version() == 1 and lib2.lfsr_built_for_sm100a() == 1


def test_missing_library_fails_loudly(monkeypatch):
    import lfsr_b200
    monkeypatch.setattr(lfsr_b200._native, "_lib", None)
    monkeypatch.setattr(lfsr_b200._native, "LIB_PATH", "/nonexistent/liblfsr_b200.so")
    with pytest.raises(lfsr_b200.LfsrError):
        lfsr_b200._native.load()


def test_plugin_modules_and_options():
    code = r"""
import sys
sys.argv = ['test.py', '--model_name', 'EPIT', '--angRes', '5', '--scale_factor', '4', '--use_pre_ckpt', '']
import importlib, option
a = option.args
assert a.use_pre_ckpt is False and a.angRes_in == 5 and a.angRes_out == 5 and not hasattr(a, 'angRes')
assert (a.patch_size_for_test, a.stride_for_test, a.minibatch_for_test) == (32, 16, 1)
for name, n in (('MyEfficientLFNet', 547540), ('EPIT', 1470080), ('DistgSSR', 3581568), ('LF_InterNet', 5482688),
                ('MyEfficientLFNetV4_5', 756553)):
    m = importlib.import_module('model.SR.' + name)
    net = m.get_model(a); net.apply(m.weights_init); m.get_loss(a)
    assert sum(p.numel() for p in net.parameters()) == n, name
    net.cpu(); net.eval(); net.load_state_dict(net.state_dict())
import utils.imresize as RI
assert all(hasattr(RI, n) for n in ('imresize', 'deriveSizeFromScale', 'deriveScaleFromSize'))
import utils.utils as U
for sym in ('LFdivide', 'LFintegrate', 'cal_metrics', 'ImageExtend', 'ycbcr2rgb', 'ExcelFile', 'create_dir', 'Logger', 'rearrange', 'np', 'torch', 'os'):
    assert hasattr(U, sym), sym
import sys as _s
_s.argv = ['x', '--use_pre_ckpt', 'False']
import importlib as il; il.reload(option)
assert option.args.use_pre_ckpt is True      # the reference's type=bool quirk (SURVEY 5)
print('ok')
"""
    r = _run(code)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_image_extend_and_colour():
    code = r"""
import sys; sys.argv = ['x']
import torch, numpy as np
import utils.utils as U
x = torch.rand(2, 1, 9, 7)
lr, ud, dg = x.flip(-1), x.flip(-2), x.flip(-1, -2)
ext = torch.cat((torch.cat((dg, ud, dg), -1), torch.cat((lr, x, lr), -1), torch.cat((dg, ud, dg), -1)), -2)
b = [3, 5, 2, 6]
want = ext[:, :, 9 - b[0]: 18 + b[1], 7 - b[2]: 14 + b[3]]
assert torch.equal(U.ImageExtend(x, b), want)
rgb = np.random.RandomState(0).random_sample((5, 6, 3))
back = U.ycbcr2rgb(U.rgb2ycbcr(rgb))
assert np.abs(back - rgb).max() < 1e-9
print('ok')
"""
    r = _run(code)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.gpu
def test_entry_point_end_to_end(tmp_path):
    cmd = [sys.executable, "test.py", "--model_name", "MyEfficientLFNet", "--angRes", "5", "--scale_factor", "4",
           "--use_pre_ckpt", "", "--synthetic", "1", "--synthetic_size", "40", "--path_log", str(tmp_path) + "/",
           "--data_name", "Synthetic"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "The mean psnr on testsets" in r.stdout
    out = tmp_path / "SR_5x5_4x" / "Synthetic" / "MyEfficientLFNet" / "results" / "TEST"
    assert out.is_dir()
    r2 = subprocess.run([sys.executable, "inference.py", "--model_name", "DistgSSR", "--angRes", "5", "--scale_factor", "2",
                         "--use_pre_ckpt", "", "--synthetic", "1", "--synthetic_size", "32", "--path_log", str(tmp_path) + "/",
                         "--data_name", "Synthetic"], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r2.returncode == 0, r2.stderr[-3000:]
    # the BMP tail (train.py:337-341): 25 CodaBench-named views per scene, 24-bit BMP of the x4 / x2 view size
    for root, side in ((out, 160), (tmp_path / "SR_5x5_2x" / "Synthetic" / "DistgSSR" / "results" / "TEST", 64)):
        bmps = sorted(root.rglob("View_*_*.bmp"))
        assert len(bmps) == 25 and bmps[0].name == "View_0_0.bmp", (root, len(bmps))
        head = bmps[7].read_bytes()[:30]
        assert head[:2] == b"BM" and int.from_bytes(head[18:22], "little") == side and int.from_bytes(head[28:30], "little") == 24
