"""ORACLE helper, build-container only: import the UNMODIFIED reference from /root/reference.

The reference parses sys.argv at import time (option.py:36, pulled in by utils/utils.py:8) and
imports packages that are absent here (skimage, matplotlib, xlwt, h5py, imageio, fvcore). This
shim registers empty stub modules, injects the scipy restatement of skimage.metrics from
oracle/lf_oracle.py, sets sys.argv and puts /root/reference first on sys.path. Run it in its own
process (make_golden.py does) - this repo's drop-in `model/`, `utils/`, `option.py` share the
reference's module names.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("LFSR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model", "SR"))


def install(argv=("--angRes", "5", "--scale_factor", "4", "--use_pre_ckpt", "", "--device", "cpu")):
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    for name in ("utils", "utils.utils", "model", "model.SR", "option", "train"):
        sys.modules.pop(name, None)
    sys.argv = ["ref_shim"] + list(argv)
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(here)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (repo,)]
    sys.path.insert(0, REF_ROOT)
    sys.path.append(repo)  # for `oracle.*` only; the reference's names win
    # The reference's model/ and model/SR/ have no __init__.py (namespace packages), and a regular package of the
    # same name (this repo's drop-in model/) would beat them whatever the sys.path order: pin both by hand.
    for name, sub in (("model", "model"), ("model.SR", os.path.join("model", "SR")), ("utils", "utils")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF_ROOT, sub)]
        sys.modules[name] = m
    sys.modules["model"].SR = sys.modules["model.SR"]

    def stub(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    for name in ("matplotlib", "matplotlib.pyplot", "xlwt", "h5py", "imageio", "fvcore", "fvcore.nn"):
        if name not in sys.modules:
            stub(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["fvcore"].nn = sys.modules["fvcore.nn"]
    sys.modules["fvcore.nn"].FlopCountAnalysis = None
    sk, skm = stub("skimage"), stub("skimage.metrics")
    sk.metrics = skm
    from oracle import lf_oracle
    skm.peak_signal_noise_ratio = lambda a, b, data_range=1.0: lf_oracle.psnr_view(a, b, data_range)
    skm.structural_similarity = lambda a, b, gaussian_weights=True, data_range=1.0, **kw: lf_oracle.ssim_view(a, b, data_range)


def ref_model(name: str, ang: int, scale: int):
    mod = importlib.import_module("model.SR." + name)
    assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REF_ROOT)), mod.__file__

    class A:
        angRes_in = ang
        angRes_out = ang
        scale_factor = scale
    return mod, mod.get_model(A())
