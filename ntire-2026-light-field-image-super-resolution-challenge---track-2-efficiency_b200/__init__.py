"""lfsr-b200: B200-native (sm_100a) kernels and drop-in host modules for BasicLFSR's patch-wise
light-field SR inference path (LFdivide -> get_model(args).forward -> LFintegrate -> PSNR/SSIM).

The directory name carries the reference repository's name and is not a Python identifier; import
it as ``lfsr_b200`` (thin alias package at the repo root).
"""
from . import _native
from ._native import LfsrError, build_native
from . import kernels
from . import lfutils
from . import scene

__all__ = ["_native", "kernels", "LfsrError", "build_native", "load_net", "NETS"]

#: reference model name (model/SR/<name>.py) -> module under lfnets/
NETS = {
    "MyEfficientLFNet": "my_efficient_lfnet",
    "MyEfficientLFNetV4_5": "my_efficient_lfnet_v4_5",
    "EPIT": "epit",
    "DistgSSR": "distgssr",
    "LF_InterNet": "lf_internet",
}


def net_module(model_name: str):
    """the host module mirroring /root/reference/model/SR/<model_name>.py"""
    import importlib
    if model_name not in NETS:
        raise LfsrError(f"model {model_name!r} is outside the accelerated path; known: {sorted(NETS)}")
    return importlib.import_module(f"{__name__}.lfnets.{NETS[model_name]}")


def load_net(model_name: str, ang: int = 5, scale: int = 4):
    class _Args:
        angRes_in = ang
        angRes_out = ang
        scale_factor = scale
    return net_module(model_name).get_model(_Args())
