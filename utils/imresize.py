"""Drop-in for the reference's utils/imresize.py (MATLAB-style antialiased bicubic / bilinear resize used by the data
generators, Generate_Data_for_inference.py:74-84): `imresize` keeps the reference's signature and return types and runs
its two separable fp64 passes on liblfsr_b200 (lfsr_resample_f64); the size helpers are plain arithmetic."""
from math import ceil

from lfsr_b200.lfutils import imresize, resize_contributions  # noqa: F401


def deriveSizeFromScale(img_shape, scale):
    return [int(ceil(scale[k] * img_shape[k])) for k in range(2)]


def deriveScaleFromSize(img_shape_in, img_shape_out):
    return [1.0 * img_shape_out[k] / img_shape_in[k] for k in range(2)]
