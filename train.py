"""Inference half of the reference's train.py: only `test()` (train.py:286-347), which test.py
imports (`from train import test`). Training (train.py:20-282) is outside this repository's scope.

`test()` keeps the reference's signature and return value but runs the patch loop on the device:
LFdivide -> batched forward (args.minibatch patches per call instead of minibatch_for_test = 1 with
a host round trip per patch) -> LFintegrate -> PSNR/SSIM, all on liblfsr_b200 kernels; one D2H copy
of the stitched SR mosaic per scene feeds the unchanged YCbCr->RGB->BMP tail."""
import numpy as np
import torch
from einops import rearrange

from utils.utils import ycbcr2rgb
from lfsr_b200 import scene as _scene


def _write_views(save_dir, name, sr_sai_y, cbcr, ang):
    import imageio
    d = save_dir.joinpath(name)
    d.mkdir(exist_ok=True)
    ycbcr = torch.cat((sr_sai_y, cbcr), dim=1)
    rgb = (ycbcr2rgb(ycbcr.squeeze().permute(1, 2, 0).numpy()).clip(0, 1) * 255).astype("uint8")
    views = rearrange(rgb, "(a1 h) (a2 w) c -> a1 a2 h w c", a1=ang, a2=ang)
    for i in range(ang):
        for j in range(ang):
            imageio.imwrite(str(d) + "/View_%d_%d.bmp" % (i, j), views[i, j])


def test(test_loader, device, net, args, save_dir=None):
    names, psnrs, ssims = [], [], []
    net.eval()
    for Lr_SAI_y, Hr_SAI_y, Sr_SAI_cbcr, data_info, LF_name in test_loader:
        ang = int(data_info[0][0].item()) if torch.is_tensor(data_info[0]) else int(data_info[0])
        lr = Lr_SAI_y.squeeze().to(device, non_blocking=True)
        hr = Hr_SAI_y.squeeze().to(device, non_blocking=True)
        with torch.no_grad():
            psnr, ssim, sr = _scene.test_scene(net, lr, hr, ang, args.scale_factor, args.patch_size_for_test,
                                               args.stride_for_test, getattr(args, "minibatch", 64))
        psnrs.append(psnr)
        ssims.append(ssim)
        names.append(LF_name[0])
        if save_dir is not None:
            try:
                _write_views(save_dir, LF_name[0], sr.cpu()[None, None], Sr_SAI_cbcr, args.angRes_out)
            except ImportError:
                np.save(str(save_dir.joinpath(LF_name[0] + "_Sr_SAI_y.npy")), sr.cpu().numpy())  # imageio absent
    return psnrs, ssims, names
