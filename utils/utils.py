"""Drop-in for the reference's utils/utils.py on the inference path: LFdivide / LFintegrate /
cal_metrics run on liblfsr_b200 kernels (lfsr_b200.lfutils); ImageExtend, ycbcr2rgb, rgb2ycbcr,
create_dir, Logger and ExcelFile keep the reference's behaviour (utils/utils.py:14-88,137-204) for
the callers that still use them (BMP/xls tail)."""
import logging
import os  # noqa: F401  (re-exported through `from utils.utils import *`, test.py:4)
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F  # noqa: F401
from einops import rearrange  # noqa: F401

from option import args
from lfsr_b200.lfutils import LFdivide, LFintegrate, cal_metrics  # noqa: F401


def ImageExtend(Im, bdr):
    """symmetric mirror extension of [N,C,h,w] by bdr=[top,bottom,left,right] (utils.py:137-149),
    written as an index gather instead of flip+cat."""
    h, w = Im.shape[-2:]
    mirror = lambda i, n: torch.where(i < 0, -i - 1, torch.where(i >= n, 2 * n - 1 - i, i))
    ys = mirror(torch.arange(-bdr[0], h + bdr[1], device=Im.device), h)
    xs = mirror(torch.arange(-bdr[2], w + bdr[3], device=Im.device), w)
    return Im[..., ys, :][..., xs]


_YCBCR = np.array([[65.481, 128.553, 24.966], [-37.797, -74.203, 112.0], [112.0, -93.786, -18.214]])


def rgb2ycbcr(x):
    y = x.astype("double") @ _YCBCR.T + np.array([16.0, 128.0, 128.0])
    return y / 255.0


def ycbcr2rgb(x):
    inv = np.linalg.inv(_YCBCR)
    offset = inv @ np.array([16, 128, 128])
    return x.astype("double") @ (inv * 255).T - offset


class ExcelFile:
    """evaluation.xls writer (utils.py:14-44); needs xlwt, otherwise falls back to a .csv with the same rows."""

    def __init__(self):
        self.rows = [("Datasets", "Scenes", "PSNR", "SSIM")]
        try:
            import xlwt
            self.xlsx_file = xlwt.Workbook()
            self.worksheet = self.xlsx_file.add_sheet("sheet1", cell_overwrite_ok=True)
            for i, t in enumerate(self.rows[0]):
                self.worksheet.write(0, i, t)
        except ImportError:
            self.xlsx_file = self
            self.worksheet = None
        self.sum = 1

    def write_sheet(self, test_name, LF_name, psnr_iter_test, ssim_iter_test):
        for n, p, s in zip(LF_name, psnr_iter_test, ssim_iter_test):
            self.add_sheet(test_name, n, p, s)
        self.add_sheet(test_name, "average", float(np.mean(psnr_iter_test)), float(np.mean(ssim_iter_test)))
        self.sum += 1

    def add_sheet(self, test_name, LF_name, psnr, ssim):
        row = (test_name, LF_name, "%.6f" % psnr, "%.6f" % ssim)
        self.rows.append(row)
        if self.worksheet is not None:
            for i, t in enumerate(row):
                self.worksheet.write(self.sum, i, t)
        self.sum += 1

    def save(self, path):  # only reached through the csv fallback (self.xlsx_file is self)
        with open(str(path).replace(".xls", ".csv"), "w") as f:
            for r in self.rows:
                f.write(",".join(str(c) for c in r) + "\n")


def create_dir(args):
    """log/SR_<A>x<A>_<s>x/<data>/<model>/{checkpoints,results} (utils.py:59-78)."""
    task = "SR_%dx%d_%dx" % (args.angRes_in, args.angRes_in, args.scale_factor)
    log_dir = Path(args.path_log) / task / args.data_name / args.model_name
    ckpt, res = log_dir / "checkpoints", log_dir / "results"
    for d in (ckpt, res):
        d.mkdir(parents=True, exist_ok=True)
    return log_dir, ckpt, res


class Logger:
    def __init__(self, log_dir, args):
        self.logger = logging.getLogger(args.model_name)
        self.logger.setLevel(logging.INFO)
        fh = logging.FileHandler("%s/%s.txt" % (log_dir, args.model_name))
        fh.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
        self.logger.addHandler(fh)

    def log_string(self, s):
        if args.local_rank <= 0:
            self.logger.info(s)
            print(s)
