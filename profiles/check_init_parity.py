"""Parity of the CUDA forward against the oracle under each model's own weights_init (not the golden weight scaling):
relative max error and output magnitude. usage: python profiles/check_init_parity.py [model ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from oracle import nets as onets
names = sys.argv[1:] or ["MyEfficientLFNetV4_5", "MyEfficientLFNet"]
for name in names:
    torch.manual_seed(7)
    net = lfsr_b200.load_net(name, 5, 4).eval()
    net.apply(lfsr_b200.net_module(name).weights_init)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.rand(1, 1, 80, 80)
    y_or = onets.forward(name, x, sd, 5, 4)
    net = net.to("cuda")
    y = net(x.cuda()).cpu()
    mag = y_or.abs().max().item()
    from lfsr_b200 import kernels as K
    net.set_backend(K.CudaOps(use_tc=False))
    y32 = net(x.cuda()).cpu()
    print(f"{name}: oracle |y|max {mag:.4g}, finite {bool(torch.isfinite(y).all())}, max|cuda - oracle| {(y - y_or).abs().max().item():.3e} "
          f"(relative {(y - y_or).abs().max().item() / max(mag, 1e-9):.2e}); fp32-only kernels: relative "
          f"{(y32 - y_or).abs().max().item() / max(mag, 1e-9):.2e}")
