"""Forward time per minibatch size with and without CUDA-graph replay; run on the GPU box.
usage: python profiles/run_minibatch.py [model]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
model = sys.argv[1] if len(sys.argv) > 1 else "MyEfficientLFNet"
net = lfsr_b200.load_net(model, 5, 4).eval().to("cuda")
common = sys.modules[[c for c in type(net).__mro__ if c.__name__ == "LFNetBase"][0].__module__]   # the module forward() reads
for B in (1, 4, 16, 64):
    x = torch.rand(B, 1, 160, 160, device="cuda")
    row = []
    for use in (False, True):
        common.USE_CUDA_GRAPH = use
        for _ in range(3):
            net(x)
        torch.cuda.synchronize()
        n = 20
        t0 = time.perf_counter()
        for _ in range(n):
            y = net(x)
        torch.cuda.synchronize()
        row.append((time.perf_counter() - t0) / n * 1e3)
    print(f"{model} minibatch {B:3d}: direct launches {row[0]:7.3f} ms/forward ({B / row[0] * 1e3:7.0f} patches/s) | "
          f"graph replay {row[1]:7.3f} ms/forward ({B / row[1] * 1e3:7.0f} patches/s)")
