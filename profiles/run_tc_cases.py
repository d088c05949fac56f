"""Run every TC_CASES entry of tests/test_kernels_gpu.py in its own process (a CUDA fault is sticky) and report."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [root, os.path.join(root, "tests")]
import test_kernels_gpu as t
for i, c in enumerate(t.TC_CASES):
    code = ("import sys; sys.path[:0]=['.','tests']; import torch, opref, test_kernels_gpu as t; "
            "torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False; "
            f"t._run_tc_case(opref.RefOps(), t.TC_CASES[{i}]); torch.cuda.synchronize(); print('ok')")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    status = "ok" if r.returncode == 0 and "ok" in r.stdout else "FAIL " + (r.stderr.strip().splitlines() or ["?"])[-1][:150]
    print(i, c, status, flush=True)
