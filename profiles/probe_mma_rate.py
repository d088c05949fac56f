"""cycles per tf32 tcgen05.mma (M=128) as a function of N, accumulator dependence and K addressing."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LFSR_PROBE_LIB", "1")   # liblfsr_probe.so: probe kernels + debug hooks (not in the product library)
import lfsr_b200
lib = lfsr_b200._native.load()
fn = lib.lfsr_debug_mma_rate
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for N in (32, 64, 128, 224):
    for alt in (0, 4, 8, 16):
        for kmode in (0,):
            res = []
            for chain in (16, 64):
                fn(out.data_ptr(), N, chain, alt, kmode, None); torch.cuda.synchronize()
                fn(out.data_ptr(), N, chain, alt, kmode, None); torch.cuda.synchronize()
                res.append(out.tolist())
            per = (res[1][1] - res[0][1]) / 48.0
            print(f"N={N:3d} alternate={alt} kmode={kmode}: chain16 issue/total = {res[0]}, chain64 = {res[1]}  -> {per:.1f} cycles per MMA (marginal)")
