"""cycles per tcgen05.mma (M = 128) when the chain is issued like conv_tc_kernel issues (converged warp, elect.sync, uniform
descriptor adds): N sweep x {tf32, f16 over the same bytes} x {fresh A and B every MMA, B repeated, A and B repeated}.
usage: python profiles/probe_mma_rate2.py"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LFSR_PROBE_LIB", "1")   # liblfsr_probe.so: probe kernels + debug hooks (not in the product library)
import lfsr_b200
lib = lfsr_b200._native.load()
fn = lib.lfsr_debug_mma_rate2
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for kind, kname in ((0, "tf32 K=8 "), (1, "f16  K=16")):
    for N in (32, 64, 128, 224, 256):
        row = []
        for afix, bfix in ((0, 0), (0, 1), (1, 1)):
            res = []
            for chain in (32, 128):
                for _ in range(2):
                    fn(out.data_ptr(), N, chain, kind, afix, bfix, None); torch.cuda.synchronize()
                res.append(out.tolist())
            row.append(((res[1][1] - res[0][1]) / 96.0, (res[1][0] - res[0][0]) / 96.0))
        print(f"{kname} N={N:3d}: cycles per MMA total (issue-side) | fresh A+B {row[0][0]:6.1f} ({row[0][1]:5.1f}) | "
              f"B repeated {row[1][0]:6.1f} ({row[1][1]:5.1f}) | A and B repeated {row[2][0]:6.1f} ({row[2][1]:5.1f})")

# cost of a stage boundary between groups of MMAs (N = 64 and 224, tf32): cycles per GROUP minus the MMAs' own time
for N, per in ((64, 48), (224, 112)):
    for grp in (4, 8, 16):
        row = []
        for bmode in (0, 1, 2, 3):
            res = []
            for chain in (64, 256):
                for _ in range(2):
                    fn(out.data_ptr(), N, chain, 0, (bmode << 4) | (grp << 8), 0, None); torch.cuda.synchronize()
                res.append(out.tolist())
            row.append((res[1][1] - res[0][1]) / (192.0 / grp))
        print(f"tf32 N={N:3d} groups of {grp:2d} MMAs ({grp * per} cycles of MMA): cycles per group | one elect block {row[0]:7.1f} | "
              f"re-elect + syncwarp {row[1]:7.1f} | + commit {row[2]:7.1f} | + completed wait + fence {row[3]:7.1f}")
