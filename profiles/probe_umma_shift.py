"""Run the UMMA row-shift probe (csrc/lfsr_debug.cu) and report which descriptor variant works."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LFSR_PROBE_LIB", "1")   # liblfsr_probe.so: probe kernels + debug hooks (not in the product library)
import lfsr_b200
lib = lfsr_b200._native.load()
fn = lib.lfsr_debug_umma_shift
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
a = (torch.arange(256 * 32, dtype=torch.float32).reshape(256, 32) % 2039).cuda()   # exactly representable in tf32
b = torch.eye(32).cuda()
for variant in (0, 1):
    for shift in (0, 1, 2, 3, 5, 8, 9, 34, 35, 70, 127):
        out = torch.zeros(128, 32, device="cuda")
        rc = fn(a.data_ptr(), b.data_ptr(), out.data_ptr(), shift, variant, None)
        torch.cuda.synchronize()
        want = a[shift:shift + 128]
        bad = int((out != want).sum())
        first = ""
        if bad:
            rows = (out != want).any(dim=1).nonzero().flatten()[:6].tolist()
            first = f" first bad rows {rows}; row0 got {out[rows[0]][:4].tolist()} want {want[rows[0]][:4].tolist()}"
        print(f"variant {variant} shift {shift:3d}: rc={rc} mismatches={bad}{first}")
