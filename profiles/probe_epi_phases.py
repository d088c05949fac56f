"""Per-CTA phase time stamps of mel_epi_branch_mma_kernel (probe build, LFSR_PROBE_LIB=1): where a CTA's lifetime goes."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LFSR_PROBE_LIB", "1")
import lfsr_b200
from lfsr_b200 import kernels as K
lib = lfsr_b200._native.load()
ops = K.CudaOps()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
trunk = torch.rand(B, 160, 160, 60, device="cuda")
cat = torch.zeros(B, 160, 160, 60, device="cuda")
w = (torch.rand((2 * 11 + 9) * 18 + 6 * 18 * 18, device="cuda") - 0.5) * 0.3
img = ops.mel_epi_pack(w, 11, "cuda")
t16 = K.alloc_nhwc16(B, 160, 160, 64, "cuda")
t16[..., :60].copy_(trunk)
run = lambda: ops.mel_epi_branch_mma(trunk[..., 40:58], t16[..., 40:58], img, cat[..., 40:58], 11, 5, 0.1)
for _ in range(3):
    run()
dbg = torch.zeros(4096 * 16, dtype=torch.int64, device="cuda")
fn = lib.lfsr_debug_set_em_dbg
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p]
assert fn(dbg.data_ptr()) == 0
run()
torch.cuda.synchronize()
fn(None)
d = dbg.view(4096, 16).cpu().double()
names = ["start", "after setup sync", "ch16/17 staged", "extras done (arrive IN)", "D1 ready (worker 0)", "A2 written", "D2 ready",
         "all done", "MMA: image landed", "MMA: tile landed", "MMA: workers ready", "MMA: block 0 issued", "MMA: block 1 issued",
         "MMA: stage 2 block 0 issued", "MMA: stage 2 block 1 issued"]
rel = d[:, 1:15] - d[:, 0:1]
sel = rel[300:4000]          # steady state (CTAs that started behind others)
print("median cycles since CTA start (CTAs 300..4000):")
for i, nme in enumerate(names[1:]):
    print(f"  {nme:32s} {sel[:, i].median().item():9.0f}")
life = (d[:, 7] - d[:, 0])[300:4000]
print(f"CTA lifetime median {life.median().item():.0f}, p90 {life.quantile(0.9).item():.0f} cycles")
