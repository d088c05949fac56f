"""Issuer time stamps of the persistent mel_epi_branch_mma_kernel (probe build, LFSR_PROBE_LIB=1): cycles per tile.
(The per-phase stamps of the first, non-persistent version are in git history: commit "All-MMA EPI kernel v2".)"""
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
dbg = torch.zeros(2 * 4096 * 16, dtype=torch.int64, device="cuda")
fn = lib.lfsr_debug_set_em_dbg
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p]
assert fn(dbg.data_ptr()) == 0
run()
torch.cuda.synchronize()
fn(None)
ph = dbg[65536:].view(4096, 16).cpu().double()[:148]
d = dbg[:65536].view(4096, 16).cpu().double()[:148]
per = (d[:, 1:12] - d[:, 0:11])            # cycles between the issuer's tile starts (persistent kernel, tiles 0..11 of each CTA)
print("issuer: cycles per tile, median over the CTAs, tiles 1..11:", [int(per[:, i].median().item()) for i in range(11)])
print(f"steady state (tiles 4..11): {per[:, 3:].median().item():.0f} cycles per tile")
t0 = d[:, 8:9]                              # issuer's start of tile 8
names = {0: "issuer: inputs + tensor memory ready", 1: "issuer: 62 tap MMAs issued", 2: "issuer: stage 2 of tile 7 issued", 3: "issuer: side channel ready",
         4: "issuer: extra MMAs + commits issued", 8: "worker 0: iteration start", 9: "worker 0: epilogue 2 of tile 7 done",
         10: "worker 0: D1 of tile 8 ready", 11: "worker 0: epilogue 1 done (A2 ready)", 14: "side channel: inputs ready", 15: "side channel: extras done"}
print("tile 8, median cycles since the issuer started it:")
for k, nme in names.items():
    print(f"  {nme:40s} {(ph[:, k:k + 1] - t0).median().item():9.0f}")
