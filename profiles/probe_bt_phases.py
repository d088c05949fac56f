"""Cycle accounting of the fused BasicTrans kernel per tile: where one epilogue warp (warp 0) and the MMA-issuing lane spend
their time (probe build: liblfsr_probe.so, LFSR_BT_DBG_PTR). usage: python profiles/probe_bt_phases.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dbg = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
os.environ["LFSR_BT_DBG_PTR"] = hex(dbg.data_ptr())
os.environ.setdefault("LFSR_PROBE_LIB", "1")
import lfsr_b200
from lfsr_b200 import kernels as K
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ops = K.default_ops()
torch.manual_seed(1234)
net = lfsr_b200.load_net("EPIT", 5, 4).eval().to("cuda")
al = net._get_packed(torch.device("cuda", 0), ops)["alt"][0]
A, h, W = 5, 32, 160
x = (torch.rand(B, W, W, 64, device="cuda") - 0.5).half()
y = torch.empty_like(x)
p = dict(A=A, S=h, stride_a=h, stride_s=1, stride_b=W * W, stride_p=h * W, stride_q=W, np_=A, nq=h)
call = lambda: ops.basictrans(x, al["bt"][0], al["bt"][1], y, p["A"], p["S"], 5, B, p["np_"], p["nq"], p["stride_a"], p["stride_s"],
                              p["stride_b"], p["stride_p"], p["stride_q"])
call(); torch.cuda.synchronize()
dbg.zero_()
call(); torch.cuda.synchronize()
d = dbg.view(148, 64).double()
tiles = d[:, 32 + 14].clamp(min=1)
per = (d / tiles[:, None]).mean(0)
epi = ["wait R#1 (in-proj; incl. tile turn-around)", "X16 store + LN1 + N store", "wait V^T", "V^T drain", "wait Q", "Q drain", "wait K",
       "K drain", "softmax work (4 heads: exp/store/arrive + O drain)", "wait S (4x)", "softmax load + max (4x)", "wait P_EMPTY (4x)",
       "wait P_EMPTY (last)", "last O drain", "wait R#2 (Wo)", "LN2", "wait F", "F drain", "wait R#3 (FF2)", "X3", "wait y", "y store"]
print(f"epilogue warp 0, cycles per tile (mean over CTAs), batch {B}:")
tot = 0
for i, n in enumerate(epi):
    print(f"  {n:55s} {per[i].item():8.0f}")
    tot += per[i].item()
print(f"  {'sum':55s} {tot:8.0f}")
mma = ["W_FULL (weight ring)", "X_FULL (input tile)", "X_READY", "N_READY", "QKV_READY", "P_FULL (8x)", "O_READY", "N2_READY", "F_READY (2x)", "X3_READY"]
print("MMA lane: cycles per tile waiting on")
w = 0
for i, n in enumerate(mma):
    print(f"  {n:55s} {per[32 + i].item():8.0f}")
    w += per[32 + i].item()
print(f"  total {per[32 + 15].item():.0f} cycles per tile, of which waiting {w:.0f}")
names = ["W_in c0", "W_in c1", "Wv c0", "Wv c1", "Wq c0", "Wq c1", "Wk c0", "Wk c1", "Wo c0", "Wo c1", "W1 n0c0", "W1 n0c1", "W1 n1c0", "W1 n1c1",
         "W2 c0", "W2 c1", "W2 c2", "W2 c3", "Wout"]
print("weight-ring wait per block:", ", ".join(f"{n} {per[44 + i].item():.0f}" for i, n in enumerate(names)))
