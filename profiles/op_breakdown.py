"""Per-op device-time breakdown of one forward (CUDA events around every C-ABI call).
usage: python profiles/op_breakdown.py [model] [batch]   (run on the GPU box)"""
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lfsr_b200
from lfsr_b200 import kernels as K


class ProfilingOps(K.CudaOps):
    def __init__(self):
        super().__init__(use_tc=True)
        self.records = []
        self.lib_calls = collections.Counter()

    def _wrap(self, name, fn, label):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before = self.lib.lfsr_launch_count()
        e0.record()
        fn()
        e1.record()
        self.records.append((name, label, e0, e1))

    def conv(self, x, pc, out, **kw):
        tc = "tc" if (pc.w_tc is not None) else "f32"
        io = ("h" if x.dtype == torch.float16 else "f") + ">" + ("f" if out is not None else "") + ("h" if kw.get("out16") is not None else "")
        extra = "".join(f" +{k}" for k in ("res", "mul", "in_scale", "tail") if kw.get(k) is not None) + (" +shuffle" if kw.get("shuffle", (1, 1, 0))[0] > 1 else "")
        label = f"conv {pc.kh}x{pc.kw} {pc.cin}->{pc.cout} s{pc.stride[0]} d{pc.dil[0]} @{x.shape[1]}x{x.shape[2]} [{tc} {io}]{extra}"
        self._wrap("conv", lambda: K.CudaOps.conv(self, x, pc, out, **kw), label)

    def dwconv(self, x, w, out, kh, kw, **k2):
        self._wrap("dwconv", lambda: K.CudaOps.dwconv(self, x, w, out, kh, kw, **k2), f"dwconv {kh}x{kw} d{k2.get('dil', (1, 1))[0]} c{x.shape[3]} @{x.shape[1]}x{x.shape[2]}")

    def dwconv_multi(self, x, out, branches):
        lab = "+".join(f"{b['kh']}x{b['kw']}d{b.get('dil', (1, 1))[0]}c{b['c']}" for b in branches)
        self._wrap("dwconv_multi", lambda: K.CudaOps.dwconv_multi(self, x, out, branches), f"dwconv_multi {lab} @{x.shape[1]}x{x.shape[2]}")

    def mel_epi_branch(self, x, w, out, klen, dil, slope, tc=False):
        self._wrap("mel_epi_branch", lambda: K.CudaOps.mel_epi_branch(self, x, w, out, klen, dil, slope, tc=tc), f"mel_epi_branch{'_tc' if tc else ''} c{x.shape[3]} @{x.shape[1]}x{x.shape[2]}")

    def mel_epi_branch_mma(self, x, x16, img, out, klen, dil, slope):
        self._wrap("mel_epi_branch_mma", lambda: K.CudaOps.mel_epi_branch_mma(self, x, x16, img, out, klen, dil, slope), f"mel_epi_branch_mma c{x.shape[3]} @{x.shape[1]}x{x.shape[2]}")

    def block_mean(self, x, out, bh, bw):
        self._wrap("block_mean", lambda: K.CudaOps.block_mean(self, x, out, bh, bw), f"block_mean {bh}x{bw} c{x.shape[3]} @{x.shape[1]}")

    def sa_modulate(self, *a, **kw):  # noqa
        self._wrap("sa_modulate", lambda: K.CudaOps.sa_modulate(self, *a, **kw), "sa_modulate" + (" + fp16 copy" if kw.get("out16") is not None else ""))

    def ang_expand(self, x, w, res, out, A, *a, **kw):
        self._wrap("ang_expand", lambda: K.CudaOps.ang_expand(self, x, w, res, out, A, *a, **kw), f"ang_expand c{x.shape[3]}->{out.shape[3]} x{A}x{A} @{x.shape[1]}x{x.shape[2]}")

    def pooled_mlp(self, x, out, *a, **kw):
        self._wrap("pooled_mlp", lambda: K.CudaOps.pooled_mlp(self, x, out, *a, **kw), f"pooled_mlp c{x.shape[3]} @{x.shape[1]}x{x.shape[2]}{' pooled' if kw.get('pool') else ''}")

    def to_f16(self, *a):
        self._wrap("to_f16", lambda: K.CudaOps.to_f16(self, *a), "to_f16")

    def scale_add(self, *a):
        self._wrap("scale_add", lambda: K.CudaOps.scale_add(self, *a), f"scale_add c{a[0].shape[3]} @{a[0].shape[1]}")

    def tap_gather(self, *a):
        self._wrap("tap_gather", lambda: K.CudaOps.tap_gather(self, *a), f"tap_gather @{a[0].shape[1]}x{a[0].shape[2]}")

    def interp(self, *a):
        self._wrap("interp", lambda: K.CudaOps.interp(self, *a), "interp")

    def layernorm(self, *a):
        self._wrap("layernorm", lambda: K.CudaOps.layernorm(self, *a), "layernorm")

    def epi_attention(self, *a):
        self._wrap("epi_attention", lambda: K.CudaOps.epi_attention(self, *a), "epi_attention")


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "MyEfficientLFNet"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    dev = torch.device("cuda:0")
    net = lfsr_b200.load_net(model, 5, 4).eval()
    net = net.to(dev)
    ops = ProfilingOps()
    net.set_backend(ops)
    x = torch.rand(batch, 1, 160, 160, device=dev)
    for _ in range(2):
        net(x)
    torch.cuda.synchronize()
    ops.records.clear()
    net(x)
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    total = 0.0
    for name, label, e0, e1 in ops.records:
        ms = e0.elapsed_time(e1)
        total += ms
        a = agg.setdefault(label, [0, 0.0])
        a[0] += 1
        a[1] += ms
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    print(f"{model} batch {batch}: {len(ops.records)} launches, {total:.2f} ms summed device time")
    for label, (cnt, ms) in rows:
        print(f"  {ms:9.3f} ms  {100 * ms / total:5.1f}%  x{cnt:<3d} {label}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"model": model, "batch": batch, "total_ms": total, "rows": [[l, c, m] for l, (c, m) in rows]},
              open(f"gpurun_out/op_breakdown_{model}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
