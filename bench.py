#!/usr/bin/env python
"""bench.py - LF patches/s of the patch-wise LF-SR inference path on N B200s (one rank per GPU).

A step = one pass of the hot path over one synthetic scene that LFdivide cuts into `--batch`
(default 64) 5x5x32x32 patches:  LFdivide -> model forward (batch 64) -> LFintegrate -> PSNR/SSIM.
Workload = BASELINE.json configs[1] (Track-2 model MyEfficientLFNet, 5x5, x4, batch 64, random
weights). `value` times the step with the LR scene already in HBM; `e2e` times the same call with
the LR scene in pinned host memory (H2D inside) and the stitched SR mosaic + metric sums read back
(D2H inside). N > 1: every rank runs its own scenes (weak scaling, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model NAME]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ANG, PATCH, STRIDE = 5, 32, 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="MyEfficientLFNet")
    ap.add_argument("--scale", type=int, default=4)
    ap.add_argument("--batch", type=int, default=64, help="patches per step per GPU (must be a square number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="keep every conv on the fp32 CUDA-core kernels")
    return ap.parse_args()


def scene_side(batch: int) -> int:
    """view size h0 such that LFdivide yields sqrt(batch)^2 patches: numU = (h0 + 15)//16."""
    n = int(round(batch ** 0.5))
    assert n * n == batch, "--batch must be a square number"
    return n * STRIDE


def synthetic_scene(h0: int, scale: int, seed: int):
    rs = np.random.RandomState(seed)
    lr = rs.random_sample((ANG * h0, ANG * h0)).astype(np.float32)
    hr = rs.random_sample((ANG * h0 * scale, ANG * h0 * scale)).astype(np.float32)
    return lr, hr


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
def cpu_reference_rate(model: str, scale: int, patches: int, steps: int, warmup: int):
    """The reference algorithm on the host cores (oracle port: numpy divide/integrate/metrics + the
    torch-CPU restatement of the reference forward, == reference bit-exactly). Returns patches/s."""
    from oracle import lf_oracle, nets as onets, weights
    torch.set_num_threads(os.cpu_count() or 1)
    sd = weights.make_state_dict(model, scale, 1234)
    n = int(round(patches ** 0.5))
    h0 = n * STRIDE
    lr, hr = synthetic_scene(h0, scale, 0)

    def step():
        sub = lf_oracle.lfdivide(lr, ANG, PATCH, STRIDE)
        nu, nv = sub.shape[:2]
        x = torch.from_numpy(sub.reshape(nu * nv, 1, ANG * PATCH, ANG * PATCH))
        ys = [onets.forward(model, x[i:i + 1], sd, ANG, scale) for i in range(nu * nv)]   # minibatch 1: option.py:45
        y = torch.cat(ys).numpy().reshape(nu, nv, ANG * PATCH * scale, ANG * PATCH * scale)
        sr = lf_oracle.to_sai(lf_oracle.lfintegrate(y, ANG, PATCH * scale, STRIDE * scale, h0 * scale, h0 * scale))
        return lf_oracle.cal_metrics(hr, sr, ANG)[:2], nu * nv

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        _, k = step()
        done += k
    dt = time.perf_counter() - t0
    return done / dt, dt / steps * 1e3, done // steps


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    patches = 16           # bounded sample: a 4x4-patch scene per step (~1.5 s of host work on the GPU box)
    rate, ms, per_step = cpu_reference_rate(a.model, a.scale, patches, max(a.steps, 1), a.warmup)
    cores = os.cpu_count() or 1
    sample = f"{per_step}-patch scene (5x5x32x32 views) per step, minibatch 1 as option.py:45, {a.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": "LF patches/sec (5x5x32x32 x%d SR)" % a.scale, "value": rate, "unit": "patches/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the product arm's workload (same pipeline, same model, same patch geometry); each step is a bounded sample of it
        "config": {"workload": f"{a.model} 5x5 x{a.scale}: LFdivide -> forward(patches 5x5x32x32) -> LFintegrate -> PSNR/SSIM "
                               f"(BASELINE configs[1])",
                   "implementation": f"reference algorithm on the host CPU cores (oracle port, torch {torch.__version__})",
                   "sample": sample},
        "cpu_baseline": {"value": rate, "unit": "patches/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# -------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import lfsr_b200
    from lfsr_b200 import kernels as K, lfutils as U

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    lib = lfsr_b200._native.load()
    if a.no_tc:
        K._default_ops = K.CudaOps(use_tc=False)
    ops = K.default_ops()
    torch.manual_seed(1234)                                  # seeded constructor-default init
    net = lfsr_b200.load_net(a.model, ANG, a.scale).eval().to(dev)

    h0 = scene_side(a.batch)
    s = a.scale
    n_scenes = 4                                              # rotate inputs so no step reuses the previous one
    host_lr, dev_lr, dev_hr = [], [], []
    for i in range(n_scenes):
        lr, hr = synthetic_scene(h0, s, 100 * rank + i)
        host_lr.append(torch.from_numpy(lr).pin_memory())
        dev_lr.append(torch.from_numpy(lr).to(dev))
        dev_hr.append(torch.from_numpy(hr).to(dev))
    pz, ss = PATCH * s, STRIDE * s
    _, nu, nv = U.divide_geometry(h0, h0, PATCH, STRIDE)
    assert nu * nv == a.batch
    sub = torch.empty((nu * nv, 1, ANG * PATCH, ANG * PATCH), device=dev)
    # two sets of result buffers: the D2H copy of scene i (copy stream) overlaps the kernels of scene i + 1
    mosaics = [torch.empty((ANG * h0 * s, ANG * h0 * s), device=dev) for _ in range(2)]
    accs = [torch.zeros(2 * ANG * ANG, dtype=torch.float64, device=dev) for _ in range(2)]
    stage_lrs = [torch.empty((ANG * h0, ANG * h0), device=dev) for _ in range(2)]
    host_srs = [torch.empty((ANG * h0 * s, ANG * h0 * s)).pin_memory() for _ in range(2)]
    host_accs = [torch.empty(2 * ANG * ANG, dtype=torch.float64).pin_memory() for _ in range(2)]
    host_sr, host_acc = host_srs[0], host_accs[0]
    copy_stream = torch.cuda.Stream(device=dev)
    done = [None, None]

    def hot_path(lr_dev, hr_dev, mosaic, acc):
        ops.divide_rows(lr_dev, sub, ANG, h0, h0, PATCH, STRIDE, 0, nu)
        sr = net(sub, [ANG, ANG])
        ops.integrate_rows(sr, mosaic, ANG, pz, ss, h0 * s, h0 * s, nu, nv, 0, nu)
        acc.zero_()
        ops.metric_sums(hr_dev, mosaic, ANG, h0 * s, h0 * s, acc)

    def step_resident(i):
        hot_path(dev_lr[i % n_scenes], dev_hr[i % n_scenes], mosaics[0], accs[0])

    def step_e2e(i):
        # every scene: H2D of its LR mosaic from pinned memory, the hot path, D2H of the stitched SR mosaic and the metric
        # sums. Two-deep pipeline: the host consumes the results of scene i - 2 (event wait) right before their buffers
        # are reused, so the D2H of one scene runs under the kernels of the next (the reference consumes them per scene,
        # train.py:322; a serving loop reads them one scene late).
        k = i & 1
        main = torch.cuda.current_stream()
        if done[k] is not None:
            done[k].synchronize()
        stage_lrs[k].copy_(host_lr[i % n_scenes], non_blocking=True)
        hot_path(stage_lrs[k], dev_hr[i % n_scenes], mosaics[k], accs[k])
        ev = torch.cuda.Event()
        ev.record(main)
        copy_stream.wait_event(ev)
        with torch.cuda.stream(copy_stream):
            host_srs[k].copy_(mosaics[k], non_blocking=True)
            host_accs[k].copy_(accs[k], non_blocking=True)
            done[k] = torch.cuda.Event()
            done[k].record(copy_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps, warmup):
        for i in range(warmup):
            step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.lfsr_launch_count() + net.graph_launches     # direct launches + kernels replayed from CUDA graphs
        e0.record()
        for i in range(steps):
            step(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.lfsr_launch_count() + net.graph_launches - l0
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, launches = timed(step_resident, a.steps, max(a.warmup, 3))
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = timed(step_e2e, a.steps, 2)
    value = a.batch * world * a.steps / (ms_total * 1e-3)
    e2e = a.batch * world * a.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel, timed alone with CUDA events on its launch stream ----
    roof = None
    if rank == 0 and hasattr(net, "dominant_kernel"):
        call, info = net.dominant_kernel(a.batch)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        kms = e0.elapsed_time(e1) / reps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        gbs = info["bytes"] / (kms * 1e-3) / 1e9
        # DRAM traffic of this kernel per launch from the committed `ncu --set full` capture (taken at batch 64)
        traffic = None
        try:
            nc = json.load(open(os.path.join(ROOT, "profiles", "r01_dominant_kernel_ncu.json")))
            if nc.get("kernel_variant", "") != ("fused-tail" if "bytes_conv_layer_only" in info else "conv"):
                raise KeyError("capture is of the other kernel variant")
            gb = lambda k: float(nc[k]["value"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[nc[k]["unit"]]
            traffic = (gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")) * a.batch / 64.0
        except (OSError, KeyError, ValueError):
            pass
        roof = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                "traffic": traffic, "kernel": info["name"], "kernel_ms": kms, "algorithmic_bytes": info["bytes"],
                "tflops": info["flops"] / (kms * 1e-3) / 1e12,
                # when the kernel fuses the head conv's contraction, also the fraction counted on this conv layer alone
                "frac_conv_layer_only": (info["bytes_conv_layer_only"] / (kms * 1e-3) / 1e9 / hbm_peak
                                         if "bytes_conv_layer_only" in info else None),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}

    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        rate, ms_cpu, per_step = cpu_reference_rate(a.model, a.scale, 16, 8, 1)     # ~128 patches, 10-20 s of host work
        cpu = {"value": rate, "unit": "patches/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"8 steps of a {per_step}-patch scene (same pipeline, minibatch 1), torch-CPU oracle port"}

    if rank == 0:
        sr_bytes = host_sr.numel() * 4 + host_acc.numel() * 8
        print(json.dumps({
            "metric": "LF patches/sec (5x5x32x32 x%d SR)" % s, "value": value, "unit": "patches/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if ops.use_tc else "f32", "data": "synthetic",
            "config": {"workload": f"{a.model} 5x5 x{s}: LFdivide -> forward(batch {a.batch} patches 5x5x32x32) -> LFintegrate -> "
                                   f"PSNR/SSIM, one {h0}x{h0}-view synthetic scene per GPU per step (BASELINE configs[1])",
                       "patches_per_step_per_gpu": a.batch, "parallelism": f"scene-parallel x{world}",
                       "l2": "activations per step (>6 GB at batch 64) exceed the 126 MB L2; inputs rotate over 4 scenes",
                       "weights": "random init (torch.manual_seed(1234), constructor defaults)"},
            "e2e": {"value": e2e, "unit": "patches/s", "h2d_bytes_per_step": host_lr[0].numel() * 4,
                    "d2h_bytes_per_step": sr_bytes, "ms_per_step": ms_e2e / a.steps,
                    "pipeline": "2-deep: D2H of scene i on a copy stream under the kernels of scene i+1; all copies complete "
                                "inside the timed region"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
