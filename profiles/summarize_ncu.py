"""Turn ncu outputs into the small summaries committed under profiles/ (run here, on files brought back in gpurun_out/).

  python profiles/summarize_ncu.py launches <launches.csv> <first> <last> "<header text>" > profiles/rNN_launch_summary.csv
      per-kernel totals of the `--metrics gpu__time_duration.sum` launch list, launches first..last (the timed steps)
  python profiles/summarize_ncu.py raw <raw.csv> "<command text>" "<kernel text>" <variant> > profiles/rNN_dominant_kernel_ncu.json
      selected metrics of `ncu -i X.ncu-rep --page raw --csv`
"""
import csv
import json
import re
import sys
from collections import defaultdict

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__block_size", "launch__grid_size",
        "launch__cluster_dim_x", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def rows_of(path):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("lfsr::", "").strip()


def launches(path, first, last, header):
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows_of(path):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        i = int(r["ID"])
        if first <= i <= last:
            k = short(r["Kernel Name"])
            tot[k] += float(r["Metric Value"].replace(",", "")) / 1e6
            cnt[k] += 1
    total = sum(tot.values())
    print(f"# {header}")
    print(f"# total {total:.2f} ms over {sum(cnt.values())} launches")
    print("kernel,launches,total_ms,share_pct")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"{k},{cnt[k]},{tot[k]:.3f},{100 * tot[k] / total:.1f}")


def raw(path, command, kernel, variant):
    rs = rows_of(path)
    units, vals = rs[0], rs[1]
    out = {"kernel_variant": variant, "command": command, "kernel": kernel}
    for k in KEEP:
        if k in vals:
            out[k] = {"value": vals[k], "unit": units[k]}
    print(json.dumps(out, indent=1))


def zoo(path, command):
    """every profiled launch of `ncu -i X.ncu-rep --page raw --csv`: the LAST instance of each kernel name (the launch
    after its warm-up), selected metrics + derived DRAM GB/s."""
    rs = rows_of(path)
    units = rs[0]
    last = {}
    for r in rs[1:]:
        last[short(r["Kernel Name"]) + " | grid " + r.get("launch__grid_size", "?") + " smem " +
             r.get("launch__shared_mem_per_block_dynamic", "?")] = r
    out = {"command": command, "kernels": {}}
    for name, r in last.items():
        ent = {}
        for k in KEEP:
            if k in r and r[k] != "":
                ent[k] = {"value": r[k], "unit": units[k]}
        try:
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tmul = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
            byt = sum(float(r[k].replace(",", "")) * mult[units[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            sec = float(r["gpu__time_duration.sum"].replace(",", "")) * tmul[units["gpu__time_duration.sum"]]
            ent["derived_dram_gbs"] = byt / sec / 1e9
        except (KeyError, ValueError):
            pass
        out["kernels"][name] = ent
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "zoo":
        zoo(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
    else:
        raw(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
