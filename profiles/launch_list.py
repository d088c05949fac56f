"""per-kernel totals of ONE forward from an `ncu --metrics gpu__time_duration.sum --csv` launch list of
`LFSR_CUDA_GRAPH=0 python profiles/run_minibatch2.py MODEL SCALE BATCH` (2 warm-up + 5 timed forwards + 1 allocation pass).
usage: python profiles/launch_list.py launches.csv [n_forwards=8]"""
import csv, collections, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.DictReader(rows))
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ours = [x for x in r if "lfsr" in x["Kernel Name"] or "conv_tc" in x["Kernel Name"] or "bt::" in x["Kernel Name"] or "_kernel" in x["Kernel Name"]]
per = len(ours) // nf
tot, cnt, shp = collections.Counter(), collections.Counter(), {}
for x in ours[-per:]:
    k = x["Kernel Name"].split("(")[0].replace("void ", "").replace("lfsr::", "") + " grid " + x.get("Grid Size", "?").replace(" ", "")
    tot[k] += float(x["Metric Value"].replace(",", "")) / 1e6
    cnt[k] += 1
print(f"# launches per forward {per}, serialised cold-cache sum {sum(tot.values()):.2f} ms")
print("kernel,launches,total_ms,share_pct")
s = sum(tot.values())
for k, v in tot.most_common(25):
    print(f"{k},{cnt[k]},{v:.3f},{100 * v / s:.1f}")
