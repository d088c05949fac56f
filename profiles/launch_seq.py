"""the launches of the LAST forward in an ncu launch list, in order, with durations (us).
usage: python profiles/launch_seq.py launches.csv [n_forwards=8]"""
import csv, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.DictReader(rows))
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ours = [x for x in r if "lfsr" in x["Kernel Name"] or "conv_tc" in x["Kernel Name"] or "bt::" in x["Kernel Name"] or "_kernel" in x["Kernel Name"]]
per = len(ours) // nf
for i, x in enumerate(ours[-per:]):
    k = x["Kernel Name"].split("(")[0].replace("void ", "").replace("lfsr::", "")
    print(f"{i:3d} {float(x['Metric Value'].replace(',', '')) / 1e3:9.1f} us  {k} grid {x.get('Grid Size', '?')} block {x.get('Block Size', '?')}")
