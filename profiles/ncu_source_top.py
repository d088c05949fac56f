"""Hot spots of an `ncu --page source --csv` dump (SASS view): instruction totals and the lines with the most stall samples.
usage: ncu_source_top.py <source.csv> [n_lines] [min_inst_executed]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
h = rows[hi]
ix = {k: h.index(k) for k in h}
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
def num(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except ValueError: return 0.0
tot_inst = sum(num(r, 'Instructions Executed') for r in body)
tot_samp = sum(num(r, '# Samples') for r in body)
print(f"SASS lines {len(body)}, warp instructions executed {tot_inst:.0f}, stall samples {tot_samp:.0f}")
print("-- top lines by samples: idx, samples%, inst_executed, main stall, SASS")
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
order = sorted(range(len(body)), key=lambda i: -num(body[i], '# Samples'))[:n]
for i in sorted(order):
    r = body[i]
    top = max(stalls, key=lambda k: num(r, k))
    print(f"{i:5d} {100 * num(r, '# Samples') / tot_samp:5.1f}% {num(r, 'Instructions Executed'):10.0f} {top:18s} {r[ix['Source']][:110]}")
