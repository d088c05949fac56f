"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP; legacy HMMA must be absent), from `cuobjdump -sass` of the
shipped library. Runs without a GPU.  usage: python profiles/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections, glob, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else glob.glob(os.path.join(root, "*_b200", "liblfsr_b200.so"))[0]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pats = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "ELECT",
        "FFMA2", "HMMA", "HGMMA", "LDGSTS", "REDUX", "SHFL"]
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    per[cur]["_total"] += 1
    for p in pats:
        if op == p or op.startswith(p + "."):
            if p == "UTCHMMA" and op.startswith("UTCHMMA.2CTA"):
                continue
            per[cur][p] += 1
print(f"# cuobjdump -sass {os.path.basename(lib)} (sm_100a): instruction counts per kernel; HMMA/HGMMA = legacy tensor paths (must be 0)")
tot = collections.Counter()
for fn, c in per.items():
    dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem)
    cols = " ".join(f"{p}={c[p]}" for p in pats if c[p])
    print(f"{dem:70s} total={c['_total']:6d} {cols}")
    tot.update(c)
print("# library totals: " + " ".join(f"{p}={tot[p]}" for p in pats))
