"""Aggregate an ncu source-page dump by code region (function) and by line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:K | python profiles/ncu_regions.py SRC.cu [top]"""
import collections
import csv
import re
import sys

src_path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# function regions of the source file: a line that starts a __device__/__global__ function opens a region
regions = []
name_re = re.compile(r"(?:__device__|__global__)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
lines = open(src_path).read().split("\n")
for i, l in enumerate(lines, 1):
    if l.startswith(("__device__", "__global__", "template")) or "__global__ void" in l:
        mm = name_re.search(re.sub(r"__launch_bounds__\([^)]*\)", "", " ".join(lines[i - 1:i + 2])))
        if mm:
            regions.append((i, mm.group(1)))
    else:   # "// ---- phase N ..." comments split a long function into sub-regions
        ph = re.search(r"// ---- (phase [0-9a-z +]+?)[,:]", l)
        if ph and regions:
            regions.append((i, regions[-1][1].split(" / ")[0] + " / " + ph.group(1).strip()))
base = src_path.split("/")[-1]


def region(f, l):
    if f != base:
        return f
    cur = "top"
    for start, nm in regions:
        if start <= l:
            cur = nm
        else:
            break
    return cur


rows = list(csv.reader(sys.stdin))
cur = hdr = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].strip().isdigit() and cur:
        try:
            s = int(r[hdr.index("# Samples")] or 0)
            i = int(r[hdr.index("Instructions Executed")] or 0)
        except ValueError:
            continue
        key = (cur, int(r[0]), r[1].strip()[:110])
        agg[key][0] += s
        agg[key][1] += i
        for name in hdr:
            if name.startswith("stall_") and "Not Issued" not in name:
                try:
                    agg[key][2][name] += int(r[hdr.index(name)] or 0)
                except ValueError:
                    pass
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
print(f"samples {ts}  warp-instructions {ti}")
reg, regi = collections.Counter(), collections.Counter()
for k, v in agg.items():
    reg[region(k[0], k[1])] += v[0]
    regi[region(k[0], k[1])] += v[1]
for k, v in reg.most_common(28):
    print(f"  {k:30s} samples {100 * v / ts:5.1f}%  inst {100 * regi[k] / ti:5.1f}%")
st = collections.Counter()
for v in agg.values():
    st.update(v[2])
print("stalls:", [(k, round(100 * v / ts, 1)) for k, v in st.most_common(8)])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {k[0][:16]}:{k[1]:4d} samp {100 * v[0] / ts:5.2f}% inst {100 * v[1] / ti:5.2f}% {[(a[6:], b) for a, b in v[2].most_common(2)]} | {k[2]}")
