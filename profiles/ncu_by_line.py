"""Aggregate an ncu `--page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K | python profiles/ncu_by_line.py [top]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
hdr = None
out = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and r and r[0].strip().isdigit():
        try:
            out.append((int(r[0]), r[1].strip()[:90], int(r[hdr.index("# Samples")] or 0), int(r[hdr.index("Instructions Executed")] or 0)))
        except ValueError:
            pass
ti = sum(o[3] for o in out) or 1
ts = sum(o[2] for o in out) or 1
print(f"total warp-instructions {ti}  samples {ts}")
print("by instructions:")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"  L{o[0]:4d} inst {100*o[3]/ti:5.1f}%  samples {100*o[2]/ts:5.1f}%  {o[1]}")
print("by stall samples:")
for o in sorted(out, key=lambda o: -o[2])[:top]:
    print(f"  L{o[0]:4d} inst {100*o[3]/ti:5.1f}%  samples {100*o[2]/ts:5.1f}%  {o[1]}")
