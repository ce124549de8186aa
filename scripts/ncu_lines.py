"""Per-source-line instruction shares of an ncu source page (csv from --print-source cuda,sass) for a range of lines.
usage: ncu_lines.py src.csv file first last [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
f, a, b = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
mn = float(sys.argv[5]) if len(sys.argv) > 5 else 0.2
cur = hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    try: line = int(r[0])
    except ValueError: continue
    i = hdr.index("Instructions Executed")
    try: wi = int(r[i])
    except ValueError: wi = 0
    out.append((cur, line, wi, r[1].strip()[:120]))
tot = sum(o[2] for o in out)
print("total warp instr", tot)
files = {}
for o in out: files[o[0]] = files.get(o[0], 0) + o[2]
print({k: round(v / tot * 100, 1) for k, v in files.items()})
acc = 0
for o in sorted([o for o in out if o[0] == f and a <= o[1] <= b], key=lambda o: o[1]):
    acc += o[2]
    if o[2] / tot * 100 >= mn: print(f"{o[1]:4d} {o[2] / tot * 100:5.2f}% {o[3]}")
print("range total", round(acc / tot * 100, 2), "%")
