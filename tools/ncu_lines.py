#!/usr/bin/env python3
"""Aggregate an `ncu -i rep --page source --print-source cuda,sass --csv` dump: warp stall samples per CUDA source line.
   python tools/ncu_lines.py page.csv [N]
The dump lists, file by file, a row per source line (with the line's totals) followed by its SASS rows."""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, newline="", errors="replace")))
agg = collections.Counter()
stall = {}
file = "?"
header = None
for r in rows:
	if len(r) >= 2 and r[0] == "File Path":
		file = r[1].split("/")[-1]
		continue
	if len(r) > 5 and r[0] == "Line No":
		header = r
		si = header.index("# Samples")
		stall_cols = [(i, c) for i, c in enumerate(header) if c.startswith("stall_") and "Not Issued" not in c]
		continue
	if header is None or len(r) < len(header) or r[0] == "":
		continue
	try:
		n = float(r[si].replace(",", "") or 0)
	except ValueError:
		continue
	key = f"{file}:{r[0]}: {r[1].strip()[:110]}"
	agg[key] += n
	best = sorted(((float(r[i].replace(",", "") or 0), c) for i, c in stall_cols if r[i] not in ("", "-")), reverse=True)[:2]
	stall[key] = " ".join(f"{c[6:]}:{v:.0f}" for v, c in best if v > 0)
total = sum(agg.values()) or 1
print("total samples", total)
for k, v in agg.most_common(top):
	print(f"{v:7.0f} {100 * v / total:5.1f}%  {k}   [{stall.get(k, '')}]")
