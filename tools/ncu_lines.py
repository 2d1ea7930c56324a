#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` dump by source line: warp stall samples per line of OUR code, top N.
   ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass? > page.csv ; python tools/ncu_lines.py page.csv [N]"""
import csv, sys, collections, re
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, newline="", errors="replace")))
header = None
for i, r in enumerate(rows):
	if any("Sampl" in c for c in r):
		header = r
		start = i + 1
		break
if header is None:
	print("no sampling column found; first rows:", rows[:3])
	sys.exit(1)
print("columns:", header)
col_s = [i for i, c in enumerate(header) if "Sampl" in c and "Not" not in c][0]
col_src = [i for i, c in enumerate(header) if c.strip() in ("Source", "Source Line", "File")] 
col_line = [i for i, c in enumerate(header) if c.strip() in ("Source", "#", "Line")]
agg = collections.Counter()
for r in rows[start:]:
	if len(r) != len(header):
		continue
	try:
		n = float(r[col_s].replace(",", "") or 0)
	except ValueError:
		continue
	key = " | ".join(r[i] for i in range(min(3, len(r))))
	agg[key] += n
total = sum(agg.values()) or 1
for k, v in agg.most_common(top):
	print(f"{v:9.0f} {100*v/total:5.1f}%  {k[:200]}")
