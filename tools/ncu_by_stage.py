#!/usr/bin/env python
"""Like ncu_by_line.py but sums over named line ranges: python tools/ncu_by_stage.py sass.csv k.sass kernel file.cu name:lo-hi ..."""
import csv, re, sys
sass_csv, disasm, kernel, src = sys.argv[1:5]
stages = []
for spec in sys.argv[5:]:
    name, rng = spec.split(":")
    lo, hi = rng.split("-")
    stages.append((name, int(lo), int(hi)))
lines, cur, inside = [], None, False
for text in open(disasm):
    if text.startswith("//---") and ".text." in text:
        inside = kernel in text
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', text)
    if m:
        # inlined intrinsics: keep attributing to the last line of OUR file
        if m.group(1).endswith(src):
            cur = int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", text):
        lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ix = {k: i for i, k in enumerate(rows[h])}
body = rows[h + 1:]
assert len(body) == len(lines)
tot = {}
ti = ts = 0
for r, ln in zip(body, lines):
    inst = int(r[ix["Instructions Executed"]]); samp = int(r[ix["Warp Stall Sampling (All Samples)"]])
    name = "other"
    for n, lo, hi in stages:
        if ln is not None and lo <= ln <= hi:
            name = n
            break
    a = tot.setdefault(name, [0, 0]); a[0] += inst; a[1] += samp; ti += inst; ts += samp
for n, (i, s) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%-12s %5.1f%% inst  %5.1f%% stall samples   %.0f inst per warp-tile (312500 tiles)" % (n, 100.0 * i / ti, 100.0 * s / ts, i / 312500.0))
