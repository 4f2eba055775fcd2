#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    nvdisasm -c -g kernel.cubin > k.sass
    python tools/ncu_by_line.py sass.csv k.sass <kernel-substring> <file.cu>

Matches the instructions of both listings by order (same cubin), sums executed warp
instructions and stall samples per source line and prints the lines by cost."""
import csv
import re
import sys

sass_csv, disasm, kernel, src = sys.argv[1:5]
# 1. line number of every instruction of the kernel, in address order
lines, cur, inside = [], None, False
inline_depth = None
for text in open(disasm):
    if text.startswith("//---") and ".text." in text:
        inside = kernel in text
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', text)
    if m:
        cur = (m.group(1), int(m.group(2)), m.group(3))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", text):
        lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[hdr_i + 1:]
assert len(body) == len(lines), (len(body), len(lines))
agg = {}
tot_inst = tot_samp = 0
for r, ln in zip(body, lines):
    inst = int(r[ix["Instructions Executed"]])
    samp = int(r[ix["Warp Stall Sampling (All Samples)"]])
    key = ln[1] if ln and ln[0].endswith(src) else ("%s:%d" % (ln[0].split("/")[-1], ln[1]) if ln else "?")
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += inst
    a[1] += samp
    a[2] += 1
    tot_inst += inst
    tot_samp += samp
text = open(src if "/" in src else "/root/repo/annealing-sign-problem_b200/csrc/" + src).read().splitlines()
print("total warp instructions %d, samples %d" % (tot_inst, tot_samp))
for key, (inst, samp, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 45]:
    code = text[key - 1].strip()[:90] if isinstance(key, int) and key <= len(text) else ""
    print("%-22s %5.1f%% inst %5.1f%% stall  %3d sass | %s" % (key, 100.0 * inst / tot_inst, 100.0 * samp / max(tot_samp, 1), n, code))
