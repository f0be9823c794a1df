"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line (several launches are
summed): samples, share, file:line, source text.  usage: python scripts/ncu_src_lines.py src.csv [n_top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
cur_file = cur_line = cur_src = None
tot = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] in ('Function Name', 'Line No', 'Kernel Name') or len(r) < 8:
        continue
    if r[0]:
        try:
            cur_line, cur_src = int(r[0]), r[1]
        except ValueError:
            continue
    if r[2]:
        try:
            s = int(float(r[6] or 0))
        except ValueError:
            s = 0
        k = (cur_file, cur_line)
        tot.setdefault(k, [cur_src, 0])
        tot[k][1] += s
total = sum(v[1] for v in tot.values()) or 1
print("total samples", total)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
for (f, ln), (src, s) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:n]:
    print(str(s).rjust(8), "%5.1f%%" % (100.0 * s / total), (f or "")[:18].ljust(18), str(ln).rjust(5), src.strip()[:110])
