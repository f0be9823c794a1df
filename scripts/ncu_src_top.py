"""Summarise an `ncu --page source --csv` export: hottest SASS instructions and stall-reason totals.
usage: python scripts/ncu_src_top.py src.csv [n_top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]


def S(r, k):
    try:
        return int(float(r[ix[k]] or 0))
    except ValueError:
        return 0


tot = sum(S(r, "# Samples") for r in data)
ex = sum(S(r, "Instructions Executed") for r in data)
print("instructions", len(data), "total samples", tot, "warp-instr executed", ex)
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("stall totals:", sorted([(sum(S(r, h) for r in data), h) for h in st], reverse=True)[:8])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for r in sorted(data, key=lambda r: -S(r, "# Samples"))[:n]:
    reasons = sorted([(S(r, h), h[6:]) for h in st], reverse=True)[:2]
    print(str(S(r, "# Samples")).rjust(6), str(S(r, "Instructions Executed")).rjust(9),
          r[ix["Source"]].strip()[:64].ljust(64), reasons)
