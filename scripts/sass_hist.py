"""Executed-instruction histogram per kernel from `ncu --page source --csv --print-source sass`."""
import csv
import sys
from collections import Counter, OrderedDict

path = sys.argv[1]
kern = None
tables = OrderedDict()
hdr = None
for row in csv.reader(open(path)):
    if len(row) >= 2 and row[0] == "Kernel Name":
        kern = row[1][:90]
        tables.setdefault(kern, Counter())
        hdr = None
        continue
    if row and row[0] == "Address":
        hdr = row
        si, ei = hdr.index("Source"), hdr.index("Instructions Executed")
        continue
    if hdr and kern and len(row) > ei:
        try:
            n = int(row[ei])
        except ValueError:
            continue
        ins = row[si].strip()
        op = ins.split()[0] if ins else "?"
        if op.startswith("@"):
            op = ins.split()[1]
        tables[kern][op.split(".")[0]] += n
seen = set()
for k, c in tables.items():
    key = (k, sum(c.values()))
    if key in seen:
        continue
    seen.add(key)
    tot = sum(c.values())
    print(k, "total warp-instr", tot)
    print("   ", ", ".join(f"{o}:{n / tot * 100:.1f}%" for o, n in c.most_common(14)))
