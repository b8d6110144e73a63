"""Summarise an ncu --page source --csv dump: hottest SASS instructions and stall reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ci["# Samples"]]), r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
idx = {id(r): i for i, (_, r) in enumerate(data)}
for n, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("%6d %5.1f%% line%5d  %-70s %s exec=%s" % (n, 100.0 * n / tot, idx[id(r)], r[ci["Source"]].strip()[:70],
                                             st, r[ci["Instructions Executed"]]))
