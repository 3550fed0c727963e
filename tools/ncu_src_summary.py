"""Summarise an `ncu --page source --csv` dump: samples per opcode class and the hottest instructions."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
# a dump of several launches repeats the ("Kernel Name", ...) + header lines: keep section `sys.argv[3]` (default 0)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = rows[starts[sec]:starts[sec + 1]]
print(rows[0][1])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot = 0
by_op = collections.Counter(); by_stall = collections.Counter(); hot = []
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    s = int(r[ix["# Samples"]] or 0)
    tot += s
    op = r[ix["Source"]].strip().split()
    opn = op[0] if not op[0].startswith("@") else op[1]
    by_op[opn.split(".")[0]] += s
    for c in stall_cols:
        by_stall[c] += int(r[ix[c]] or 0)
    hot.append((s, r[ix["Address"]][-5:], r[ix["Source"]].strip()[:90], {c[6:]: int(r[ix[c]]) for c in stall_cols if int(r[ix[c]] or 0) > 0.2 * max(s, 1)}))
print("total samples", tot, " instructions", len(rows) - 2)
print("by opcode:", [(k, round(100 * v / tot, 1)) for k, v in by_op.most_common(14)])
print("by stall :", [(k[6:], round(100 * v / tot, 1)) for k, v in by_stall.most_common(10)])
for h in sorted(hot, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(h)
