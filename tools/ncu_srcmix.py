"""Dynamic instruction mix (and hottest instructions) from an `ncu --page source --csv --print-source sass` export."""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
print(rows[0][1][:70])
h = rows[1]; si = h.index('Source'); ei = h.index('Instructions Executed'); smp = h.index('# Samples')
data = [(r[si].strip(), int(r[ei]), int(r[smp])) for r in rows[2:] if len(r) > ei and r[ei].isdigit()]
tot = sum(e for _, e, _ in data); ts = sum(s for *_, s in data)
print("total warp instr", tot, "samples", ts, "static", len(data))
byop = collections.Counter(); bysmp = collections.Counter()
for s, e, sm in data:
    op = re.sub(r'^@!?U?P\d+\s+', '', s).split()[0]
    op = op if op.startswith('IMAD.MOV') else op.split('.')[0]
    byop[op] += e; bysmp[op] += sm
for op, c in byop.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    print(f"{op:14s} {c:10d} {100*c/tot:5.1f}%  samples {100*bysmp[op]/ts:5.1f}%")
