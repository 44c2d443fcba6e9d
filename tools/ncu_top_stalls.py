"""Top stalled SASS instructions from `ncu -i rep --page source --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(float(r[ix['# Samples']] or 0) for r in data)
print("total samples", tot)
agg = {c: sum(float(r[ix[c]] or 0) for r in data) for c in stall_cols}
print("by reason:", sorted(((round(v), k) for k, v in agg.items() if v > 0), reverse=True)[:8])
top = sorted(data, key=lambda r: -float(r[ix['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    reasons = sorted(((float(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print("%6s %5.1f%%  %-70s %s" % (r[ix['# Samples']], 100 * float(r[ix['# Samples']] or 0) / tot, r[ix['Source']][:70], [(c, int(v)) for v, c in reasons if v > 0]))
