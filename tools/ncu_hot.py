"""Hot SASS lines of an `ncu --page source --csv` export: executed share, stall-sample share, main stall reason.
usage: python tools/ncu_hot.py src.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
hdr = rows[1]
end = next((i for i in range(2, len(rows)) if rows[i] and rows[i][0] in ('Address', 'Kernel Name')), len(rows))
ia, isrc, ie, iss = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:end]:
    if len(r) <= ie: continue
    stalls = {hdr[i]: int(r[i] or 0) for i in stall_cols}
    data.append((r[isrc].strip(), int(r[ie] or 0), int(r[iss] or 0), stalls))
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
print(f"instructions {tot}, samples {ts}, lines {len(data)}")
agg = {}
for d in data:
    for k, v in d[3].items(): agg[k] = agg.get(k, 0) + v
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for i, d in enumerate(data):
    if d[1] > tot * thr / 100 or d[2] > ts * thr / 100:
        top = max(d[3].items(), key=lambda kv: kv[1])
        print(f"{i:5d} {d[0][:64]:64s} {100*d[1]/tot:5.2f}% exec {100*d[2]/ts:5.2f}% samp  {top[0]}={top[1]}")
