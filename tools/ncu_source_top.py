"""Top stall sites of an `ncu --page source --csv` dump (developer aid).  usage: ncu_source_top.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[ix["# Samples"]])
    except ValueError: continue
    data.append((n, r))
tot = sum(n for n, _ in data)
print("total samples", tot)
agg = {s: 0 for s in stalls}
for n, r in data:
    for s in stalls:
        try: agg[s] += int(r[ix[s]])
        except ValueError: pass
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for n, r in sorted(data, key=lambda t: -t[0])[:N]:
    top = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{n:7d} {100.0*n/tot:5.1f}%  {r[ix['Address']][-6:]}  {r[ix['Source']][:70]:70s} {top}")
