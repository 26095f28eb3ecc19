"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the hottest SASS instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
tot = {s: 0 for s in stalls}; data = []
def num(x):
    try: return int(float(x))
    except Exception: return 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address": continue
    n = num(r[idx["# Samples"]])
    data.append((n, r[idx["Source"]].strip(), r))
    for s in stalls: tot[s] += num(r[idx[s]])
T = max(1, sum(tot.values()))
print("total samples", sum(d[0] for d in data), " instructions", len(data))
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:28s} {v:8d} {100 * v / T:5.1f}%")
print("top instructions by samples:")
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for n, src, r in sorted(data, key=lambda d: -d[0])[:topn]:
    top = sorted(((num(r[idx[s]]), s.replace("stall_", "")) for s in stalls), reverse=True)[:2]
    print(f"  {n:7d}  {src[:64]:64s} {top}")
