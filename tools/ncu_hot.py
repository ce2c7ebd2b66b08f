"""Top stall sites of an ncu --page source --csv dump (SASS view): python tools/ncu_hot.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for idx, r in enumerate(rows[2:]):
    try:
        s = int(r[ci["# Samples"]])
    except Exception:
        continue
    data.append((s, idx, r))
tot = sum(s for s, _, _ in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for s, idx, r in sorted(data, key=lambda x: -x[0])[:n]:
    top = sorted(((int(r[ci[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{s:6d} {100.0*s/tot:5.1f}%  #{idx:5d} {r[ci['Source']][:90]:90s} {top}")
