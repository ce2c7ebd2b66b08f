"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py in.csv out.csv"""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
tot = collections.OrderedDict()
for r in rows[1:]:
    try:
        name = r[ci["Kernel Name"]]
        val = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
    except Exception:
        continue
    if unit in ("ns", "nsecond"):
        val /= 1e3
    elif unit in ("ms", "msecond"):
        val *= 1e3
    short = re.sub(r"\(.*", "", name).replace("adsr::", "").replace("<unnamed>::", "").replace("void ", "")
    n, t = tot.get(short, (0, 0.0))
    tot[short] = (n + 1, t + val)
total = sum(t for _, t in tot.values())
with open(sys.argv[2], "w") as f:
    f.write("kernel,launches,total_us,share\n")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{t:.1f},{t / total:.4f}\n")
print(open(sys.argv[2]).read())
