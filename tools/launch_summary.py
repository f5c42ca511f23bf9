"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, re, sys
path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if "gpu__time_duration" not in row["Metric Name"]:
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    if row["Metric Unit"] == "ns":
        v /= 1000.0
    elif row["Metric Unit"] == "ms":
        v *= 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(title)
print("per-launch times are cold-cache and serialised by ncu: compare SHARES, not absolutes")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:28s} launches={c:4d} total_us={t:10.1f} avg_us={t / c:9.1f} share={t / tot:6.3f}")
