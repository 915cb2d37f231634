"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (and optionally grid)."""
import collections, csv, re, sys
path = sys.argv[1]; by_grid = len(sys.argv) > 2
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("mpl::", "")
    v = float(row["Metric Value"].replace(",", "")) / 1e3
    key = (name, row["Grid Size"]) if by_grid else name
    agg[key][0] += 1; agg[key][1] += v; tot += v
print(f"total {tot/1e3:.2f} ms, {sum(n for n,_ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print(f"{t:9.1f} us {100*t/tot:5.1f}% n={n:5d} avg={t/n:8.2f}  {str(k)[:100]}")
