"""profiles/rNN_parity_errors.md from the JSON records tests/parity.py writes when MPL_PARITY_LOG is set:
   MPL_PARITY_LOG=gpurun_out/parity.jsonl python -m pytest tests -m gpu -q ; python tools/parity_table.py gpurun_out/parity.jsonl"""
import collections
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
by = collections.OrderedDict()
for r in rows:
    t = r["test"].split("::")
    key = (t[0].replace("tests/", ""), t[-1].split("[")[0])
    by.setdefault(key, []).append(r)
print("| file | test | checks | worst measured rel. error (check) | in bf16 ulps (2^-8) | stated rtol | flipped rows |")
print("|---|---|---|---|---|---|---|")
for (f, t), rs in by.items():
    w = max(rs, key=lambda r: r["rel"] / max(r["rtol"], 1e-30))
    flips = sum(r.get("rows_flipped") or 0 for r in rs)
    nrows = sum(r.get("rows") or 0 for r in rs)
    print(f"| {f} | {t} | {len(rs)} | {w['rel']:.2e} ({w['check'][:60]}) | {w['rel'] / 2 ** -8:.2f} | {w['rtol']:g} | "
          f"{str(flips) + ' / ' + str(nrows) if nrows else '-'} |")
