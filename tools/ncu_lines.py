"""Aggregate an `ncu --page source --csv` export by CUDA source line, using nvdisasm -g line info of the cubin.
usage: python tools/ncu_lines.py src.csv lines.sass file.cu [lo hi]   (lo/hi: restrict to a source line range)"""
import csv, re, collections, sys
srccsv, sass, cu = sys.argv[1:4]
lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 10**9)
line_of = {}; cur = None
for l in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # keep the OUTERMOST non-inlined frame of our file when present
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', l)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if r and r[col['# Samples']].isdigit()]
base = min(int(r[col['Address']], 16) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.defaultdict(collections.Counter)
for r in data:
    off = int(r[col['Address']], 16) - base
    ln = line_of.get(off); n = int(r[col['# Samples']])
    agg[ln]['n'] += n; agg[ln]['inst'] += int(r[col['Instructions Executed']] or 0)
    for s in stalls: agg[ln][s] += int(r[col[s]] or 0)
src = open(cu).read().split('\n'); name = cu.split('/')[-1]
tot = sum(a['n'] for a in agg.values())
sel = {k: v for k, v in agg.items() if k and k[0] == name and lo <= k[1] <= hi} if len(sys.argv) > 5 else agg
print("total samples", tot, "selected", sum(a['n'] for a in sel.values()))
ssum = collections.Counter()
for a in sel.values():
    for s in stalls: ssum[s] += a[s]
print("stall mix of selection:", [(s[6:], v) for s, v in ssum.most_common(6)])
order = sorted(sel.items(), key=(lambda x: x[0][1]) if len(sys.argv) > 5 else (lambda x: -x[1]['n']))
for ln, a in order[:60 if len(sys.argv) <= 5 else 400]:
    top = sorted(((s, a[s]) for s in stalls), key=lambda x: -x[1])[:2]
    text = src[ln[1] - 1].strip()[:64] if ln and ln[0] == name else str(ln)
    print(f"{str(ln[1] if ln else None):>5} {a['n']:7d} {100 * a['n'] / tot:5.1f}% inst {a['inst']:9d} {top[0][0][6:]:>10}:{top[0][1]:6d} {top[1][0][6:]:>10}:{top[1][1]:5d} | {text}")
