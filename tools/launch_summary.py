"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel time for one train step."""
import csv, collections, re, sys
path = sys.argv[1]; per_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; data = rows[hi + 1:]
ki, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
recs = []
for r in data:
    if len(r) <= mi: continue
    v = float(r[mi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    recs.append((re.sub(r'\(.*', '', r[ki]).replace('<unnamed>::', '').replace('void ', ''), v))
idx = [i for i, (n, _) in enumerate(recs) if 'adamw_flat' in n]
if len(idx) >= 2: step = recs[idx[0] + 1: idx[1] + 1]
elif per_step: step = recs[-per_step:]
else: step = recs[idx[0] + 1 - (len(recs) - idx[0] - 1) - 1000:] if False else recs[-(idx[0] + 1):] if idx else recs
tot = sum(v for _, v in step)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in step:
    agg[n][0] += 1; agg[n][1] += v
print(f'{len(recs)} launches captured; one step = {len(step)} launches, {tot / 1e3:.3f} ms (cold-cache, serialised)')
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f'{v / 1e3:8.3f} ms {100 * v / tot:5.1f}% x{c:3d}  {n[:90]}')
