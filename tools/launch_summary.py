"""Summarise an ncu launch-list CSV (gpu__time_duration.sum per launch): the last forward of the run (from its g_a head kernel)."""
import csv, collections, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
recs = collections.OrderedDict()
for row in csv.DictReader(lines):
    d = recs.setdefault(int(row['ID']), {'name': row['Kernel Name']})
    d[row['Metric Name']] = float(row['Metric Value'].replace(',', ''))
L = list(recs.values()); half = len(L) // 2; S = L[half:]
heads = [i for i, d in enumerate(L) if 'ga_head_kernel' in d['name']]
if heads:                      # the last forward of the run: from its g_a head kernel on, this repo's kernels only
    S = [d for d in L[heads[-1]:] if not d['name'].startswith(('at::', 'void at::'))]
tot = sum(d['gpu__time_duration.sum'] for d in S)
print(f"launches in the measured forward: {len(S)}; summed kernel time {tot/1e3:.1f} us")
agg = collections.defaultdict(lambda: [0, 0.0])
for d in S:
    n = re.sub(r'\(.*', '', d['name']).replace('void ', ''); agg[n][0] += 1; agg[n][1] += d['gpu__time_duration.sum']
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/tot*100:6.2f}% {t/1e3:10.1f} us  x{c:4d}  {n[:100]}")
if len(sys.argv) > 2:
    for i, d in enumerate(S):
        print(i, d['name'][:60], d['gpu__time_duration.sum'] / 1e3)
