import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hdr = None; tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get('Metric Name') == 'gpu__time_duration.sum':
            v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
            v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(u, 1e-6)
            k = d['Kernel Name'][:60]; tot[k] += v; cnt[k] += 1
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]): print('%10.3f ms %5d  %s' % (v, cnt[k], k))
print('total %.3f ms' % sum(tot.values()))
