"""Per-launch table of one bench frame from an ncu launch list (gpu__time_duration.sum): trace / shade launches in order."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if r]
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
ix = {n: i for i, n in enumerate(rows[h])}
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
names, vals = [], []
for r in rows[h + 1:]:
    if len(r) <= ix['Metric Value']:
        continue
    name = r[ix['Kernel Name']].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    if name.startswith(('k_trace<', 'k_shade')):
        names.append(name[:28]); vals.append(float(r[ix['Metric Value']].replace(',', '')) / 1e3)
names, vals = names[skip:], vals[skip:]
for i in range(0, min(len(vals), 45), 3):
    print('  '.join(f'{n:>22s} {v:8.1f}us' for n, v in zip(names[i:i + 3], vals[i:i + 3])))
