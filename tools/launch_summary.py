"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel time of one production frame."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
h = rows[hdr]; rows = rows[hdr + 1:]
ki = h.index('Kernel Name'); vi = h.index('Metric Value')
names = [r[ki] for r in rows]; vals = [float(r[vi].replace(',', '')) for r in rows]
idx = [i for i, n in enumerate(names) if n.startswith('shade_kernel')]
for j in range(len(idx) - 1, 0, -1):
    s, e = idx[j - 1] + 1, idx[j] + 1
    if any('cone_kernel_fast' in n for n in names[s:e]):
        break
tot = sum(vals[s:e])
for n, v in zip(names[s:e], vals[s:e]):
    print(f"{n[:70]:70s} {v / 1000:8.1f} us {100 * v / tot:5.1f}%")
print('total', tot / 1000, 'us (cold-cache, serialised launches)')
