"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel time of one production frame."""
import csv, sys
from collections import Counter
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
h = rows[hdr]; rows = rows[hdr + 1:]
ki = h.index('Kernel Name'); vi = h.index('Metric Value')
names = [r[ki] for r in rows]; vals = [float(r[vi].replace(',', '')) for r in rows]
idx = [i for i, n in enumerate(names) if n.startswith('shade_kernel')]
# a frame = the launches between two consecutive shade kernels that contain the production cone kernel; take the last frame of the
# most common length (the bench also issues stand-alone mip builds and instrumented cone kernels between some frames)
wins = [(idx[j - 1] + 1, idx[j] + 1) for j in range(1, len(idx)) if any('cone_kernel_fast' in n for n in names[idx[j - 1] + 1:idx[j] + 1])]
common = Counter(e - s for s, e in wins).most_common(1)[0][0]
s, e = [w for w in wins if w[1] - w[0] == common][-1]
tot = sum(vals[s:e])
for n, v in zip(names[s:e], vals[s:e]):
    print(f"{n[:70]:70s} {v / 1000:8.1f} us {100 * v / tot:5.1f}%")
print('total', tot / 1000, 'us (cold-cache, serialised launches)')
