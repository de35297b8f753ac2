"""Print the per-launch table (time, issue %, warps active %) of an `ncu --csv` launch list; last N launches."""
import csv, sys
path = sys.argv[1]; last = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); mi = h.index('Metric Name'); vi = h.index('Metric Value'); ii = h.index('ID')
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(',', ''))
tot = 0
for (i, k), m in sorted(d.items())[-last:]:
    t = m['gpu__time_duration.sum'] / 1000; tot += t
    print(i, k[:72].ljust(72), round(t, 1), round(m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0), 1),
          round(m.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0), 1))
print('total us', round(tot, 1))
