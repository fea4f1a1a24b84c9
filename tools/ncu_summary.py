"""Summarise an ncu --csv launch list (gpu__time_duration.sum) per kernel/grid."""
import csv, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.OrderedDict()
unit = ''
for row in csv.DictReader(lines):
    try:
        t = float(row['Metric Value'].replace(',', ''))
    except Exception:
        continue
    unit = row['Metric Unit']
    key = (row['Kernel Name'][:70], row.get('Grid Size', ''), row.get('Block Size', ''))
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
for k, (n, t) in agg.items():
    print("%-72s grid %-16s blk %-12s n=%3d avg %12.1f %s  share %5.1f%%" % (k[0], k[1], k[2], n, t / n, unit, 100 * t / tot))
