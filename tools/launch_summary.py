"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict(); n = 0; tot = 0.0
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    name = re.sub(r'\(.*', '', r[ki]).replace('sd::', '').replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; n += 1; tot += v
print("total %.0f us over %d launches" % (tot, n))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%8.0f us %5.1f%%  n=%3d  avg %7.1f us  %s" % (v, 100 * v / tot, c, v / c, k[:100]))
