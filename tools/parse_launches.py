import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); gi = hdr.index('Grid Size')
agg = collections.OrderedDict(); total = 0
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki]); t = float(r[vi].replace(',', '')) / 1e3
    total += t
    a = agg.setdefault(name, [0, 0.0, []]); a[0] += 1; a[1] += t; a[2].append(round(t, 1))
print("total us %.1f  launches %d" % (total, len(rows) - 1))
for k, (n, t, l) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {100*t/total:5.1f}% {n:4d}x  {k[:70]:70s} {l if n <= 14 else ''}")
