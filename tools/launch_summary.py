import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split('(')[0]
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1e3
    elif r[ui] == 'ms': v *= 1e3
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("launches", len(rows) - 1, "total ms", tot / 1e3)
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:55s} n={c:4d} total={t/1e3:9.3f} ms  avg={t/c:10.1f} us  share={t/tot*100:5.1f}%")
