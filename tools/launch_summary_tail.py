import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
last = rows[1:][-int(sys.argv[2]):]
agg = collections.OrderedDict()
tot = 0
for r in last:
    v = float(r[vi].replace(",", "")); 
    if r[ui] == "ns": v /= 1e3
    elif r[ui] == "ms": v *= 1e3
    k = r[ki][:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
for k, (n, v) in agg.items(): print(f"{v:8.1f} us n={n:3d} avg={v/n:6.2f}  {k}")
print("total", tot)
