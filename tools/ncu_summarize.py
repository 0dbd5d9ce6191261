import csv,sys,collections
rows=list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ii=h.index("ID")
d=collections.OrderedDict()
for r in rows[1:]:
    try: d.setdefault(r[ki][:70],[]).append(float(r[vi].replace(',','')))
    except: pass
tot=0
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])):
    print(f"{sum(v)/1000:10.1f} us total  n={len(v):4d}  avg={sum(v)/len(v)/1000:8.2f} us  {k}")
    tot+=sum(v)
print("total us",tot/1000)
