import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
d=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki][:60]; v=float(r[vi].replace(",",""))
    if r[ui]=="ns": v/=1e3
    elif r[ui]=="ms": v*=1e3
    d.setdefault(k,[]).append(v)
for k,v in d.items(): print("%-62s n=%3d last=%9.1f us min=%9.1f"%(k,len(v),v[-1],min(v)))
