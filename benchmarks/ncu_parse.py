import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
d={}
for r in rows[1:]:
    if len(r)<=vi: continue
    d.setdefault((int(r[ii]),r[ki][:36]),{})[r[mi]]=r[vi]
for k,v in sorted(d.items()): print(k, v)
