import csv,collections,sys
rows=list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki][:100]; agg.setdefault(k,[]).append(v)
for k,v in agg.items(): print(len(v), round(sum(v)/len(v)/1000,1),'us', k)
