"""Per-source-line stall samples from `ncu --page source --print-source cuda,sass --csv` output."""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="Line No"][0]
h=rows[hdr]
wi=[j for j,x in enumerate(h) if x.startswith("Warp Stall Sampling (All")][0]
ii=h.index("Instructions Executed")
tot=0; out=[]
for r in rows[hdr+1:]:
    if len(r)<=wi or r[2]!="-": continue          # source-line rows have "-" as address
    try: v=int(r[wi])
    except: continue
    tot+=v; out.append((v,int(r[0]),r[ii],r[1][:105]))
print("total samples",tot)
n=int(sys.argv[2]) if len(sys.argv)>2 else 40
for v,l,ie,s in sorted(out,reverse=True)[:n]: print("%6d %5.1f%% inst=%8s L%4d: %s"%(v,100*v/tot,ie,l,s))
