import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 45
cur=None; hdr=None
agg=collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] in ("File Name","File Path"): cur=r[1].split('/')[-1]; continue
    if r[0]=="Line No": hdr=r; iex=hdr.index("Instructions Executed"); ism=hdr.index("# Samples"); continue
    if hdr is None or cur is None: continue
    try:
        ln=int(r[0]); ex=int(r[iex] or 0); sm=int(r[ism] or 0)
    except: continue
    k=(cur,ln)
    a=agg.setdefault(k,[0,0,r[1]])
    a[0]+=ex; a[1]+=sm
tot=sum(a[0] for a in agg.values()); ts=sum(a[1] for a in agg.values())
print("total instr",tot,"samples",ts)
items=sorted(agg.items(), key=lambda kv:-kv[1][1])[:top]
for (f,ln),(ex,sm,src) in sorted(items):
    print(f"{f}:{ln:4d} instr%={100*ex/tot:5.1f} samp%={100*sm/ts:5.1f}  {src.strip()[:110]}")
