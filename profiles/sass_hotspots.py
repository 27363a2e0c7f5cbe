import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
# multiple kernels instances may be concatenated: take first block
hdr=None; data=[]
for r in rows:
    if r and r[0]=='Address':
        if hdr is not None: break
        hdr=r; continue
    if hdr is not None and len(r)==len(hdr): data.append(r)
ia=hdr.index('Address'); isrc=hdr.index('Source'); iex=hdr.index('Instructions Executed')
tot=sum(float(r[iex]) for r in data)
print('total instr executed', tot, 'sass lines', len(data))
N=int(sys.argv[2]) if len(sys.argv)>2 else 60
for start in range(0,len(data),N):
    ch=data[start:start+N]; a=sum(float(r[iex]) for r in ch)
    ops={}
    for r in ch:
        t=r[isrc].split(); 
        if not t: continue
        op=t[1] if t[0].startswith('@') else t[0]
        ops[op]=ops.get(op,0)+float(r[iex])
    top=sorted(ops.items(), key=lambda kv:-kv[1])[:6]
    if a/tot>0.004: print(f"{start:5d} {a/tot*100:6.2f}%  " + ', '.join(f"{k}:{v/tot*100:.1f}" for k,v in top))
