import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]
ci={h:i for i,h in enumerate(hdr)}
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot={s:0 for s in stalls}; samples=0
data=[]
def I(x):
    try: return int(x)
    except: return 0
for r in rows[hi+1:]:
    if len(r)<len(hdr): continue
    n=I(r[ci['# Samples']]); samples+=n
    for s in stalls: tot[s]+=I(r[ci[s]])
    data.append((n,r[ci['Source']],I(r[ci['Instructions Executed']]),r))
print('total samples',samples, 'instr rows', len(data), 'warp-instr', sum(d[2] for d in data))
for s,v in sorted(tot.items(), key=lambda x:-x[1])[:9]: print('%-24s %6d %5.1f%%'%(s,v,100*v/max(samples,1)))
print('--- top instructions by samples')
N=int(sys.argv[2]) if len(sys.argv)>2 else 25
for n,src,ie,r in sorted(data,key=lambda x:-x[0])[:N]:
    top=sorted(((I(r[ci[s]]),s) for s in stalls),reverse=True)[:2]
    print('%5d %8d  %-64s %s'%(n,ie,src.strip()[:64],top))
