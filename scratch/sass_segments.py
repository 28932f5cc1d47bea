import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)==len(hdr)]
tot=sum(int(r[idx['Instructions Executed']]) for r in data)
print('total inst',tot, 'n sass', len(data))
cnts=[int(r[idx['Instructions Executed']]) for r in data]
i=0
segs=[]
while i<len(data):
    j=i
    while j+1<len(data) and abs(cnts[j+1]-cnts[i])<=0.02*max(cnts[i],1): j+=1
    segs.append((i,j,cnts[i]))
    i=j+1
big=[s for s in segs if s[2]*(s[1]-s[0]+1)>0.005*tot]
for s in big:
    n=s[1]-s[0]+1
    print(f"sass {s[0]:5d}-{s[1]:5d} n={n:4d} count={s[2]:.3e} share={s[2]*n/tot:.3f}  samples={sum(int(data[k][idx['# Samples']]) for k in range(s[0],s[1]+1))}")
if len(sys.argv)>3:
    a,b=int(sys.argv[2]),int(sys.argv[3])
    for k in range(a,b+1):
        r=data[k]
        print(k, r[idx['Source']].strip()[:90].ljust(90), r[idx['Instructions Executed']], r[idx['# Samples']], r[idx['L1 Wavefronts Shared']] if 'L1 Wavefronts Shared' in idx else '', r[idx.get('L1 Wavefronts Shared Ideal', 0)])
