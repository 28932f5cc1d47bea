import numpy as np, sys
sys.path.insert(0,'scratch')
from bank_sim_fp import ray_pq, wf, N
rng=np.random.default_rng(2)
dphi=2*np.pi/720
pitches=list(range(36,104,4))
acc={(s,p):0.0 for s in (-1,1) for p in pitches}; cnt={-1:0,1:0}
for t in range(3000):
    phi=rng.uniform(-np.pi/4,np.pi/4); k=rng.integers(0,512)
    u0=rng.integers(0,48)*16; v=rng.integers(0,512)+0.5
    cu=np.arange(u0,u0+16)+0.5; cv=np.full(16,v)
    p1,q1=ray_pq(phi,cu,cv,k); p2,q2=ray_pq(phi+dphi,cu,cv,k)
    p=np.concatenate([p1,p2]); q=np.concatenate([q1,q2])
    if p.max()<0 or p.min()>N or q.max()<0 or q.min()>N: continue
    ip=np.floor(p).astype(int); iq=np.floor(q).astype(int)
    s=1 if (p1[-1]-p1[0])*(q1[-1]-q1[0])>=0 else -1
    cnt[s]+=1
    for pt in pitches: acc[(s,pt)]+=wf(ip,iq,pt)
for s in (-1,1):
    print('sign',s,'n',cnt[s], ' '.join(f"{pt}:{acc[(s,pt)]/cnt[s]:.2f}" for pt in pitches))
