"""Shared-memory wavefronts per tap instruction of the TMA-staged FP kernel, cfg 3 geometry,
for different lane mappings and row pitches (simulation; banks = word address mod 32)."""
import numpy as np
rng=np.random.default_rng(1)
N=512; SOD=4.0; SDD=6.0; pu=2.8125/768; pv=1.875/512; U=768; V=512; vox=1.0/N
def ray_pq(phi, cu, cv, k):
    # y-marching group (|phi| <= 45deg): m=y, p=x, q=z. returns index coords (p,q) at slice k
    s=np.array([np.sin(phi)*SOD,-np.cos(phi)*SOD,0.0])
    dc=np.array([-np.sin(phi)*(SDD-SOD),np.cos(phi)*(SDD-SOD),0.0])
    eu=np.array([np.cos(phi),np.sin(phi),0.0])*pu; ev=np.array([0,0,1.0])*pv
    pix=dc[None,:]+(cu[:,None]-U/2)*eu[None,:]+(cv[:,None]-V/2)*ev[None,:]
    d=pix-s[None,:]
    ym=(k+0.5-N/2)*vox
    t=(ym-s[1])/d[:,1]
    x=s[0]+t*d[:,0]; z=s[2]+t*d[:,2]
    return x/vox+N/2-0.5, z/vox+N/2-0.5
def wf(ip,iq,pitch):
    w=iq*pitch+ip; b=w%32; mx=0
    for kk in np.unique(b): mx=max(mx,len(set(w[b==kk])))
    return mx
res={}
cnt=0
dphi=2*np.pi/720
for t in range(4000):
    phi=rng.uniform(-np.pi/4,np.pi/4); k=rng.integers(0,512)
    u0=rng.integers(0,24)*32; v=rng.integers(0,512)+0.5
    # mapping A: 32 consecutive u, one angle
    cu=np.arange(u0,u0+32)+0.5; cv=np.full(32,v)
    p,q=ray_pq(phi,cu,cv,k)
    if p.max()<0 or p.min()>N or q.max()<0 or q.min()>N: continue
    ip=np.floor(p).astype(int); iq=np.floor(q).astype(int)
    sgn=np.sign((p[-1]-p[0])*(q[-1]-q[0]))
    for pitch in (64,68,72,76,80,60):
        res[('A',pitch)]=res.get(('A',pitch),0)+wf(ip,iq,pitch)
    res[('A','adapt')]=res.get(('A','adapt'),0)+(wf(ip,iq,68) if sgn>=0 else wf(ip,iq,60))
    # mapping B: 16 consecutive u x 2 angles
    for half in (0,1):
        cu=np.arange(u0+16*half,u0+16*half+16)+0.5; cv=np.full(16,v)
        p1,q1=ray_pq(phi,cu,cv,k); p2,q2=ray_pq(phi+dphi,cu,cv,k)
        p=np.concatenate([p1,p2]); q=np.concatenate([q1,q2])
        ip=np.floor(p).astype(int); iq=np.floor(q).astype(int)
        for pitch in (64,68,72,76,80,60):
            res[('B',pitch)]=res.get(('B',pitch),0)+0.5*wf(ip,iq,pitch)
        sgn=np.sign((p1[-1]-p1[0])*(q1[-1]-q1[0]))
        res[('B','adapt')]=res.get(('B','adapt'),0)+0.5*(wf(ip,iq,68) if sgn>=0 else wf(ip,iq,60))
    # mapping C: 16 u x 2 rows (v, v+4?) same angle -> rows r and r+1 of the thread block
    cu=np.concatenate([np.arange(u0,u0+16),np.arange(u0,u0+16)])+0.5; cv=np.concatenate([np.full(16,v),np.full(16,v+1)])
    p,q=ray_pq(phi,cu,cv,k); ip=np.floor(p).astype(int); iq=np.floor(q).astype(int)
    for pitch in (64,68,72,80,48,56):
        res[('C',pitch)]=res.get(('C',pitch),0)+wf(ip,iq,pitch)
    cnt+=1
for kk in sorted(res, key=str): print(kk, round(res[kk]/cnt,3))
