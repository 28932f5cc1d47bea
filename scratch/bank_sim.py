import numpy as np
rng=np.random.default_rng(0)
N=512; vox=1/N; SOD=4.0; SDD=6.0; pu=2.8125/768; pv=1.875/512; U=768; V=512
def proj(x,y,z,phi):
    # source at (sin*SOD, -cos*SOD), detector centre opposite; standard circular cone
    s=np.array([np.sin(phi)*SOD,-np.cos(phi)*SOD,0.0])
    # ray dir central: -s/|s| ; detector u axis: (cos, sin, 0), v axis: z
    eu=np.array([np.cos(phi),np.sin(phi),0.0]); n=np.array([-np.sin(phi),np.cos(phi),0.0])
    X=np.stack([x,y,z],-1)-s
    depth=X@n
    M=SDD/depth
    u=(X@eu)*M/pu+U/2-0.5
    v=X[...,2]*M/pv+V/2-0.5
    return u,v
def wavefronts(iu,iv,pitch):
    # one LDS: 32 lanes, word address = iv*pitch+iu ; wavefronts = max over banks of distinct words
    w=iv*pitch+iu
    b=w%32
    mx=0
    for k in range(32):
        mx=max(mx,len(set(w[b==k])))
    return mx
tot={}
cnt=0
for t in range(3000):
    phi=rng.uniform(0,2*np.pi)
    x0=rng.integers(0,16)*32; y=rng.integers(0,512); z=rng.integers(0,512)
    xs=(np.arange(x0,x0+32)+0.5-N/2)*vox; ys=np.full(32,(y+0.5-N/2)*vox); zs=np.full(32,(z+0.5-N/2)*vox)
    u,v=proj(xs,ys,zs,phi)
    iu=np.floor(u).astype(int); iv=np.floor(v).astype(int)
    s=np.sign((u[-1]-u[0])*(v[-1]-v[0]))
    res={}
    for p in (64,65,68,60,72,76,80,84,92):
        res[p]=wavefronts(iu,iv,p)
    res['adapt']=res[68] if s>=0 else res[60]
    res['adapt2']=res[60] if s>=0 else res[68]
    for k,vv in res.items(): tot[k]=tot.get(k,0)+vv
    cnt+=1
for k,vv in tot.items(): print(k, vv/cnt)
