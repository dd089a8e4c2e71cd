#!/usr/bin/env python
"""CPU emulation of the staged walk's tile sizes: bins a C3-sized uniform flock exactly as fit_grid /
grid_keys do (x-slowest keys, z slices), forms each 128-boid CTA's nine interval unions and reports
the CTAs whose total exceeds the 1904-entry tile (they take the slow global-memory path).
    python tools/tile_emulation.py [skin]           # grid anchored at the minimum corner
    N=4194304 EXTENT=1296 python tools/tile_emulation.py 0.33   # another flock (C5)
    CENTER=1 python tools/tile_emulation.py [skin]  # grid centred on the flock (FP_GRID_CENTER=1)
Positions advance ballistically (p0 + t v0): enough to show the structural cause (sliver rows)."""
import os, sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from feriphys_b200 import synth
f32=np.float32
n=int(os.environ.get('N', 1<<20)); extent=float(os.environ.get('EXTENT', 816.0))
st=synth.uniform_flock(n, extent, seed=int(os.environ.get('SEED', 11)))
p0=st[:,:3].astype(np.float64); v0=st[:,3:].astype(np.float64)
reach=16.0; skin=float(sys.argv[1]) if len(sys.argv)>1 else 0.2763
cell=reach*(1+1/512)+skin; zspan=4
lo=p0.min(axis=0); hi=p0.max(axis=0)
dims=[int(np.floor((hi[a]-lo[a])/cell))+1 for a in range(3)]
dimz=int(np.floor((hi[2]-lo[2])/(cell/zspan)))+1
import os
if os.environ.get('CENTER'):
    ext=hi-lo
    lo=lo.copy()
    lo[0]-= (dims[0]*cell-ext[0])/2; lo[1]-=(dims[1]*cell-ext[1])/2; lo[2]-=(dimz*cell/zspan-ext[2])/2
print('dims',dims,dimz,'cell',cell,'origin',lo)
def tiles(p):
    cx=np.clip(np.floor((p[:,0]-lo[0])/cell).astype(np.int64),0,dims[0]-1)
    cy=np.clip(np.floor((p[:,1]-lo[1])/cell).astype(np.int64),0,dims[1]-1)
    cz=np.clip(np.floor((p[:,2]-lo[2])/(cell/zspan)).astype(np.int64),0,dimz-1)
    key=(cx*dims[1]+cy)*dimz+cz
    order=np.argsort(key,kind='stable')
    key=key[order]; cx=cx[order]; cy=cy[order]; cz=cz[order]
    ncells=dims[0]*dims[1]*dimz
    cell_start=np.searchsorted(key,np.arange(ncells+1))
    B=128; nct=n//B
    z0=np.maximum(cz-zspan,0); z1=np.minimum(cz+zspan,dimz-1)
    total=np.zeros(nct,np.int64)
    worst=None
    for r in range(9):
        x=cx+r//3-1; y=cy+r%3-1
        ok=(x>=0)&(x<dims[0])&(y>=0)&(y<dims[1])
        rb=(np.where(ok,x,0)*dims[1]+np.where(ok,y,0))*dimz
        jb=np.where(ok,cell_start[rb+z0],0); je=np.where(ok,cell_start[rb+z1+1],0)
        has=je>jb
        lo_=np.where(has,jb,np.iinfo(np.int64).max).reshape(nct,B).min(axis=1)
        hi_=np.where(has,je,0).reshape(nct,B).max(axis=1)
        ln=np.where(hi_>0,(hi_+3)//4*4-(lo_//4*4),0)
        total+=ln
    return total,(cx,cy,cz)
for k in (0,78,300):
    p=p0+k*1e-3*v0
    t,(cx,cy,cz)=tiles(p)
    over=np.nonzero(t>1904)[0]
    print('step',k,'mean',t.mean(),'max',t.max(),'over',len(over))
    if len(over) and k in (78,300):
        for c in over[:6]:
            sl=slice(c*128,c*128+128)
            print('  cta',c,'total',t[c],'cx',np.unique(cx[sl]),'cy',np.unique(cy[sl]),'cz range',cz[sl].min(),cz[sl].max())
