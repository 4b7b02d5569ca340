import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, golden_helpers as H
from oracle import fdi_oracle as fo
for nodes in [(129,20,66),(64,64,40),(70,37,45),(66,70,130)]:
    rng=np.random.default_rng(sum(nodes)); grid=fo.Grid(np.array(nodes),np.zeros(3),np.ones(3))
    w=dict(dxx=1.3,dyy=0.8,dzz=1.1,dxy=2.0,dyz=0.5,dxz=1.5,dx=0.7,dy=1.9,dz=1.2)
    e=H.engine_for(grid,w); p=rng.standard_normal(grid.n_nodes)
    ref=fo.regularisation_normal_matvec(grid,w,p); q=e.apply_normal(p,parts=1)
    err=np.abs(q-ref).reshape(nodes[::-1]); bad=np.argwhere(err>1e-9*np.abs(ref).max())
    print(nodes,'max err',err.max(),'nbad',len(bad), 'k range',(bad[:,0].min(),bad[:,0].max()) if len(bad) else None,'j',(bad[:,1].min(),bad[:,1].max()) if len(bad) else None,'i',(bad[:,2].min(),bad[:,2].max()) if len(bad) else None)
    e.close()
