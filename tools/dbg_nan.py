import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, beamform_b200 as bf
from golden.cases import build_case
G=np.load('/root/repo/tests/golden/ref_outputs.npz')
name='mvdr_aira3_coldstart_nan'
cfg,x,ev=build_case(name)
ref=G[name+'/out']
got=bf.Beamformer(cfg,1).process(x[None])[0]
H=512
for t in range(len(ref)//H):
    r=ref[t*H:(t+1)*H]; g=got[t*H:(t+1)*H]
    print(t, 'ref nonfinite',(~np.isfinite(r)).sum(),'got nonfinite',(~np.isfinite(g)).sum(), 'rel', np.linalg.norm((g-r)[np.isfinite(r)&np.isfinite(g)])/max(1e-30,np.linalg.norm(r[np.isfinite(r)&np.isfinite(g)])))
