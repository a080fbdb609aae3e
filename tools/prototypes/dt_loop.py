import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,"foveated-instance-segmentation_b200")); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np, torch
from fovea import ops
import test_delaunay_gpu as T
rng=np.random.default_rng(0)
sets=[]
sets.append(T._sorted_unique(rng.integers(0,200,size=(1500,2))))
rr,cc=np.meshgrid(np.arange(40),np.arange(50),indexing='ij'); sets.append(T._sorted_unique(np.stack([rr.ravel(),cc.ravel()],1)))
zz=np.stack([np.arange(0,300,3),100+80*((np.arange(100)%2)*2-1)*(np.arange(100)%7)//7],1)
sets.append(T._sorted_unique(np.concatenate([zz,[[0,0],[0,250],[299,0],[299,250]]])))
sets.append(T._sorted_unique(np.array([[0,0],[0,5],[0,9],[7,2],[7,3],[7,30]])))
sets.append(T._sorted_unique(np.array([[0,0],[0,10],[0,20],[5,7]])))
sets.append(T._sorted_unique(np.concatenate([np.stack([np.arange(50),np.arange(50)],1),[[10,40]]])))
blob=np.clip(rng.normal(2000,15,size=(4000,2)).astype(np.int64),0,4095)
sets.append(T._sorted_unique(np.concatenate([blob,[[0,0],[0,4095],[4095,0],[4095,4095]]])))
print('sizes',[len(s) for s in sets])
for rep in range(int(sys.argv[1]) if len(sys.argv)>1 else 20):
    for i,s in enumerate(sets):
        try:
            mesh,ntri,ws=T._run_device(ops,[s],4096)
            if ws[1]<0: print('rep',rep,'set',i,'DBG',ws[:9]); sys.exit(0)
            T.check_mesh(s,mesh[0],int(ntri[0]))
        except AssertionError as e:
            print('rep',rep,'set',i,'ASSERT',str(e)[:150],'ws',ws[:9]); 
        except Exception as e:
            print('rep',rep,'set',i,'EXC',type(e).__name__,str(e)[:200]); sys.exit(1)
print('done')
