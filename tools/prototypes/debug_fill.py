import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,"foveated-instance-segmentation_b200"))
import numpy as np, torch
from fovea import ops
from oracle import reference_port as rp
g=dict(np.load(os.path.join(ROOT,'tests/golden/inverse_80_to_128.npz')))
grid,pred=torch.from_numpy(g['grid']),torch.from_numpy(g['pred'])
seg=(128,128); C=5
want=rp.inverse_path(pred,grid,seg,zero_residual=False,tie='max')
plan=ops.build_inverse_plan(grid.cuda(),seg,nchan=C,triangulation='host')
scores,_=ops.inverse_fill(plan,pred.cuda(),zero_residual=False)
s=scores.cpu()
win=plan.winner.cpu()
nanm=torch.isnan(s)!=torch.isnan(want)
print('nan mismatch count',nanm.sum().item(),'of',s.numel())
d=(s-want).abs(); d[torch.isnan(d)]=0
bad=(d>1e-4)|nanm
print('bad count',bad.sum().item(), 'bad at filled px', (bad.any(1)&(win>=0)).sum().item(), 'bad at unfilled', (bad.any(1)&(win<0)).sum().item())
b,y,x=[t[0].item() for t in torch.where(bad.any(1))]
print('first bad',b,y,x,'ours',s[b,:,y,x],'want',want[b,:,y,x])
# emulate
npts=plan.npts.cpu().numpy(); pts=plan.pts.cpu().numpy()[b,:npts[b]]
tris=plan.tris.cpu().numpy()[b]; nbrs=plan.nbrs.cpu().numpy()[b]; nt=plan.ntri.cpu().numpy()[b]
src=plan.src.cpu().numpy()[b]
print('npts',npts[b],'ntri',nt,'hint',plan.hints.cpu().numpy()[b,y//32,x//32])
from scipy.spatial import Delaunay
rc=np.stack([pts>>16,pts&0xFFFF],1).astype(float)
tri=Delaunay(rc)
t=tri.find_simplex(np.array([[y,x]],float))[0]
print('scipy simplex',t,tri.simplices[t],'ours tris row',tris[t],'nbrs',nbrs[t], tri.neighbors[t])
print('verts rc',rc[tri.simplices[t]])
table=ops.box4_table(pred.cuda()).cpu()
T=tri.transform[t]; c=T[:2]@(np.array([y,x])-T[2]); c=np.append(c,1-c.sum())
print('c',c,'srcs',src[tri.simplices[t]])
val=sum(table[b,src[tri.simplices[t][k]],:C]*float(np.float32(c[k])) for k in range(3))
print('recomputed from table',val)
