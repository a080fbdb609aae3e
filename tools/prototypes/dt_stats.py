import sys, os, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,"foveated-instance-segmentation_b200"))
import numpy as np, torch
from fovea import ops, _lib
from fovea.ops import _ptr,_stream
from oracle import reference_port as rp
H=W=int(sys.argv[1]) if len(sys.argv)>1 else 1024
B=int(sys.argv[2]) if len(sys.argv)>2 else 64
xs,_=rp.synthetic_saliency(B,seed=0)
R=45
g1x,g1y=(t.cuda() for t in ops.separable_factors(rp.gaussian_filter_weight(R,R,R)))
grid=ops.saliency_to_grid(xs.cuda(),g1x,g1y,80,80,R,R,"replication",(80,80))
win=ops.grid_inv_scatter(grid,(H,W))
h=w=80; cap=h*w+4; tcap=2*cap
pts=torch.empty(B,cap,device='cuda',dtype=torch.int32); src=torch.empty_like(pts); npts=torch.empty(B,device='cuda',dtype=torch.int32)
_lib.call("fovea_select_points",_ptr(grid),_ptr(win),B,h,w,H,W,51,cap,_ptr(pts),_ptr(src),_ptr(npts),_stream())
for rep in range(3):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); mesh,ntri,ws=ops.delaunay_device(pts,npts,cap,tcap,max(H,W)); e1.record(); torch.cuda.synchronize()
    print('B=%d delaunay %.3f ms'%(B,e0.elapsed_time(e1)))
ws=ws.cpu().numpy(); rounds=ws[:B]; dbg=ws[B:9*B].reshape(B,8)
print('npts min/mean/max',npts.min().item(),npts.float().mean().item(),npts.max().item())
print('flip rounds min/mean/max',rounds.min(),rounds.mean(),rounds.max())
print('pocket rounds L/R max',dbg[:,4].max(),dbg[:,5].max(),' rows max',dbg[:,2].max())
i=int(rounds.argmax()); print('worst image',i,'dbg',dbg[i])
print('worst: P0+P1 %d wait1 %d P2 %d wait2 %d per round; round0 %d round1 %d round4 %d cycles; rounds %d'%(dbg[i][4]*16/rounds[i],dbg[i][5]*16/rounds[i],dbg[i][6]*16/rounds[i],dbg[i][7]*16/rounds[i],dbg[i][1],dbg[i][2],dbg[i][3],rounds[i]))

# time single images: best and worst
for idx in (int(rounds.argmin()), i):
    p1=pts[idx:idx+1].contiguous(); n1=npts[idx:idx+1].contiguous()
    for rep in range(2):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); ops.delaunay_device(p1,n1,cap,tcap,max(H,W)); e1.record(); torch.cuda.synchronize()
    print('image',idx,'rounds',rounds[idx],'alone %.3f ms'%e0.elapsed_time(e1))
