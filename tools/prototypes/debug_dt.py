import sys, os, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,"foveated-instance-segmentation_b200"))
import numpy as np, torch
from fovea import ops
from oracle import reference_port as rp
H=W=int(sys.argv[1]) if len(sys.argv)>1 else 128
xs,_=rp.synthetic_saliency(2,seed=1)
filt,P=rp.gaussian_filter_weight(45,45,45),rp.p_basis(80,80,45,45)
grid,_=rp.create_grid(rp.pad_saliency(xs,45,45),filt,P,80,80,(80,80))
g=grid.contiguous().cuda()
win=ops.grid_inv_scatter(g,(H,W))
B,h,w,_=g.shape; cap=h*w+4; tcap=2*cap
pts=torch.empty(B,cap,device='cuda',dtype=torch.int32); src=torch.empty_like(pts); npts=torch.empty(B,device='cuda',dtype=torch.int32)
from fovea import _lib
from fovea.ops import _ptr,_stream
_lib.call("fovea_select_points",_ptr(g),_ptr(win),B,h,w,H,W,51,cap,_ptr(pts),_ptr(src),_ptr(npts),_stream())
torch.cuda.synchronize(); print('npts',npts.cpu().tolist(),flush=True)
t0=time.time()
mesh,ntri,ws=ops.delaunay_device(pts,npts,cap,tcap,max(H,W))
torch.cuda.synchronize(); print('delaunay done %.3fs'%(time.time()-t0),flush=True)
ws=ws.cpu().numpy(); print('rounds',ws[:B],'dbg',ws[B:].reshape(B,8),'ntri',ntri.cpu().tolist(),flush=True)
sys.path.insert(0,os.path.join(ROOT,'tests'))
from test_delaunay_gpu import check_mesh
p=pts.cpu().numpy(); n=npts.cpu().numpy(); m=mesh.cpu().numpy(); nt=ntri.cpu().numpy()
for b in range(B):
    rc=np.stack([p[b,:n[b]]>>16,p[b,:n[b]]&0xFFFF],1)
    try:
        check_mesh(rc,m[b],int(nt[b])); print('image',b,'mesh OK')
    except AssertionError as e:
        print('image',b,'FAILED',str(e)[:300])
