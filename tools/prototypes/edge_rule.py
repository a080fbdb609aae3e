import sys, os
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from oracle import reference_port as rp
from scipy.spatial import Delaunay
def run(H,W,seed):
    xs,_=rp.synthetic_saliency(1,seed=seed)
    filt,P=rp.gaussian_filter_weight(45,45,45),rp.p_basis(80,80,45,45)
    grid,_=rp.create_grid(rp.pad_saliency(xs,45,45),filt,P,80,80,(80,80))
    pred=rp.synthetic_pred(1,1,seed=seed)
    ps=rp.inverse_sample(pred,rp.grid_inverse(grid,(H,W)))
    mask,inv=rp.pixels_for_interp(ps[0])
    rr,cc=torch.where(mask[0]); pts=np.stack([rr.numpy(),cc.numpy()],1)
    tri=Delaunay(pts.astype(float))
    yy,xx=np.meshgrid(np.arange(H),np.arange(W),indexing='ij')
    q=np.stack([yy,xx],-1).reshape(-1,2)
    sidx=tri.find_simplex(q.astype(float))
    S=tri.simplices; P_=pts.astype(np.int64)
    def orient(a,b,qq): return (b[...,1]-a[...,1])*(qq[...,0]-a[...,0])-(b[...,0]-a[...,0])*(qq[...,1]-a[...,1])
    v=P_[S[sidx]]  # n,3,2
    A=orient(v[:,0],v[:,1],v[:,2]); s=np.sign(A)
    e=[s*orient(v[:,(k+1)%3],v[:,(k+2)%3],q) for k in range(3)]
    e=np.stack(e,1)
    assert (e>=0).all()
    onedge=(e==0).any(1)
    print(H,W,'on-edge pixels',onedge.sum(),'of',len(q), 'on vertex', ((e==0).sum(1)>=2).sum())
    # top-left style rule: e==0 counts as inside iff perturbed sign positive: delta=(dr=eps^2, dc=-eps)
    def accept(vv,ss,qq):
        ok=np.ones(len(qq),bool)
        for k in range(3):
            a=vv[:,(k+1)%3]; b=vv[:,(k+2)%3]
            ek=ss*orient(a,b,qq)
            dr_=b[:,0]-a[:,0]; dc_=b[:,1]-a[:,1]
            tie=np.where(dr_!=0, ss*dr_>0, ss*dc_>0)
            ok&=(ek>0)|((ek==0)&tie)
        return ok
    acc=accept(v,s,q)
    print(' scipy choice accepted by left-rule on on-edge pixels: %d / %d'%(acc[onedge].sum(), onedge.sum()))
    # opposite rule
    def accept2(vv,ss,qq):
        ok=np.ones(len(qq),bool)
        for k in range(3):
            a=vv[:,(k+1)%3]; b=vv[:,(k+2)%3]
            ek=ss*orient(a,b,qq)
            dr_=b[:,0]-a[:,0]; dc_=b[:,1]-a[:,1]
            tie=np.where(dr_!=0, ss*dr_<0, ss*dc_<0)
            ok&=(ek>0)|((ek==0)&tie)
        return ok
    acc2=accept2(v,s,q)
    print(' accepted by right-rule: %d / %d'%(acc2[onedge].sum(), onedge.sum()))
    return
run(128,128,1); run(256,256,2); run(1024,1024,3)
