"""Surface fixtures from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY: total order 1 (dense and multigrid)
and n_ind_dim > 1 (dense; the reference's own MultigridLayer fails with n_ind_dim > 1: "size of tensor a (2) must
match the size of tensor b (4)", so that combination is not a surface to reproduce).

    cd /tmp/scratch && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/oracle/make_golden_surface.py
"""
import sys
sys.argv=[sys.argv[0], "/root/repo/tests/golden"]
sys.path.insert(0,"/root/repo")
import numpy as np, torch, os
import oracle.make_golden as MG
from oracle.cases import IV_LISTS
from solver.pde_layer_dense import PDEDenseLayer
from solver.multigrid import MultigridLayer
MG.OUT="/root/repo/tests/golden"
def make_inputs(dims,bs,n_init,seed,M):
    d=len(dims); G=int(np.prod(dims)); g=torch.Generator().manual_seed(seed)
    coeffs=0.3*torch.randn(bs,G,M,generator=g,dtype=torch.float64); coeffs[...,1]+=1.0
    rhs=0.1*torch.randn(bs,G,generator=g,dtype=torch.float64)
    iv=0.5*torch.randn(bs,n_init,generator=g,dtype=torch.float64)
    base=[0.1,0.3,0.25][:d]
    steps=[h*(0.75+0.5*torch.rand(bs,n-1,generator=g,dtype=torch.float64)) for n,h in zip(dims,base)]
    lw=torch.randn(bs,1,G,M,generator=g,dtype=torch.float64)
    return dict(coeffs=coeffs.numpy(),rhs=rhs.numpy(),iv_rhs=iv.numpy(),steps=[s.numpy() for s in steps],loss_w=lw.numpy())
def run(name,kind,dims,iv_name,bs,seed,order,n_ind=1,n_grid=2):
    iv=IV_LISTS[iv_name]
    if kind=="dense":
        layer=PDEDenseLayer(bs=bs,coord_dims=dims,order=order,n_ind_dim=n_ind,n_iv=1,init_index_mi_list=iv,n_iv_steps=1,double_ret=True,solver_dbl=True)
    else:
        layer=MultigridLayer(bs=bs,coord_dims=dims,order=order,n_ind_dim=n_ind,n_iv=1,n_grid=n_grid,evolution=False,downsample_first=True,init_index_mi_list=iv,n_iv_steps=1,double_ret=True,solver_dbl=True)
    M=layer.n_orders
    B=bs*n_ind
    inp=make_inputs(dims,B,layer.pde.num_added_initial_constraints,seed,M)
    # layer inputs: (bs, n_ind, G, M) etc.
    G=int(np.prod(dims))
    torch.set_default_dtype(torch.float64)
    try:
        coeffs=torch.tensor(inp["coeffs"]).reshape(bs,n_ind,G,M).requires_grad_(True)
        rhs=torch.tensor(inp["rhs"]).reshape(bs,n_ind,G).requires_grad_(True)
        ivr=torch.tensor(inp["iv_rhs"]).reshape(bs,n_ind,-1).requires_grad_(True)
        steps=[torch.tensor(s).reshape(bs,n_ind,-1).requires_grad_(True) for s in inp["steps"]]
        infos=[]
        import solver.fgmres as FG
        orig=FG.fgmres_matvec
        def wrap(*a,**k):
            x,info=orig(*a,**k); infos.append((int(info[0]),float(info[1]))); return x,info
        FG.fgmres_matvec=wrap
        u0,u,_=layer(coeffs,rhs,ivr,list(steps))
        FG.fgmres_matvec=orig
        (u*torch.tensor(inp["loss_w"]).reshape(u.shape)).sum().backward()
    finally:
        torch.set_default_dtype(torch.float32)
    save=dict(kind=kind,dims=np.array(dims),iv_name=iv_name,bs=bs,n_ind=n_ind,order=order,seed=seed,n_grid=n_grid,dsf=True,
              coeffs=inp["coeffs"],rhs=inp["rhs"],iv_rhs=inp["iv_rhs"],loss_w=inp["loss_w"],
              u=u.detach().numpy(),u0=u0.detach().numpy(),d_coeffs=coeffs.grad.numpy(),d_rhs=rhs.grad.numpy(),d_iv_rhs=ivr.grad.numpy())
    for c,s in enumerate(inp["steps"]): save[f"steps{c}"]=s
    for c,s in enumerate(steps): save[f"d_steps{c}"]=s.grad.numpy()
    if infos: save["info"]=np.array(infos)
    np.savez_compressed(os.path.join(MG.OUT,f"surf_{name}.npz"),**save)
    print("surf",name,u.shape,infos,flush=True)
run("dense_2d_8x10_order1","dense",(8,10),"burgers",2,71,1)
run("mg_2d_16x16_order1","mg",(16,16),"burgers",2,72,1)
run("dense_1d_24_nind3","dense",(24,),"kamani",2,73,2,n_ind=3)
