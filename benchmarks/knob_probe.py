import sys, os, torch, statistics
sys.path.insert(0, os.getcwd())
from smow_net_b200 import _lib, ops
dev="cuda:0"; B,C,H=64,32,128
g=torch.Generator(device=dev).manual_seed(0)
x=torch.randn(B,C,2,H,H,device=dev,generator=g).requires_grad_(True)
flow=(torch.randn(B,2,2,H,H,device=dev,generator=g)*0.3).requires_grad_(True)
gout=torch.randn(B,C,4,H,H,device=dev,generator=g)
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
_lib.set_option("warp_fwd_variant",2); _lib.set_option("warp_bwd_variant",2)
out=ops.flow_warp(x,flow,(H,H))
def bwd():
    x.grad=None; flow.grad=None; out.backward(gout,retain_graph=True)
for pf in (0,1,3,6):
    _lib.set_option("cvec_prefetch",pf)
    with torch.no_grad(): tf=t(lambda: ops.flow_warp(x,flow,(H,H)))
    print("PF",pf,"fwd %.3f ms"%tf,"bwd %.3f ms"%t(bwd))
for halo in (1,2,3):
    _lib.set_option("cvec_prefetch",3); _lib.set_option("bwd_halo",halo); _lib.set_option("fwd_halo",halo)
    with torch.no_grad(): tf=t(lambda: ops.flow_warp(x,flow,(H,H)))
    print("HALO",halo,"fwd %.3f ms"%tf,"bwd %.3f ms"%t(bwd))
