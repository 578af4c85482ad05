import sys, os, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wcmc_b200 import lib
def rel(a,b): return ((a.double()-b.double()).norm()/(b.double().norm()+1e-30)).item()
g = torch.Generator(device="cuda").manual_seed(1)
dt=torch.float16
k=5
for (n,h,w,cin,cout) in [(4,96,96,100,441),(1,40,40,100,441),(4,96,96,100,256),(4,96,96,100,192)]:
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g)*0.03).to(dt).float()
    dy = torch.randn(n, cout, h-4, w-4, device="cuda", generator=g).to(dt).float()
    ref = F.conv_transpose2d(dy.double(), wt.double())
    wf, wd = lib.pack_weights(wt, dtype=dt)
    dyn = lib.nchw_to_nhwc(dy, dtype=dt)
    for flags,name in [(0,'auto'),(1<<4,'mt1'),(2<<4,'mt2')]:
        dx = lib.conv2d(dyn, wd, None, k, 4, act=0, out_dtype=torch.float32, flags=flags)
        got = dx[..., :cin].permute(0,3,1,2).float()
        e = rel(got, ref)
        msg = "n%d hw%d cout%d %s err %.3e" % (n,h,cout,name,e)
        if e > 1e-3:
            d=(got.double()-ref)
            se=(d**2).sum((0,1)); sr=(ref**2).sum((0,1))
            msg += " rowerr " + str([round(x,2) for x in (se.sum(1)/sr.sum(1)).sqrt()[:34].tolist()])
            msg += " colerr " + str([round(x,2) for x in (se.sum(0)/sr.sum(0)).sqrt()[:34].tolist()])
            msg += " cherr " + str([round(rel(got[:,c],ref[:,c]),2) for c in range(0,cin,9)])
        print(msg)
