import os, sys, types, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from wcmc_b200 import dropin, lib
from wcmc_b200.synth import make_batch
dropin.install(); lib.init()
from sbmc import KPCN
from support.interfaces import KPCNInterface
from support.losses import FeatureMSE, RelativeMSE
from support.networks import PathNet
from wcmc_b200.engine import GraphedTrainStep, DevicePrefetcher
torch.manual_seed(0)
models={"dncnn":KPCN(39).cuda(),"backbone_diffuse":PathNet(36,outc=3).cuda(),"backbone_specular":PathNet(36,outc=3).cuda()}
optims={"optim_"+k:torch.optim.Adam(m.parameters(),lr=1e-4) for k,m in models.items()}
lf={"l_diffuse":torch.nn.L1Loss(),"l_specular":torch.nn.L1Loss(),"l_recon":torch.nn.L1Loss(),"l_test":RelativeMSE(),"l_manif":FeatureMSE(non_local=True,rng="device")}
itf=KPCNInterface(models,optims,lf,types.SimpleNamespace(model_name="p"),use_llpm_buf=True,manif_learn=True,w_manif=0.1)
host={k:v.pin_memory() for k,v in make_batch(batch=8,spp=8,size=128,seed=1).items()}
dev={k:v.cuda() for k,v in host.items()}
itf.to_train_mode()
g=GraphedTrainStep(itf, dev)
def timeit(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t=time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time()-t)/n*1e3
print("graph replay only        %.2f ms" % timeit(lambda: g.graph.replay()))
print("graphed step (dev batch)  %.2f ms" % timeit(lambda: g(dev)))
def h2d():
    for k in host: dev[k].copy_(host[k], non_blocking=True)
print("H2D 197MB alone           %.2f ms" % timeit(h2d))
print("graphed step (host batch) %.2f ms" % timeit(lambda: g(host)))
def hb():
    while True: yield host
pf=DevicePrefetcher(hb())
def e2e():
    b=next(pf); g(b); pf.release()
print("prefetched e2e (no read)  %.2f ms" % timeit(e2e))
def e2e2():
    b=next(pf); g(b); pf.release(); return float(itf.m_losses["m_l_total"])
print("prefetched e2e (+read)    %.2f ms" % timeit(e2e2))
# pieces of the eager tail
def tail():
    itf._logging(g.loss); itf._optimization()
print("logging+optim eager       %.2f ms" % timeit(tail))
