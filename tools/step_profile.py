"""Host-side profile of the train step (cProfile) + GPU-idle estimate."""
import cProfile, pstats, io, os, sys, types, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from wcmc_b200 import dropin, lib
from wcmc_b200.synth import make_batch
dropin.install(); lib.init()
from sbmc import KPCN
from support.interfaces import KPCNInterface
from support.losses import FeatureMSE, RelativeMSE
from support.networks import PathNet
torch.manual_seed(0)
models={"dncnn":KPCN(39).cuda(),"backbone_diffuse":PathNet(36,outc=3).cuda(),"backbone_specular":PathNet(36,outc=3).cuda()}
optims={"optim_"+k:torch.optim.Adam(m.parameters(),lr=1e-4) for k,m in models.items()}
lf={"l_diffuse":torch.nn.L1Loss(),"l_specular":torch.nn.L1Loss(),"l_recon":torch.nn.L1Loss(),"l_test":RelativeMSE(),"l_manif":FeatureMSE(non_local=True,rng="device")}
itf=KPCNInterface(models,optims,lf,types.SimpleNamespace(model_name="p"),use_llpm_buf=True,manif_learn=True,w_manif=0.1)
batch={k:v.cuda() for k,v in make_batch(batch=8,spp=8,size=128,seed=1).items()}
itf.to_train_mode()
def step():
    itf.preprocess(batch); itf.train_batch(batch)
for _ in range(3): step()
torch.cuda.synchronize()
t=time.time()
for _ in range(5): step()
torch.cuda.synchronize(); print("ms/step", (time.time()-t)/5*1e3)
# host-only time: no sync inside except the finite check
pr=cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize()
pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
