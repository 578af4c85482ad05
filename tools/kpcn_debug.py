import sys, os, torch
import torch.nn.functional as F
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from wcmc_b200 import dropin, lib, ops
dropin.install(); lib.init()
from sbmc import KPCN
from tests._oracle_loader import load_oracle
from wcmc_b200.synth import make_batch
o=load_oracle()
def rel(a,b): return ((a.double()-b.double()).norm()/(b.double().norm()+1e-30)).item()
size,batch,n_in=128,4,34
torch.manual_seed(0)
ref=o.KPCN(n_in).cuda(); ours=KPCN(n_in).cuda(); ours.load_state_dict(ref.state_dict())
data={k:v.cuda() for k,v in make_batch(batch=batch,size=size,seed=3,paths=False).items()}
# oracle dz capture
ref_dz={}
convs=[m for m in ref.diffuse.modules() if isinstance(m,torch.nn.Conv2d)]
def mk(i):
    def fh(mod, inp, out):
        def gh(g):
            ref_dz[i] = g.detach().clone()
        if out.requires_grad:
            out.register_hook(gh)
    return fh
for i,m in enumerate(convs):
    m.register_forward_hook(mk(i))
out_r=ref(data)
ops.DEBUG_TAP=[]
out_o=ours(data)
tgt=torch.rand_like(out_r["radiance"])
wgt=torch.randn_like(tgt)
mode=os.environ.get("LOSS","smooth")
if mode=="smooth":
    (out_o["diffuse"]*wgt).mean().backward(); (out_r["diffuse"]*wgt).mean().backward()
else:
    F.l1_loss(out_o["diffuse"],tgt).backward(); F.l1_loss(out_r["diffuse"],tgt).backward()
for (k,p),(k2,p2) in zip(ours.diffuse.named_parameters(), ref.diffuse.named_parameters()):
    if "weight" in k: print(k, "grad err %.3e"%rel(p.grad,p2.grad))
for (i,t,coff,c,inv) in ops.DEBUG_TAP:
    mine=(t[...,coff:coff+c].float()*inv).permute(0,3,1,2)
    r=ref_dz[i]
    e=rel(mine,r)
    d=(mine.double()-r.double())
    H=r.shape[2]
    # spatial pattern of squared error
    se=(d**2).sum((0,1)); sr=(r.double()**2).sum((0,1))
    rows=(se.sum(1)/sr.sum(1).clamp_min(1e-300)).sqrt()
    cols=(se.sum(0)/sr.sum(0).clamp_min(1e-300)).sqrt()
    print("layer %d dz err %.3e  | rows[0:6] %s rows[mid] %s rows[-6:] %s | cols[0:6] %s cols[-6:] %s" % (i,e,
        [round(x,3) for x in rows[:6].tolist()],[round(x,3) for x in rows[H//2:H//2+3].tolist()],[round(x,3) for x in rows[-6:].tolist()],
        [round(x,3) for x in cols[:6].tolist()],[round(x,3) for x in cols[-6:].tolist()]))
# ---- in-situ check of the prediction layer's dgrad ----
tap={i:(t,coff,c,inv) for (i,t,coff,c,inv) in ops.DEBUG_TAP}
t8,_,c8,inv=tap[8]; t7,_,c7,_=tap[7]
Wp=ref.diffuse.prediction.weight.detach()
d8=t8[...,:441].permute(0,3,1,2).double()              # our scaled fp16 d_logits
dz7_from_ours=F.conv_transpose2d(d8, Wp.half().double())  # exact dgrad of OUR d_logits with fp16 weights
a7=None
acts=[]
x=data["kpcn_diffuse_in"]
with torch.no_grad():
    h=x
    for m in ref.diffuse.children():
        h=m(h); acts.append(h)
mask=(acts[7]>0)
mine7=t7[...,:100].permute(0,3,1,2).double()
print("in-situ: ours vs exact-dgrad-of-our-dlogits (masked): %.3e" % rel(mine7, dz7_from_ours*mask))
print("exact-dgrad-of-our-dlogits vs oracle dz7: %.3e" % rel(dz7_from_ours*mask*inv.double(), ref_dz[7]))
print("norms: |d8| %.3e  |dz7| %.3e  ratio %.3e ; random-W expectation ratio ~ %.3e" % (d8.norm().item(), (dz7_from_ours*mask).norm().item(), (dz7_from_ours*mask).norm().item()/d8.norm().item(), (Wp.double().pow(2).sum()/441*0.5).sqrt().item()))
