import sys, copy, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sivae_b200
from sivae_b200 import functional as F, trainer as T
DEV='cuda'
def run(fused):
    torch.manual_seed(11); F.manual_seed(11)
    net = sivae_b200.SoftIntroVAE(64, [[64,1,2],[64,1,2],[64,1,2]]).to(DEV); net.apply(T.init_weights_he); net.train()
    if fused: oe, od = sivae_b200.FusedAdam(net.encoder.parameters(), lr=2e-4), sivae_b200.FusedAdam(net.decoder.parameters(), lr=2e-4)
    else: oe, od = torch.optim.Adam(net.encoder.parameters(), lr=2e-4), torch.optim.Adam(net.decoder.parameters(), lr=2e-4)
    g = torch.Generator(device=DEV).manual_seed(5)
    real = torch.rand(2,1,16,24,16, device=DEV, generator=g); noise = torch.randn(2,1,2,3,2, device=DEV, generator=g)
    torch.cuda.manual_seed(99)
    outs=[]
    snaps=[]
    for s in range(2):
        o = T.soft_intro_train_step(net, real, noise, oe, od)
        outs.append({k: float(v) for k,v in o.items()})
        snaps.append({k: p.detach().clone() for k,p in net.named_parameters()})
    return outs, snaps
a, sa = run(True); b, sb = run(False)
print(a[0]); print(b[0]); print(a[1]); print(b[1])
for s in range(2):
    worst = sorted(((float((sa[s][k]-sb[s][k]).abs().max()), k) for k in sa[s]), reverse=True)[:6]
    print('step', s, worst)
