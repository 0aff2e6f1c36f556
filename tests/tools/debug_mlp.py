import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from tests.test_gpu_parity import rand_mlp, oracle_params
from tests import _golden as G
from oracle import gnan_lut
from gnan_b200 import ops
DEV='cuda'
def run(R,G_,H,C,L, seed=None):
    rng = np.random.default_rng(R * 7 + G_ if seed is None else seed)
    p = rand_mlp(rng, G_, H, C, L)
    u = torch.tensor(rng.normal(size=(R, G_)) * (rng.random((R, G_)) < 0.7)).float()
    dS = torch.tensor(rng.normal(size=(R, C))).float()
    q = oracle_params(p, L)
    want = gnan_lut.feature_sums(q, u.double())
    (want * dS.double()).sum().backward()
    d = {k: v.to(DEV).requires_grad_(v.numel() > 0) for k, v in p.items()}
    got = ops.mlp(u.to(DEV), d["w1"], d["b1"], d["wh"], d["bh"], d["wo"], d["bo"], L)
    (got * dS.to(DEV)).sum().backward()
    errs = {"out": G.rel_err(got.detach().cpu().numpy(), want.detach().numpy())}
    for k in p:
        if p[k].numel() and q[k] is not None:
            errs[k] = G.rel_err(d[k].grad.cpu().numpy(), q[k].grad.numpy())
    print((R,G_,H,C,L), {k: f"{v:.1e}" for k,v in errs.items()})
    return d, q
for cfg in [(1000,40,64,3,3),(1000,40,64,7,3),(1000,39,64,3,3),(1000,41,64,3,3),(896,40,64,3,3),(1000,40,32,3,3),(1000,8,64,3,3),(1000,16,64,3,3), (1000,24,64,3,3),(1000,32,64,3,3),(300,4,64,3,3),(2000,40,64,3,3)]:
    d,q = run(*cfg)
d,q = run(1000,40,64,3,3)
for k in ("w1","wh","wo","bo","b1","bh"):
    a = d[k].grad.cpu().double(); b = q[k].grad
    e = (a-b).abs()
    dims = [i for i in range(a.dim())]
    gdim = {"w1":0,"b1":0,"wh":1,"bh":1,"wo":0,"bo":0}[k]
    per_g = e.transpose(0,gdim).reshape(a.shape[gdim],-1).max(1).values
    print(k, "bad groups:", (per_g > 1e-3).nonzero().flatten().tolist()[:50])
