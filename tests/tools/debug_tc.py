import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from tests.test_gpu_parity import rand_mlp, oracle_params, relu_margin
from tests import _golden as G
from oracle import gnan_lut
from gnan_b200 import ops
DEV='cuda'
def run(R,G_,C, prec, seed=0):
    H,L=64,3
    rng = np.random.default_rng(seed)
    p = rand_mlp(rng, G_, H, C, L)
    u = torch.tensor(rng.normal(size=(R, G_)) * (rng.random((R, G_)) < 0.7)).float()
    dS = torch.tensor(rng.normal(size=(R, C))).float()
    q = oracle_params(p, L)
    want = gnan_lut.feature_sums(q, u.double())
    (want * dS.double()).sum().backward()
    res = {}
    for pr in ("fp32", prec):
        d = {k: v.to(DEV).requires_grad_(v.numel() > 0) for k, v in p.items()}
        got = ops.mlp(u.to(DEV), d["w1"], d["b1"], d["wh"], d["bh"], d["wo"], d["bo"], L, precision=pr)
        (got * dS.to(DEV)).sum().backward()
        errs = {"out": G.rel_err(got.detach().cpu().numpy(), want.detach().numpy())}
        for k in p: errs[k] = G.rel_err(d[k].grad.cpu().numpy(), q[k].grad.numpy())
        res[pr] = errs
    print((R,G_,C), "margin %.1e" % relu_margin(p,u))
    for pr,e in res.items(): print("   ", pr, {k: f"{v:.1e}" for k,v in e.items()})
for cfg in [(2708,70,7),(1000,300,7),(1000,40,4),(2560,70,7),(2708,70,1),(640,300,3),(128*9,296,2)]:
    run(*cfg, "tf32x3")
run(1000,40,4,"tf32")
