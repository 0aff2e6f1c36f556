import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from gnan_b200 import ops
rng = np.random.default_rng(0)
R, G, C = 2708, 1434, 7
dev = 'cuda'
x = torch.tensor(rng.normal(size=(R, G))).float().to(dev)
p = dict(w1=torch.randn(G,64), b1=torch.randn(G,64)*0.3, wh=torch.randn(1,G,64,64)/8, bh=torch.randn(1,G,64)*0.3, wo=torch.randn(G,C,64)/8, bo=torch.randn(G,C)*0.3)
d = {k: v.to(dev).requires_grad_(True) for k, v in p.items()}
dS = torch.randn(R, C, device=dev)
for it in range(3):
    out = ops.mlp(x, d["w1"], d["b1"], d["wh"], d["bh"], d["wo"], d["bo"], 3, precision="tf32x3")
    (out * dS).sum().backward()
torch.cuda.synchronize()
