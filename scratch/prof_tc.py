import os, sys, ctypes, torch, numpy as np
os.environ["GNAN_TC_PROF"] = "1"
sys.path.insert(0, '/root/repo')
from gnan_b200 import ops, _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
rng = np.random.default_rng(0)
R, G, C = 2708, 1434, 7
dev = 'cuda'
x = torch.tensor(rng.normal(size=(R, G))).float().to(dev)
p = dict(w1=torch.randn(G,64), b1=torch.randn(G,64)*0.3, wh=torch.randn(1,G,64,64)/8, bh=torch.randn(1,G,64)*0.3, wo=torch.randn(G,C,64)/8, bo=torch.randn(G,C)*0.3)
d = {k: v.to(dev).requires_grad_(True) for k, v in p.items()}
dS = torch.randn(R, C, device=dev)
buf = (ctypes.c_longlong * 16)()
for it in range(3):
    out = ops.mlp(x, d["w1"], d["b1"], d["wh"], d["bh"], d["wo"], d["bo"], 3, precision="tf32x3")
    lib.gnan_debug_tc_prof(buf)
    (out * dS).sum().backward()
    lib.gnan_debug_tc_prof(buf)
    v = list(buf)
    n = max(v[8], 1)
    names = {9: "A.pre+g", 10: "A.compute", 11: "A.tmem_st", 12: "A.wait_st", 0: "A.fence+arrive", 6: "B.dWo(prev)", 1: "wait MMA1", 2: "D.epiC", 3: "wait MMA2", 4: "E.epiF", 7: "loop"}
    print("tiles", v[8], {nm: int(v[i] / n) for i, nm in names.items()}, "sum/tile", int(sum(v[i] for i in names) / n))
