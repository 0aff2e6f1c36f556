import sys, torch
sys.path.insert(0, '/root/repo')
from gnan_b200.preprocess import apsp
def graph(n, e, seed=0):
    g = torch.Generator().manual_seed(seed)
    s = torch.randint(0, n, (e,), generator=g); d = torch.randint(0, n, (e,), generator=g)
    keep = s != d; s, d = s[keep], d[keep]
    key = torch.unique(torch.cat([s * n + d, d * n + s]))
    return torch.stack([key // n, key % n])
n, e = (169343, 1166243) if "arxiv" in sys.argv else (19717, 44324)
hd = apsp(graph(n, e).cuda(), n, method="msbfs")
torch.cuda.synchronize()
print(hd.nbins)
