"""Per-key errors of the golden module cases under both kernel paths (run under gpurun)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
import test_gpu_parity as T
G = T.G
for name in T.RUNNABLE:
    z = G.load(name)
    for prec in ("fp32", "tf32x3"):
        m, out = T.run_case(z, True, prec)
        errs = {"out": G.rel_err(out.detach().cpu().numpy(), z["out"])}
        for tag, mod, want in (("fs", m.fs, z["grad_fs"]), ("rho", m.rho, z["grad_rho"])):
            got = T.grads_of(mod)
            for k in got:
                w = want[k]
                if w is not None and w.size and got[k] is not None and np.linalg.norm(w) > 0:
                    errs[f"{tag}.{k}"] = G.rel_err(got[k], w)
        worst = max(errs, key=errs.get)
        print(f"{name:32s} {prec:7s} worst {worst:14s} {errs[worst]:.2e}  out {errs['out']:.2e}")
