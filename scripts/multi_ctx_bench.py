"""ChebyshevFiltering through ONE multi-device context (chefsi_create_multi) on the bench workload: host block of
256 columns per device, pinned, X copy-back off -- the call the SPARC shim makes when CHEFSI_B200_DEVICES names
several GPUs.  Prints grid-pt*vectors/s for 1 .. N devices."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sparc_b200 import problem as P
from sparc_b200.chefsi import ChefsiContext

args = argparse.Namespace(grid=160, cell_typ=0, no_nloc=False, ncell=6)
g, veff, proj, (a, b, a0) = bench.build_problem(args)
ngpu = torch.cuda.device_count()
for n in [k for k in (1, 2, 4, 8) if k <= ngpu]:
    ctx = ChefsiContext(list(range(n))) if n > 1 else ChefsiContext(0)
    ctx.set_grid(g); ctx.set_veff(veff); ctx.set_projectors(proj)
    ncol = 256 * n
    xh = torch.empty((ncol, g.Nd), dtype=torch.float64).pin_memory()
    yh = torch.empty((ncol, g.Nd), dtype=torch.float64).pin_memory()
    xh.numpy()[:] = P.random_columns(g.Nd, 1, seed=1)[0]
    ctx.ChebyshevFiltering(xh, yh, 20, a, b, a0, copy_back_x=False)
    best = None
    for _ in range(3):
        t0 = time.perf_counter(); ctx.ChebyshevFiltering(xh, yh, 20, a, b, a0, copy_back_x=False); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    t0 = time.perf_counter(); ctx.set_veff(veff); t_v = time.perf_counter() - t0
    info = ctx.multi_info()
    print(f"{n} device(s): {ncol} columns in {best:.4f} s = {g.Nd * ncol / best:.3e} grid-pt*vectors/s end to end; "
          f"Veff upload + broadcast {1e3 * t_v:.2f} ms; {info}", flush=True)
    ctx.close()
    del xh, yh
