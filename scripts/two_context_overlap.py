"""Experiment: do the projector pass of one column group and the stencil launch of another overlap when two contexts
(two streams) filter different column groups of the bench workload on ONE device at the same time?  Prints the
aggregate rate of 1 thread x 512 columns and of 2 threads x 256 columns (128-column launch groups, device-resident)."""
import argparse, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sparc_b200.chefsi import ChefsiContext

args = argparse.Namespace(grid=160, cell_typ=0, no_nloc=False, ncell=6)
g, veff, proj, (a, b, a0) = bench.build_problem(args)
m, blk = 20, 128


def make(nthreads):
    out = []
    for _ in range(nthreads):
        ctx = ChefsiContext(0)
        ctx.set_grid(g); ctx.set_veff(veff); ctx.set_projectors(proj)
        ld = ctx.device_ld
        bufs = [torch.empty(blk * ld, dtype=torch.float64, device="cuda") for _ in range(3)]
        ctx.fill_random_device(bufs[0], blk, first_col=0, seed=1)
        out.append((ctx, bufs))
    return out


def work(ctx, bufs, ngroups):
    for _ in range(ngroups):
        ctx.filter_device(bufs[0], bufs[1], bufs[2], blk, m, a, b, a0)
    ctx.synchronize()


for nthreads, ngroups in ((1, 4), (2, 2), (1, 4), (2, 2)):
    ws = make(nthreads)
    for ctx, bufs in ws:
        work(ctx, bufs, 1)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(ctx, bufs, ngroups)) for ctx, bufs in ws]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{nthreads} context(s) x {ngroups} groups of {blk} columns: {g.Nd * blk * ngroups * nthreads / dt:.3e} grid-pt*vectors/s "
          f"(GRIDSYNC={os.environ.get('CHEFSI_B200_GRIDSYNC', '1')})", flush=True)
    for ctx, bufs in ws:
        ctx.close()
    del ws
    torch.cuda.empty_cache()
