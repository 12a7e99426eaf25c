"""Band-parallel Rayleigh-Ritz products, one process per GPU (torchrun): Hp / Mp column blocks and the rotation with the
other ranks' blocks read in place over CUDA IPC / NVLink peer memory (sparc_b200/csrc/ranks.cu).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/band_parallel_bench.py [n] [ncol_total]

Prints, per rank count, the time of the projection products and of the rotation (max over ranks, device-synchronised
host clock around the calls, which synchronise their stream) and the FP64 rate over all ranks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from sparc_b200 import problem as P
from sparc_b200.band_parallel import BandParallelSubspace, rank_project, rank_rotate, _export
from sparc_b200.chefsi import ChefsiContext, _addr
from sparc_b200.partition import band_partition

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl" if world > 1 else "gloo", rank=rank, world_size=world,
                        init_method=None if "MASTER_ADDR" in os.environ else "tcp://127.0.0.1:29512")
n, ncol = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (96, 512)
L = 45.9 * n / 160.0
g = P.make_grid((n, n, n), (L, L, L))
ctx = ChefsiContext(local)
ctx.set_grid(g); ctx.set_veff(P.synthetic_veff(g)); ctx.set_projectors(None)
first, nc = band_partition(ncol, world, rank)
y = P.random_columns(g.Nd, nc, first_col=first, seed=2)
ctx._check(ctx._lib.chefsi_rank_load(ctx._h, _addr(y), y.shape[1], nc, 0))
bp = BandParallelSubspace(ctx)
info = bp._gather_objects((nc, _export(ctx, 0)))
ncols = [v[0] for v in info]
peerY = [0 if r == rank else bp.peers.address(info[r][1]) for r in range(world)]
Q = np.ascontiguousarray(np.random.default_rng(0).standard_normal((nc, ncol)))
X = torch.empty((nc, g.Nd), dtype=torch.float64).pin_memory()
t = torch.zeros(3, dtype=torch.float64, device="cuda")
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); Hp, Mp = rank_project(ctx, False, rank, ncols, peerY); t1 = time.perf_counter()
    dist.barrier()
    t2 = time.perf_counter(); rank_rotate(ctx, False, rank, ncols, peerY, None, Q, X.numpy()); t3 = time.perf_counter()
    dist.barrier()
    t4 = time.perf_counter(); rank_project(ctx, False, rank, ncols, peerY, share=True); t5 = time.perf_counter()
    t[0], t[1], t[2] = t1 - t0, t3 - t2, t5 - t4
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    fl = 2.0 * g.Nd * ncol * ncol
    chk = float(np.abs(Mp[0, first:first + 1] - y[0] @ y[0]).max() / abs(y[0] @ y[0]))
    print(f"{world} ranks, {n}^3 x {ncol}: project (H Y + 2 products + D2H of the blocks) {1e3*t[0].item():.1f} ms = "
          f"{2*fl/t[0].item()/1e12:.1f} TFLOP/s over all ranks; rotate (product + D2H of X) {1e3*t[1].item():.1f} ms = "
          f"{fl/t[1].item()/1e12:.1f} TFLOP/s; Mp check {chk:.1e}; project with every Hermitian block pair formed once "
          f"{1e3*t[2].item():.1f} ms", flush=True)
dist.barrier()
bp.close(); ctx.close()
dist.destroy_process_group()
