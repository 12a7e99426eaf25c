"""Band-parallel Rayleigh-Ritz products with one rank per GPU (sparc_b200/csrc/ranks.cu, sparc_b200/band_parallel.py):
rank I forms the column block I of Mp = Y^H Y, Hp = Y^H H Y and of X = Y Q while its GEMM kernels read the other ranks'
resident blocks in place (CUDA IPC / NVLink peer memory) -- the exchange the reference does with BP2DP / pdgemr2d before
its dgemm (src/eigenSolver.c:977-990,1504-1582; band split: src/parallelization.c:403-428).

Two set-ups: (a) two contexts of ONE process standing in for two ranks (plain device addresses), which exercises the
rank arithmetic of the C entry points; (b) two real processes (torch.multiprocessing, gloo for the plumbing) that exchange
CUDA IPC handles -- on a one-GPU box both ranks use device 0, with two or more GPUs rank r uses device r and the kernels
read over NVLink.  Checks are against numpy products with the oracle's H apply (1e-10), and the distributed step against
the single-context chain project -> eig -> rotate."""
import os
import sys

import numpy as np
import pytest

from sparc_b200 import problem as P
from tests.cases import KVEC, overlap_case, rel_fro, small_case

pytestmark = pytest.mark.gpu
TOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(ctx, g, veff, proj, kvec=(0, 0, 0)):
    ctx.set_grid(g)
    ctx.set_veff(veff)
    ctx.set_projectors(proj)
    ctx.set_kpoint(kvec)


@pytest.mark.parametrize("complex_", [False, True])
@pytest.mark.parametrize("split", [(5, 3), (1, 6), (40, 30)])
def test_rank_products_two_contexts_one_process(port, complex_, split):
    import ctypes as C
    from sparc_b200.band_parallel import rank_project, rank_rotate
    from sparc_b200.chefsi import ChefsiContext, _addr
    ncol = sum(split)
    if ncol > 16:
        g, veff, proj, y = overlap_case("stream", ncol=ncol, complex_=complex_)
    else:
        g, veff, proj, y = small_case(17, complex_=complex_, ncol=ncol)
    kvec = KVEC if complex_ else (0, 0, 0)
    ctxs = [ChefsiContext(0), ChefsiContext(0)]
    try:
        blocks = [np.ascontiguousarray(y[:split[0]]), np.ascontiguousarray(y[split[0]:])]
        for c, blk in zip(ctxs, blocks):
            _setup(c, g, veff, proj, kvec)
            c._check(c._lib.chefsi_rank_load(c._h, _addr(blk), blk.shape[1], blk.shape[0], int(complex_)))
        addr = lambda which: [int(c._lib.chefsi_resident_ptr(c._h, which) or 0) for c in ctxs]
        Hy = port.hamiltonian_mult(g, proj, veff, 0.0, y, kvec=kvec)
        Mp_want, Hp_want = y.conj() @ y.T, y.conj() @ Hy.T          # element (row r, col c) = <y_r, y_c>
        Q = np.random.default_rng(7).standard_normal((ncol, ncol)) + (1j * np.random.default_rng(8).standard_normal((ncol, ncol)) if complex_ else 0)
        Q = np.ascontiguousarray(Q)                                  # numpy Q[n] = column n of Q
        X_want = Q @ y                                               # X_n = sum_k Q[k, n] y_k
        peerY = addr(0)
        for r in range(2):
            Hp, Mp = rank_project(ctxs[r], complex_, r, split, peerY)
            c0 = 0 if r == 0 else split[0]
            # numpy [n, m] = element (row m, column c0 + n) of the matrix
            assert rel_fro(Mp, Mp_want[:, c0:c0 + split[r]].T) < TOL
            assert rel_fro(Hp, Hp_want[:, c0:c0 + split[r]].T) < TOL
        # every Hermitian block pair formed by one rank only, the rest mirrored on the host after the "all-gather"
        from sparc_b200.band_parallel import assemble_hermitian
        shared = [rank_project(ctxs[r], complex_, r, split, peerY, share=True) for r in range(2)]
        h0 = split[0] // 2   # two ranks share their one off-diagonal block: rank 1 forms rows [h0, nc0) of block (0, 1) only
        assert not shared[1][0][:, :h0].any() and not shared[1][1][:, :h0].any()
        assert not shared[0][0][h0:, split[0]:].any() and not shared[0][1][h0:, split[0]:].any()
        assert rel_fro(assemble_hermitian([sb[1] for sb in shared], split), Mp_want.T) < TOL
        assert rel_fro(assemble_hermitian([sb[0] for sb in shared], split), Hp_want.T) < TOL
        for c in ctxs:
            c._check(c._lib.chefsi_rank_rotate_prepare(c._h, int(complex_)))
        peerT = addr(2)
        for r in range(2):
            c0 = 0 if r == 0 else split[0]
            X = np.empty((split[r], g.Nd), dtype=y.dtype)
            rank_rotate(ctxs[r], complex_, r, split, peerY, peerT, np.ascontiguousarray(Q[c0:c0 + split[r]]), X)
            assert rel_fro(X, X_want[c0:c0 + split[r]]) < TOL
    finally:
        for c in ctxs:
            c.close()


def _worker(rank, world, port_no, out_dir, complex_, ncol):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from sparc_b200.band_parallel import BandParallelSubspace
    from sparc_b200.chefsi import ChefsiContext
    from sparc_b200.partition import band_partition
    from tests.cases import BOUNDS, KVEC, overlap_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank if torch.cuda.device_count() >= world else 0
    g, veff, proj, x = overlap_case("stream", ncol=ncol, complex_=complex_)
    first, n = band_partition(ncol, world, rank)
    ctx = ChefsiContext(dev)
    ctx.set_grid(g); ctx.set_veff(veff); ctx.set_projectors(proj); ctx.set_kpoint(KVEC if complex_ else (0, 0, 0))
    a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
    X = np.ascontiguousarray(x[first:first + n])
    Y = np.empty_like(X)
    ctx.subspace_reserve(n, complex_)
    ctx.ChebyshevFiltering(X, Y, 7, a, b, a0, keep_y=True, copy_back_y=False)   # Y never visits the host
    bp = BandParallelSubspace(ctx)
    lam, Xr = bp.rayleigh_ritz(n, complex_)
    np.save(os.path.join(out_dir, f"x{rank}.npy"), Xr)
    if rank == 0:
        np.save(os.path.join(out_dir, "lam.npy"), lam)
    dist.barrier()
    bp.close()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("complex_", [False, True])
def test_band_parallel_step_two_processes_over_cuda_ipc(tmp_path, complex_):
    """filter (Y kept on each rank's device) -> projection over IPC peer memory -> eigensolve on rank 0 -> rotation over
    IPC peer memory, against the same chain on ONE context holding all columns."""
    import torch.multiprocessing as mp
    from sparc_b200.chefsi import ChefsiContext
    ncol = 11
    port_no = 29700 + (os.getpid() % 200) + (50 if complex_ else 0)
    mp.spawn(_worker, args=(2, port_no, str(tmp_path), complex_, ncol), nprocs=2, join=True)
    g, veff, proj, x = overlap_case("stream", ncol=ncol, complex_=complex_)
    ctx = ChefsiContext(0)
    try:
        _setup(ctx, g, veff, proj, KVEC if complex_ else (0, 0, 0))
        a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
        X, Y = x.copy(), np.empty_like(x)
        ctx.ChebyshevFiltering(X, Y, 7, a, b, a0)
        dt = x.dtype
        Hp, Mp = np.zeros((ncol, ncol), dtype=dt), np.zeros((ncol, ncol), dtype=dt)
        ctx.DP_Project_Hamiltonian(Y, Hp, Mp)
        lam, Q = ctx.DP_Solve_Generalized_EigenProblem(ncol, Hp, Mp)
        Xw = np.empty_like(x)
        ctx.DP_Subspace_Rotation(Q, Xw)
    finally:
        ctx.close()
    got = np.concatenate([np.load(tmp_path / "x0.npy"), np.load(tmp_path / "x1.npy")])
    lam_got = np.load(tmp_path / "lam.npy")
    assert np.abs(lam_got - lam).max() <= 1e-10 * max(1.0, np.abs(lam).max())
    # Ritz vectors: equal up to a sign / phase per vector (the two eigensolves see matrices that differ in the last bits)
    for n in range(ncol):
        ov = np.vdot(Xw[n], got[n]) / (np.linalg.norm(Xw[n]) * np.linalg.norm(got[n]))
        assert abs(abs(ov) - 1.0) < 1e-8, (n, ov)
        assert rel_fro(got[n] * np.conj(ov / abs(ov)), Xw[n]) < 1e-7
