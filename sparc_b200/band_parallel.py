"""Band-parallel Rayleigh-Ritz step with ONE PROCESS PER GPU (host side of sparc_b200/csrc/ranks.cu).

SPARC's band communicator gives rank r the columns [r*NB, min((r+1)*NB, Ns)) (src/parallelization.c:403-428,
``partition.band_partition``).  The filter needs nothing from the other ranks; the projection ``Hp = Y^T H Y``,
``Mp = Y^T Y`` and the rotation ``X = Y Q`` do.  The reference re-distributes the block (BP2DP / pdgemr2d,
src/eigenSolver.c:977-990,1504-1582) and multiplies; here every rank keeps its filtered block resident, exports it with
CUDA IPC, and the GEMM kernels of rank I read the blocks of the other ranks in place -- over NVLink between GPUs --
to form the column block I of Hp / Mp / Y Q.  ``torch.distributed`` is plumbing: the exchange of the 64-byte IPC
handles, the barriers between the steps, the all-gather of the small Ns x nc blocks and the broadcast of (lambda, Q)
from rank 0, which solves the subspace eigenproblem as in the reference (src/eigenSolver.c:1262-1375: rank 0 solves,
MPI_Bcast of the eigenvectors).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .chefsi import ChefsiContext, _addr, _is_complex


class _Peers:
    """Device addresses of the other ranks' resident blocks in this process (opened IPC handles, cached by handle)."""

    def __init__(self, ctx: ChefsiContext):
        self.ctx = ctx
        self.open = {}  # handle bytes -> device address

    def address(self, handle: bytes) -> int:
        p = self.open.get(handle)
        if p is None:
            out = C.c_void_p()
            buf = C.create_string_buffer(handle, 64)
            self.ctx._check(self.ctx._lib.chefsi_ipc_open(self.ctx._h, buf, C.byref(out)))
            p = self.open[handle] = int(out.value)
        return p

    def close(self):
        for p in self.open.values():
            self.ctx._lib.chefsi_ipc_close(self.ctx._h, C.c_void_p(p))
        self.open = {}


def _export(ctx: ChefsiContext, which: int) -> bytes:
    buf = C.create_string_buffer(64)
    ctx._check(ctx._lib.chefsi_ipc_export(ctx._h, which, buf))
    return buf.raw


def _ptr_array(addrs):
    arr = (C.c_void_p * len(addrs))()
    for i, a in enumerate(addrs):
        arr[i] = a
    return arr


def rank_block_part(J: int, rank: int, nranks: int, ncJ: int, ncI: int):
    """Hp and Mp are Hermitian, so of every mirrored element pair only one is formed.  Rank I forms block (J, I) -- rows of
    rank J, its own columns -- for the ranks J that follow it cyclically at a distance below P / 2; for even P the block at
    distance P / 2 is shared by its two ranks: the lower one forms the first half of its columns, the upper one the rows
    that mirror the other half.  Returns the part of block (J, rank) that `rank` forms as local (r0, r1, c0, c1): rows
    [r0, r1) x columns [c0, c1), empty when r1 <= r0 or c1 <= c0 (the rule of ``rank_block_part`` in chefsi_internal.h,
    restated here for the host)."""
    if J == rank:
        return 0, ncJ, 0, ncI
    dist = (J - rank) % nranks
    if 2 * dist < nranks:
        return 0, ncJ, 0, ncI
    if 2 * dist == nranks:
        return (0, ncJ, 0, ncI // 2) if rank < J else (ncJ // 2, ncJ, 0, ncI)
    return 0, 0, 0, 0


def assemble_hermitian(blocks, ncols):
    """Full column-major matrix (as a C-ordered [Ns, Ns] array G with G[col, row]) from the per-rank column blocks of a
    shared projection (``rank_project(..., share=True)``): every element that its column's rank did not form is the
    conjugate of its mirror image, which the other rank formed."""
    P = len(ncols)
    G = np.ascontiguousarray(np.concatenate([np.asarray(b) for b in blocks if b.shape[0] > 0]))
    off = np.concatenate([[0], np.cumsum(ncols)]).astype(int)
    for I in range(P):
        for J in range(P):
            if J == I or ncols[I] == 0 or ncols[J] == 0:
                continue
            r0, r1, c0, c1 = rank_block_part(J, I, P, int(ncols[J]), int(ncols[I]))
            formed = np.zeros((ncols[I], ncols[J]), dtype=bool)      # [column of I, row of J], the layout of G
            formed[c0:c1, r0:r1] = True
            mirror = np.conj(G[off[J]:off[J + 1], off[I]:off[I + 1]]).T   # [column of I, row of J] from rank J's block (I, J)
            blk = G[off[I]:off[I + 1], off[J]:off[J + 1]]
            blk[~formed] = mirror[~formed]
    return G


def rank_project(ctx, is_complex, rank, ncols, peerY, share=False):
    """Column block `rank` of (Hp, Mp): numpy arrays [ncols[rank], Ns] (row n = column n of the block: column-major).
    share: form only this rank's share of the Hermitian element pairs (``rank_block_part``); the rest comes back as zeros
    and is mirrored by ``assemble_hermitian`` after the all-gather."""
    ns, nc = int(sum(ncols)), int(ncols[rank])
    dt = np.complex128 if is_complex else np.float64
    Hp, Mp = np.zeros((nc, ns), dtype=dt), np.zeros((nc, ns), dtype=dt)
    nc_arr = (C.c_int * len(ncols))(*[int(v) for v in ncols])
    fn = ctx._lib.chefsi_rank_project_shared if share else ctx._lib.chefsi_rank_project
    ctx._check(fn(ctx._h, int(is_complex), len(ncols), rank, nc_arr, _ptr_array(peerY), _addr(Hp), _addr(Mp), ns))
    return Hp, Mp


def rank_rotate(ctx, is_complex, rank, ncols, peerY, peerT, Q_blk, X_blk):
    """X_blk[ncols[rank], ld] = the rank's columns of Y Q; Q_blk [ncols[rank], Ns] = the rank's columns of Q."""
    ns = int(sum(ncols))
    assert Q_blk.shape == (int(ncols[rank]), ns) and Q_blk.flags.c_contiguous
    nc_arr = (C.c_int * len(ncols))(*[int(v) for v in ncols])
    ctx._check(ctx._lib.chefsi_rank_rotate(ctx._h, int(is_complex), len(ncols), rank, nc_arr, _ptr_array(peerY),
                                           _ptr_array(peerT) if is_complex else None, _addr(Q_blk), ns, _addr(X_blk),
                                           X_blk.shape[1]))


class BandParallelSubspace:
    """One rank's end of the band-parallel projection / eigensolve / rotation.

    ``ctx`` is the rank's single-device context (grid, Veff, projectors, k-point already set); ``group`` a
    ``torch.distributed`` process group (default: the world).  The rank's block must be resident: filter with
    ``keep_y=True`` or pass ``Y_blk`` to :meth:`rayleigh_ritz`.
    """

    def __init__(self, ctx: ChefsiContext, group=None):
        import torch.distributed as dist
        self.ctx, self.group, self.dist = ctx, group, dist
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.peers = _Peers(ctx)

    def close(self):
        self.peers.close()

    # -- plumbing --------------------------------------------------------------------------------------------------
    def _gather_objects(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def _broadcast_array(self, arr, src=0):
        """In-place broadcast of a numpy array through a CPU tensor (gloo) or a CUDA tensor (nccl)."""
        import torch
        t = torch.from_numpy(arr.view(np.float64) if np.iscomplexobj(arr) else arr)
        if self.dist.get_backend(self.group) == "nccl":
            d = t.cuda()
            self.dist.broadcast(d, src=src, group=self.group)
            t.copy_(d.cpu())
        else:
            self.dist.broadcast(t, src=src, group=self.group)
        return arr

    # -- the step ----------------------------------------------------------------------------------------------------
    def rayleigh_ritz(self, nc, is_complex, Y_blk=None, X_blk=None):
        """Projection, eigensolve and rotation of the block whose `nc` local columns are resident on this rank's device
        (or given as the host block ``Y_blk``).  Returns (lambda[Ns], X_blk[nc, ld])."""
        ctx, rank = self.ctx, self.rank
        if Y_blk is not None:
            is_complex = _is_complex(Y_blk)
            nc = Y_blk.shape[0]
            ctx._check(ctx._lib.chefsi_rank_load(ctx._h, _addr(Y_blk), Y_blk.shape[1], nc, int(is_complex)))
        info = self._gather_objects((int(nc), _export(ctx, 0), _export(ctx, 2) if is_complex else b""))
        ncols = [v[0] for v in info]
        peerY = [0 if r == rank or ncols[r] == 0 else self.peers.address(info[r][1]) for r in range(self.world)]
        peerT = [0 if r == rank or ncols[r] == 0 or not is_complex else self.peers.address(info[r][2]) for r in range(self.world)]
        ns, c0 = int(sum(ncols)), int(sum(ncols[:rank]))
        self.dist.barrier(group=self.group)                       # every rank's Y is resident
        Hp_blk, Mp_blk = rank_project(ctx, is_complex, rank, ncols, peerY, share=True)  # each Hermitian block pair once
        blocks = self._gather_objects((Hp_blk, Mp_blk))           # Ns x nc blocks: small next to the orbitals
        dt = np.complex128 if is_complex else np.float64
        lam, Q = np.zeros(ns), np.zeros((ns, ns), dtype=dt)
        if rank == 0:                                             # eigenSolver.c:1262-1375: rank 0 solves, then MPI_Bcast
            Hp = assemble_hermitian([b[0] for b in blocks], ncols)
            Mp = assemble_hermitian([b[1] for b in blocks], ncols)
            lam, Q = ctx.DP_Solve_Generalized_EigenProblem(ns, Hp, Mp)
        self._broadcast_array(lam)
        self._broadcast_array(Q)
        ctx._check(ctx._lib.chefsi_rank_rotate_prepare(ctx._h, int(is_complex)))
        self.dist.barrier(group=self.group)                       # complex: every rank's T = i Y is in place
        if X_blk is None:
            X_blk = np.empty((nc, ctx.grid.Nd), dtype=dt)
        rank_rotate(ctx, is_complex, rank, ncols, peerY, peerT, np.ascontiguousarray(Q[c0:c0 + nc]), X_blk)
        self.dist.barrier(group=self.group)                       # nobody overwrites its Y before all products are done
        return lam, X_blk
