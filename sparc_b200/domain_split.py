"""Optional DOMAIN SPLIT of the CheFSI filter: z slabs of the grid over the ranks of a process group.

Only for grids whose column blocks exceed one GPU's HBM (BASELINE.json north_star); every
listed configuration fits one B200, so the band split (``partition.py``) is the production path and this is
the escape hatch.  It mirrors what the reference does when its domain communicator has more than one rank:

  * before every stencil application the FDn boundary planes of the input are exchanged with the two z
    neighbours (``Lap_plus_diag_vec_mult_orth``: pack 6 faces, ``MPI_Ineighbor_alltoallv``, unpack,
    src/lapVecRoutines.c:387-442,494-534) -- here ``torch.distributed`` send/recv (NCCL over NVLink on the
    box, gloo in the CPU tests) of the two z faces; x and y stay whole on every rank;
  * the projector inner products are partial sums over the local sphere points and are all-reduced over the
    domain communicator before the scaling by Gamma (``Vnl_vec_mult``: ``MPI_Allreduce`` of alpha,
    src/nlocVecRoutines.c:834-838) -- here ``dist.all_reduce`` of the per-atom alpha buffer.

Each rank stores its slab WITH its halo planes (nzl + 2*FDn planes) and sets the local problem with a Dirichlet
z face, so the local kernels never wrap in z; whatever they compute on the halo planes is overwritten by the
next exchange.  The three-term recurrence (src/eigenSolver.c:747-796) runs here, one exchange + one all-reduce
per degree.  The arithmetic stays in the engine (``GpuSlabEngine`` -> libchefsi_b200.so).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .problem import Grid, Projectors

FDN = 6


def z_slabs(Nz: int, nparts: int):
    """Contiguous, balanced plane ranges [(z0, z1), ...]; every slab needs at least FDn planes."""
    base, rem = divmod(Nz, nparts)
    out, z = [], 0
    for r in range(nparts):
        n = base + (1 if r < rem else 0)
        out.append((z, z + n))
        z += n
    if min(b - a for a, b in out) < FDN:
        raise ValueError(f"{Nz} planes over {nparts} ranks leaves a slab thinner than the FD radius {FDN}")
    return out


def slab_grid(grid: Grid, z0: int, z1: int) -> Grid:
    """The rank's local problem: its planes plus FDn halo planes on either side, Dirichlet in z (the halo planes
    carry the neighbours' data, nothing is wrapped locally); stencil tables unchanged."""
    if grid.FDn != FDN:
        raise ValueError("domain split is built for FD radius 6")
    return dataclasses.replace(grid, N=(grid.N[0], grid.N[1], z1 - z0 + 2 * FDN), BC=(grid.BC[0], grid.BC[1], 1))


def _ext_planes(grid: Grid, z0: int, z1: int):
    """Global plane index (or -1 = outside a Dirichlet cell) of every local plane."""
    Nz = grid.N[2]
    idx = np.arange(z0 - FDN, z1 + FDN)
    if grid.BC[2] == 0:
        return idx % Nz
    return np.where((idx >= 0) & (idx < Nz), idx, -1)


def slab_field(grid: Grid, field: np.ndarray, z0: int, z1: int) -> np.ndarray:
    """A grid function (..., Nd) cut to the slab with halo planes (periodic images, or zeros outside a Dirichlet cell)."""
    Nx, Ny, Nz = grid.N
    f = field.reshape(field.shape[:-1] + (Nz, Ny * Nx))
    planes = _ext_planes(grid, z0, z1)
    out = np.zeros(field.shape[:-1] + (planes.size, Ny * Nx), dtype=field.dtype)
    ok = planes >= 0
    out[..., ok, :] = f[..., planes[ok], :]
    return np.ascontiguousarray(out.reshape(field.shape[:-1] + (-1,)))


def slab_projectors(grid: Grid, proj: Projectors | None, z0: int, z1: int) -> Projectors | None:
    """Projector tables of the rank: the sphere points of its OWN planes with local indices; atoms, IP_displ and
    Gamma stay global (alpha is all-reduced over atoms).  Images without local points are dropped."""
    if proj is None:
        return None
    Nx, Ny, _ = grid.N
    plane = Nx * Ny
    img_atom, img_ndc, img_coords, pos_list, chi_list = [], [], [], [], []
    for J in range(proj.n_img):
        pos = proj.grid_pos[proj.pos_off[J]:proj.pos_off[J + 1]]
        ndc = pos.size
        a = int(proj.img_atom[J])
        nproj = int(proj.IP_displ[a + 1] - proj.IP_displ[a])
        chi = proj.chi[proj.chi_off[J]:proj.chi_off[J + 1]].reshape(nproj, ndc)
        k = pos // plane
        keep = (k >= z0) & (k < z1)
        if not keep.any():
            continue
        img_atom.append(a)
        img_ndc.append(int(keep.sum()))
        img_coords.append(proj.img_coords[3 * J:3 * J + 3])
        pos_list.append((pos[keep] - (z0 - FDN) * plane).astype(np.int32))
        chi_list.append(np.ascontiguousarray(chi[:, keep]).reshape(-1))
    n_img = len(img_atom)
    pos_off = np.zeros(n_img + 1, dtype=np.int64)
    chi_off = np.zeros(n_img + 1, dtype=np.int64)
    for q in range(n_img):
        pos_off[q + 1] = pos_off[q] + img_ndc[q]
        chi_off[q + 1] = chi_off[q] + chi_list[q].size
    return Projectors(
        n_atom=proj.n_atom, IP_displ=proj.IP_displ, gamma=proj.gamma,
        img_atom=np.asarray(img_atom, dtype=np.int32), img_ndc=np.asarray(img_ndc, dtype=np.int32),
        img_coords=np.ascontiguousarray(np.concatenate(img_coords)) if n_img else np.zeros(0),
        pos_off=pos_off, chi_off=chi_off,
        grid_pos=np.concatenate(pos_list).astype(np.int32) if n_img else np.zeros(0, dtype=np.int32),
        chi=np.ascontiguousarray(np.concatenate(chi_list)) if n_img else np.zeros(0))


class GpuSlabEngine:
    """The local pieces on one B200 through the C ABI (chefsi_stencil_step_device / chefsi_nloc_*_device)."""

    def __init__(self, device: int, grid_loc: Grid, veff_loc, proj_loc):
        import torch
        from .chefsi import ChefsiContext
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.ctx = ChefsiContext(device)
        self.ctx.set_grid(grid_loc)
        self.ctx.set_veff(veff_loc)
        self.have_proj = proj_loc is not None and proj_loc.n_img > 0
        self.ctx.set_projectors(proj_loc if self.have_proj else None)
        self.ld = self.ctx.device_ld
        self.nd = grid_loc.Nd

    def block(self, ncol):
        return self.torch.zeros((ncol, self.ld), dtype=self.torch.float64, device=self.device)

    def upload(self, x_np):
        b = self.block(x_np.shape[0])
        b[:, :self.nd] = self.torch.from_numpy(x_np).to(self.device)
        return b

    def download(self, b):
        self.ctx.synchronize()
        return b[:, :self.nd].cpu().numpy()

    def alpha_buffer(self, ntot, ncol):
        return self.torch.zeros(ntot * ncol, dtype=self.torch.float64, device=self.device)

    def stencil_step(self, x, xprev, out, c, s1, s2):
        self.ctx.stencil_step_device(x, xprev, out, x.shape[0], c, s1, s2)

    def project(self, x, alpha):
        if self.have_proj:
            self.ctx.nloc_project_device(x, x.shape[0], alpha)
        else:
            alpha.zero_()

    def expand(self, out, scale, alpha):
        if self.have_proj:
            self.ctx.nloc_expand_device(out, out.shape[0], scale, alpha)

    def sync(self):
        """Wait for the library's stream (its kernels) -- before torch / NCCL touch the blocks."""
        self.ctx.synchronize()

    def fence(self):
        """Wait for torch's streams (copies, NCCL) -- before the library's stream touches the blocks: the library
        runs on its own non-blocking stream, which torch's stream semantics know nothing about."""
        self.torch.cuda.synchronize(self.device)

    def close(self):
        self.ctx.close()


class DomainSplitFilter:
    """ChebyshevFiltering over z slabs.  ``engine`` does the local arithmetic on blocks of shape (ncol, >= Nd_local)
    (torch tensors: CUDA for ``GpuSlabEngine``); ``group`` is the torch.distributed group of the split (None with
    ``world == 1``: the exchange degenerates to local copies, which is how the single-process tests run it)."""

    def __init__(self, engine, grid: Grid, slab, rank: int, world: int, n_proj_total: int, group=None):
        self.e = engine
        self.grid = grid
        self.z0, self.z1 = slab
        self.rank, self.world = rank, world
        self.ntot = int(n_proj_total)
        self.group = group
        self.plane = grid.N[0] * grid.N[1]
        self.nzl = self.z1 - self.z0

    # -- halo exchange of one block (planes are contiguous in a column: a face is one strided copy) ------------
    def exchange(self, blk):
        import torch
        import torch.distributed as dist
        ncol = blk.shape[0]
        v = blk[:, :(self.nzl + 2 * FDN) * self.plane].view(ncol, self.nzl + 2 * FDN, self.plane)
        lo_halo, hi_halo = v[:, :FDN], v[:, FDN + self.nzl:]
        lo_own, hi_own = v[:, FDN:2 * FDN], v[:, self.nzl:self.nzl + FDN]
        periodic = self.grid.BC[2] == 0
        prev, nxt = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        has_prev = periodic or self.rank > 0
        has_next = periodic or self.rank < self.world - 1
        self.e.sync()
        if self.world == 1:
            if periodic:
                lo_halo.copy_(hi_own.clone())
                hi_halo.copy_(lo_own.clone())
            else:
                lo_halo.zero_(); hi_halo.zero_()
            self.e.fence()
            return
        send_dn, send_up = lo_own.contiguous(), hi_own.contiguous()
        recv_hi, recv_lo = torch.empty_like(send_dn), torch.empty_like(send_up)
        ops = []
        # order matters when prev == next (two ranks): a peer's first message is its "down" face
        if has_prev: ops.append(dist.P2POp(dist.isend, send_dn, prev, self.group))
        if has_next: ops.append(dist.P2POp(dist.isend, send_up, nxt, self.group))
        if has_next: ops.append(dist.P2POp(dist.irecv, recv_hi, nxt, self.group))
        if has_prev: ops.append(dist.P2POp(dist.irecv, recv_lo, prev, self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if has_next: hi_halo.copy_(recv_hi)
        else: hi_halo.zero_()
        if has_prev: lo_halo.copy_(recv_lo)
        else: lo_halo.zero_()
        self.e.fence()

    def _alpha(self, x, alpha):
        import torch.distributed as dist
        self.e.project(x, alpha)
        if self.world > 1 and self.ntot:
            self.e.sync()
            dist.all_reduce(alpha, group=self.group)
        self.e.fence()

    def ChebyshevFiltering(self, X, Y, W, m, a, b, a0):
        """Blocks X (input, destroyed), Y, W of the engine; returns (result block, block holding p_{m-1}(H)X0)."""
        e_ = 0.5 * (b - a)
        c = 0.5 * (b + a)
        sigma = sigma1 = e_ / (a0 - c)
        gamma = 2.0 / sigma1
        alpha = self.e.alpha_buffer(self.ntot, X.shape[0])
        self.e.fence()
        self.exchange(X)
        self._alpha(X, alpha)
        self.e.stencil_step(X, None, Y, -c, sigma1 / e_, 0.0)
        self.e.expand(Y, sigma1 / e_, alpha)
        for _ in range(1, m):
            sigma2 = 1.0 / (gamma - sigma)
            self.exchange(Y)
            self._alpha(Y, alpha)
            self.e.stencil_step(Y, X, W, -c, 2.0 * sigma2 / e_, sigma * sigma2)
            self.e.expand(W, 2.0 * sigma2 / e_, alpha)
            X, Y, W = Y, W, X
            sigma = sigma2
        self.e.sync()
        return Y, X

    def own_planes(self, blk_np):
        """The rank's own planes of a downloaded block: (ncol, nzl * Nx * Ny)."""
        ncol = blk_np.shape[0]
        v = blk_np[:, :(self.nzl + 2 * FDN) * self.plane].reshape(ncol, self.nzl + 2 * FDN, self.plane)
        return np.ascontiguousarray(v[:, FDN:FDN + self.nzl].reshape(ncol, -1))
