"""Replication of Veff and the projector tables across the band-parallel ranks.

SPARC re-broadcasts Veff every SCF iteration with three MPI_Bcast calls
(Transfer_Veff_loc, src/electronicGroundState.c:1313-1385) and recomputes the projector tables
on every rank once per ionic step.  Here the ranks are the GPUs of one B200 box: rank ``src``
owns the tables and one ``torch.distributed`` broadcast per array (NCCL over NVLink on the GPU
box, gloo in the CPU tests) replicates them.  This is the only collective on the path; the filter
itself has none.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .problem import Grid, Projectors

_PROJ_FIELDS = ("IP_displ", "gamma", "img_atom", "img_ndc", "img_coords", "pos_off", "chi_off", "grid_pos", "chi")


def _bcast_array(arr, meta, src, device):
    dtype, shape = meta
    if arr is None:
        arr = np.empty(shape, dtype=dtype)
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
    if t.numel():
        dist.broadcast(t, src=src)
    return t.cpu().numpy()


def broadcast_veff(veff, nd: int, src: int = 0, device=None):
    """Per-SCF step: replicate Veff (Nd doubles)."""
    device = device or torch.device("cpu")
    rank = dist.get_rank()
    return _bcast_array(veff if rank == src else None, (np.float64, (nd,)), src, device)


def broadcast_problem(grid: Grid, veff, proj: Projectors | None, src: int = 0, device=None):
    """Replicate Veff and the projector tables from ``src``; returns (veff, proj) on every rank."""
    device = device or torch.device("cpu")
    rank = dist.get_rank()
    veff = broadcast_veff(veff, grid.Nd, src, device)
    meta = [None]
    if rank == src:
        meta = [None if proj is None else
                {"n_atom": proj.n_atom, "arrays": {f: (getattr(proj, f).dtype.str, getattr(proj, f).shape) for f in _PROJ_FIELDS}}]
    dist.broadcast_object_list(meta, src=src)
    if meta[0] is None:
        return veff, None
    out = {}
    for f in _PROJ_FIELDS:
        dt, shape = meta[0]["arrays"][f]
        out[f] = _bcast_array(getattr(proj, f) if rank == src else None, (np.dtype(dt), tuple(shape)), src, device)
    return veff, Projectors(n_atom=meta[0]["n_atom"], **out)
