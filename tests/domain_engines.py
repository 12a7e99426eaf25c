"""CPU stand-in for the local pieces of the domain split (tests only): the oracle's stencil routine on the rank's
slab + a numpy restatement of the projector project / expand halves (nlocVecRoutines.c:807-831 / :866-881), so that
the orchestration of sparc_b200/domain_split.py (halo exchange, alpha all-reduce, recurrence) can be checked with
gloo on a machine without a GPU."""
import numpy as np
import torch

from oracle.bindings import Port  # checker only


class OracleSlabEngine:
    def __init__(self, grid_loc, veff_loc, proj_loc):
        self.port = Port()
        self.g = grid_loc
        self.veff = np.ascontiguousarray(veff_loc)
        self.proj = proj_loc
        self.nd = grid_loc.Nd

    def block(self, ncol):
        return torch.zeros((ncol, self.nd), dtype=torch.float64)

    def upload(self, x_np):
        return torch.from_numpy(np.ascontiguousarray(x_np).copy())

    def download(self, b):
        return b.numpy().copy()

    def alpha_buffer(self, ntot, ncol):
        return torch.zeros(ntot * ncol, dtype=torch.float64)

    def stencil_step(self, x, xprev, out, c, s1, s2):
        hx = self.port.lap_plus_diag(self.g, -0.5, 1.0, c, self.veff, x.numpy())
        res = s1 * hx
        if xprev is not None and s2 != 0.0:
            res -= s2 * xprev.numpy()
        out.copy_(torch.from_numpy(res))

    def project(self, x, alpha):
        alpha.zero_()
        p = self.proj
        if p is None or p.n_img == 0:
            return
        xn, al = x.numpy(), alpha.numpy()
        ncol = xn.shape[0]
        for J in range(p.n_img):
            a = int(p.img_atom[J])
            npj = int(p.IP_displ[a + 1] - p.IP_displ[a])
            pos = p.grid_pos[p.pos_off[J]:p.pos_off[J + 1]]
            chi = p.chi[p.chi_off[J]:p.chi_off[J + 1]].reshape(npj, pos.size)
            blk = al[p.IP_displ[a] * ncol:(p.IP_displ[a] + npj) * ncol].reshape(ncol, npj)
            blk += self.g.dV * (xn[:, pos] @ chi.T)

    def expand(self, out, scale, alpha):
        p = self.proj
        if p is None or p.n_img == 0:
            return
        on, al = out.numpy(), alpha.numpy()
        ncol = on.shape[0]
        for J in range(p.n_img):
            a = int(p.img_atom[J])
            npj = int(p.IP_displ[a + 1] - p.IP_displ[a])
            pos = p.grid_pos[p.pos_off[J]:p.pos_off[J + 1]]
            chi = p.chi[p.chi_off[J]:p.chi_off[J + 1]].reshape(npj, pos.size)
            blk = al[p.IP_displ[a] * ncol:(p.IP_displ[a] + npj) * ncol].reshape(ncol, npj)
            on[:, pos] += scale * ((blk * p.gamma[p.IP_displ[a]:p.IP_displ[a] + npj]) @ chi)

    def sync(self):
        pass

    def fence(self):
        pass

    def close(self):
        pass
