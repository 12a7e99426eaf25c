"""Host-side problem description for the CheFSI filter path.

Mirrors (in numpy) the pieces of SPARC's initialisation that produce the inputs of the
hot path, so that tests and the benchmark can build the same tables SPARC would hand to
``ChebyshevFiltering``:

* finite-difference weights           -- src/initialization.c:2112-2122, fract() src/tools.c:725
* per-axis / mixed stencil tables     -- src/initialization.c:2129-2178
* lattice metric, cell_typ            -- src/initialization.c:3525-3609 (Cart2nonCart_transformMat)
* max eigenvalue of -1/2 Lap (orth)   -- src/initialization.c:2186-2197
* synthetic Kleinman-Bylander tables  -- same *layout* as GetInfluencingAtoms_nloc /
  CalculateNonlocalProjectors (src/nlocVecRoutines.c:43-571): per periodic image the grid
  points inside the rc-sphere, an ``ndc x nproj`` column-major Chi, Gamma per projector.
  The radial shapes are synthetic (no pseudopotential files travel to the GPU box).

The ctypes structures mirror ``include/chefsi_b200.h`` field for field.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

CHEFSI_MAX_FDN = 12
_COEF = C.c_double * (CHEFSI_MAX_FDN + 1)

_COEF_NAMES = (
    "D2_x", "D2_y", "D2_z", "D2_xy", "D2_xz", "D2_yz",
    "D1_x", "D1_y", "D1_z", "D1_xy", "D1_yx", "D1_xz", "D1_zx", "D1_yz", "D1_zy",
)


class ChefsiGridC(C.Structure):
    """``chefsi_grid_t``"""

    _fields_ = (
        [("Nx", C.c_int), ("Ny", C.c_int), ("Nz", C.c_int),
         ("BCx", C.c_int), ("BCy", C.c_int), ("BCz", C.c_int),
         ("FDn", C.c_int), ("cell_typ", C.c_int),
         ("dV", C.c_double),
         ("range_x", C.c_double), ("range_y", C.c_double), ("range_z", C.c_double)]
        + [(n, _COEF) for n in _COEF_NAMES]
    )


class ChefsiNlocC(C.Structure):
    """``chefsi_nloc_t``"""

    _fields_ = [
        ("n_atom", C.c_int),
        ("IP_displ", C.POINTER(C.c_int)),
        ("gamma", C.POINTER(C.c_double)),
        ("n_img", C.c_int),
        ("img_atom", C.POINTER(C.c_int)),
        ("img_ndc", C.POINTER(C.c_int)),
        ("img_coords", C.POINTER(C.c_double)),
        ("pos_off", C.POINTER(C.c_longlong)),
        ("chi_off", C.POINTER(C.c_longlong)),
        ("grid_pos", C.POINTER(C.c_int)),
        ("chi", C.POINTER(C.c_double)),
    ]


def fract(n: int, k: int) -> float:
    """src/tools.c:725"""
    nr = 1.0
    dr = 1.0
    for i in range(n - k + 1, n + 1):
        nr *= i
    for i in range(n + 1, n + k + 1):
        dr *= i
    return nr / dr


def fd_weights(FDn: int):
    """First / second derivative central weights, src/initialization.c:2112-2122."""
    w1 = np.zeros(FDn + 1)
    w2 = np.zeros(FDn + 1)
    for p in range(1, FDn + 1):
        w1[p] = (2 * (p % 2) - 1) * fract(FDn, p) / p
        w2[0] -= 2.0 / (p * p)
        w2[p] = (2 * (p % 2) - 1) * 2 * fract(FDn, p) / (p * p)
    return w1, w2


def lattice_transforms(latvec: np.ndarray):
    """LatUVec, metric (u_i.u_j), lapcT and cell_typ; src/initialization.c:3525-3609."""
    latvec = np.asarray(latvec, dtype=float).reshape(3, 3)
    U = latvec / np.linalg.norm(latvec, axis=1, keepdims=True)
    jac = float(np.linalg.det(U))
    if jac <= 0:
        raise ValueError("lattice vectors must be right handed")
    gradT = np.linalg.inv(U)  # gradT[3*j+i] = cofactor / det  ==  inverse of LatUVec
    lapcT = gradT.T @ gradT  # lapcT[3i+j] = sum_k gradT[3i+k] gradT[3j+k] with gradT stored transposed
    # the reference stores gradT[3*j+i]; lapcT[3i+j] = sum_k gradT[3i+k]*gradT[3j+k]
    g = np.empty((3, 3))
    for i in range(3):
        for j in range(3):
            g[j, i] = (U[(j + 1) % 3, (i + 1) % 3] * U[(j + 2) % 3, (i + 2) % 3]
                       - U[(j + 1) % 3, (i + 2) % 3] * U[(j + 2) % 3, (i + 1) % 3]) / jac
    lapcT = g @ g.T
    tol = 1e-6  # TEMP_TOL
    t12, t13, t23 = (abs(lapcT[0, 1]) > tol, abs(lapcT[0, 2]) > tol, abs(lapcT[1, 2]) > tol)
    cell_typ = {
        (False, False, False): 0,
        (True, False, False): 11, (False, True, False): 12, (False, False, True): 13,
        (True, True, False): 14, (False, True, True): 15, (True, False, True): 16,
        (True, True, True): 17,
    }[(t12, t13, t23)]
    metric = U @ U.T
    return U, jac, metric, lapcT, cell_typ


@dataclass
class Grid:
    """Discretisation of one unsplit domain (everything ``chefsi_grid_t`` carries)."""

    N: tuple
    BC: tuple
    FDn: int
    cell_typ: int
    dV: float
    L: tuple  # range_x, range_y, range_z
    h: tuple
    coefs: dict
    latuvec: np.ndarray = field(default_factory=lambda: np.eye(3))
    metric: np.ndarray = field(default_factory=lambda: np.eye(3))
    jacbdet: float = 1.0

    @property
    def Nd(self) -> int:
        return int(self.N[0]) * int(self.N[1]) * int(self.N[2])

    def to_c(self) -> ChefsiGridC:
        g = ChefsiGridC()
        g.Nx, g.Ny, g.Nz = (int(v) for v in self.N)
        g.BCx, g.BCy, g.BCz = (int(v) for v in self.BC)
        g.FDn = int(self.FDn)
        g.cell_typ = int(self.cell_typ)
        g.dV = float(self.dV)
        g.range_x, g.range_y, g.range_z = (float(v) for v in self.L)
        for name in _COEF_NAMES:
            arr = getattr(g, name)
            src = self.coefs[name]
            for p in range(self.FDn + 1):
                arr[p] = float(src[p])
        return g

    def max_eig_mhalf_lap(self) -> float:
        """src/initialization.c:2186-2197 (orthogonal cells, periodic)."""
        c = self.coefs
        val = c["D2_x"][0] + c["D2_y"][0] + c["D2_z"][0]
        sc = [(n - n % 2) / float(n) for n in self.N]
        for p in range(1, self.FDn + 1):
            val += 2.0 * (c["D2_x"][p] * math.cos(math.pi * p * sc[0])
                          + c["D2_y"][p] * math.cos(math.pi * p * sc[1])
                          + c["D2_z"][p] * math.cos(math.pi * p * sc[2]))
        return -0.5 * val


def make_grid(N, L, BC=(0, 0, 0), FDn=6, latvec=None) -> Grid:
    """Build the stencil tables exactly as SPARC's ``Initialize`` does.

    ``L`` are the cell lengths along the (unit) lattice vectors, ``latvec`` the lattice
    vectors (rows; any length).  Mesh spacing h = L/N for periodic and L/(N-1) for
    Dirichlet axes (SPARC counts the boundary nodes on Dirichlet axes).
    """
    N = tuple(int(v) for v in N)
    BC = tuple(int(v) for v in BC)
    L = tuple(float(v) for v in L)
    if FDn > CHEFSI_MAX_FDN:
        raise ValueError("FDn too large")
    h = tuple(L[d] / (N[d] - BC[d]) for d in range(3))
    w1, w2 = fd_weights(FDn)
    if latvec is None:
        latvec = np.eye(3)
    U, jac, metric, lapcT, cell_typ = lattice_transforms(latvec)
    inv = [1.0 / v for v in h]
    inv2 = [1.0 / (v * v) for v in h]
    z = np.zeros(FDn + 1)
    c = {n: z.copy() for n in _COEF_NAMES}
    c["D1_x"], c["D1_y"], c["D1_z"] = (w1 * inv[0], w1 * inv[1], w1 * inv[2])
    if cell_typ == 0:
        c["D2_x"], c["D2_y"], c["D2_z"] = (w2 * inv2[0], w2 * inv2[1], w2 * inv2[2])
    else:
        # src/initialization.c:2164-2177
        c["D2_x"] = lapcT[0, 0] * w2 * inv2[0]
        c["D2_y"] = lapcT[1, 1] * w2 * inv2[1]
        c["D2_z"] = lapcT[2, 2] * w2 * inv2[2]
        c["D2_xy"] = 2 * lapcT[0, 1] * w1 * inv[0]
        c["D2_xz"] = 2 * lapcT[0, 2] * w1 * inv[0]
        c["D2_yz"] = 2 * lapcT[1, 2] * w1 * inv[1]
        c["D1_xy"] = 2 * lapcT[0, 1] * w1 * inv[1]
        c["D1_yx"] = 2 * lapcT[0, 1] * w1 * inv[0]
        c["D1_xz"] = 2 * lapcT[0, 2] * w1 * inv[2]
        c["D1_zx"] = 2 * lapcT[0, 2] * w1 * inv[0]
        c["D1_yz"] = 2 * lapcT[1, 2] * w1 * inv[2]
        c["D1_zy"] = 2 * lapcT[1, 2] * w1 * inv[1]
    dV = h[0] * h[1] * h[2] * jac
    return Grid(N=N, BC=BC, FDn=FDn, cell_typ=cell_typ, dV=dV, L=L, h=h, coefs=c,
                latuvec=U, metric=metric, jacbdet=jac)


# lattice of tests/Si8/standard/Si8.inpt:3-6 (cell_typ 17)
SI8_LATVEC = np.array([
    [1.000000000000000, 0.000000000000000, 0.000000000000000],
    [-0.292371704722737, 0.956304755963035, 0.000000000000000],
    [0.173648177666930, -0.110492654830881, 0.978589640054184],
])

# synthetic lattices that hit every non-orthogonal flavour (the
# reference has no test for 13-16; these are pinned by the compiled-reference oracle)
LATVEC_BY_CELL_TYP = {
    0: np.eye(3),
    11: np.array([[1, 0, 0], [0.3, 1, 0], [0, 0, 1.0]]),
    12: np.array([[1, 0, 0], [0, 1, 0], [0.25, 0, 1.0]]),
    13: np.array([[1, 0, 0], [0, 1, 0], [0, 0.35, 1.0]]),
    17: SI8_LATVEC,
}


def _orthogonal_pairs_lattice(zero_pair: str) -> np.ndarray:
    """Lattice whose *reciprocal* metric has exactly one vanishing off-diagonal entry
    (cell_typ 14, 15, 16).  Built from the dual basis: choose gradT rows g_i with the
    wanted orthogonality, then LatUVec = inv(g)^T normalised."""
    g = {
        "23": np.array([[1.0, 0.35, 0.3], [0.35, 1.0, 0.0], [0.3, 0.0, 1.0]]),  # T23 = 0 -> 14
        "12": np.array([[1.0, 0.0, 0.3], [0.0, 1.0, 0.35], [0.3, 0.35, 1.0]]),  # T12 = 0 -> 15
        "13": np.array([[1.0, 0.3, 0.0], [0.3, 1.0, 0.35], [0.0, 0.35, 1.0]]),  # T13 = 0 -> 16
    }[zero_pair]
    # g is a symmetric positive-definite "lapcT"; take gradT = chol(g) so gradT gradT^T = g
    Lc = np.linalg.cholesky(g)
    U = np.linalg.inv(Lc).T  # LatUVec rows (up to normalisation, which rescales T consistently)
    if np.linalg.det(U) < 0:
        U[2] *= -1
    return U


LATVEC_BY_CELL_TYP[14] = _orthogonal_pairs_lattice("23")
LATVEC_BY_CELL_TYP[15] = _orthogonal_pairs_lattice("12")
LATVEC_BY_CELL_TYP[16] = _orthogonal_pairs_lattice("13")


def synthetic_veff(grid: Grid) -> np.ndarray:
    """Veff[i] = -0.5 + 0.4 cos(2 pi x/L) cos(2 pi y/L) cos(2 pi z/L)  (SURVEY.md 8d)."""
    ax = [np.cos(2 * np.pi * np.arange(n) / n) for n in grid.N]
    v = -0.5 + 0.4 * ax[2][:, None, None] * ax[1][None, :, None] * ax[0][None, None, :]
    return np.ascontiguousarray(v.reshape(-1))


@dataclass
class Projectors:
    """Flat Kleinman-Bylander tables (everything ``chefsi_nloc_t`` carries)."""

    n_atom: int
    IP_displ: np.ndarray
    gamma: np.ndarray
    img_atom: np.ndarray
    img_ndc: np.ndarray
    img_coords: np.ndarray
    pos_off: np.ndarray
    chi_off: np.ndarray
    grid_pos: np.ndarray
    chi: np.ndarray

    @property
    def n_img(self) -> int:
        return int(self.img_atom.shape[0])

    def to_c(self) -> ChefsiNlocC:
        s = ChefsiNlocC()
        s.n_atom = int(self.n_atom)
        s.n_img = self.n_img
        # keep the arrays alive on the struct
        s._keep = (self.IP_displ, self.gamma, self.img_atom, self.img_ndc, self.img_coords,
                   self.pos_off, self.chi_off, self.grid_pos, self.chi)
        s.IP_displ = self.IP_displ.ctypes.data_as(C.POINTER(C.c_int))
        s.gamma = self.gamma.ctypes.data_as(C.POINTER(C.c_double))
        s.img_atom = self.img_atom.ctypes.data_as(C.POINTER(C.c_int))
        s.img_ndc = self.img_ndc.ctypes.data_as(C.POINTER(C.c_int))
        s.img_coords = self.img_coords.ctypes.data_as(C.POINTER(C.c_double))
        s.pos_off = self.pos_off.ctypes.data_as(C.POINTER(C.c_longlong))
        s.chi_off = self.chi_off.ctypes.data_as(C.POINTER(C.c_longlong))
        s.grid_pos = self.grid_pos.ctypes.data_as(C.POINTER(C.c_int))
        s.chi = self.chi.ctypes.data_as(C.POINTER(C.c_double))
        return s


def _angular(p: int, d: np.ndarray, r: np.ndarray, rc: float) -> np.ndarray:
    """Real-harmonic-like angular factors l = 0,1,2 (synthetic stand-in for
    RealSphericalHarmonic, src/tools.c:1231)."""
    x, y, z = d[:, 0] / rc, d[:, 1] / rc, d[:, 2] / rc
    q = p % 9
    if q == 0:
        return np.ones_like(r)
    if q <= 3:
        return (x, y, z)[q - 1]
    return (x * y, y * z, z * x, x * x - y * y, 3 * z * z - (r / rc) ** 2)[q - 4]


def make_projectors(grid: Grid, frac_coords, rc, nproj, seed=7) -> Projectors:
    """Influence lists + Chi for atoms at fractional coordinates ``frac_coords``.

    ``rc`` and ``nproj`` may be scalars or per-atom sequences.  For each atom every
    periodic image (along periodic axes) whose rc-sphere reaches into the cell gets an
    entry, as in GetInfluencingAtoms_nloc (src/nlocVecRoutines.c:43-386); distances use the
    lattice metric (CalculateDistance, src/initialization.c:3687).
    """
    frac = np.atleast_2d(np.asarray(frac_coords, dtype=float))
    n_atom = frac.shape[0]
    rc = np.broadcast_to(np.asarray(rc, dtype=float), (n_atom,))
    nproj = np.broadcast_to(np.asarray(nproj, dtype=int), (n_atom,))
    rng = np.random.default_rng(seed)
    N, L, h = grid.N, grid.L, grid.h
    G = grid.metric
    IP = np.zeros(n_atom + 1, dtype=np.int32)
    IP[1:] = np.cumsum(nproj)
    gamma = np.empty(int(IP[-1]))
    for a in range(n_atom):
        sign = np.where(np.arange(nproj[a]) % 3 == 2, -1.0, 1.0)
        gamma[IP[a]:IP[a + 1]] = sign * (0.5 + 4.0 * rng.random(nproj[a]))
    # conservative image search range in grid units: |offset_d| <= rc * sqrt(Ginv_dd)
    Ginv = np.linalg.inv(G)
    img_atom, img_ndc, img_coords, pos_list, chi_list = [], [], [], [], []
    U = grid.latuvec
    for a in range(n_atom):
        s0 = frac[a] * np.asarray(L)
        reach = rc[a] * np.sqrt(np.diag(Ginv)) * 1.0001
        shifts = [(-1, 0, 1) if grid.BC[d] == 0 else (0,) for d in range(3)]
        for nz in shifts[2]:
            for ny in shifts[1]:
                for nx in shifts[0]:
                    s = s0 + np.array([nx, ny, nz]) * np.asarray(L)
                    lo, hi = [], []
                    for d in range(3):
                        lo.append(max(0, int(math.ceil((s[d] - reach[d]) / h[d]))))
                        hi.append(min(N[d] - 1, int(math.floor((s[d] + reach[d]) / h[d]))))
                    if any(lo[d] > hi[d] for d in range(3)):
                        continue
                    ii = np.arange(lo[0], hi[0] + 1)
                    jj = np.arange(lo[1], hi[1] + 1)
                    kk = np.arange(lo[2], hi[2] + 1)
                    K, J, I = np.meshgrid(kk, jj, ii, indexing="ij")
                    dn = np.stack([I * h[0] - s[0], J * h[1] - s[1], K * h[2] - s[2]], axis=-1)
                    dn = dn.reshape(-1, 3)
                    r2 = np.einsum("ni,ij,nj->n", dn, G, dn)
                    mask = r2 <= rc[a] ** 2
                    if not mask.any():
                        continue
                    pos = (K.reshape(-1) * N[1] * N[0] + J.reshape(-1) * N[0] + I.reshape(-1))[mask]
                    dcart = dn[mask] @ U  # Cartesian offsets
                    r = np.sqrt(r2[mask])
                    t = r / rc[a]
                    chi = np.empty((nproj[a], pos.size))
                    for p in range(nproj[a]):
                        radial = (1 - t * t) ** 2 * np.cos((p // 9 + 1) * 0.5 * np.pi * t)
                        chi[p] = radial * _angular(p, dcart, r, rc[a]) * (1.0 + 0.1 * p)
                    img_atom.append(a)
                    img_ndc.append(pos.size)
                    img_coords.append(s)
                    pos_list.append(pos.astype(np.int32))
                    chi_list.append(chi.reshape(-1))  # [p][i] == column-major ndc x nproj
    n_img = len(img_atom)
    pos_off = np.zeros(n_img + 1, dtype=np.int64)
    chi_off = np.zeros(n_img + 1, dtype=np.int64)
    for q in range(n_img):
        pos_off[q + 1] = pos_off[q] + img_ndc[q]
        chi_off[q + 1] = chi_off[q] + chi_list[q].size
    return Projectors(
        n_atom=n_atom,
        IP_displ=IP,
        gamma=np.ascontiguousarray(gamma),
        img_atom=np.asarray(img_atom, dtype=np.int32),
        img_ndc=np.asarray(img_ndc, dtype=np.int32),
        img_coords=np.ascontiguousarray(np.asarray(img_coords, dtype=float).reshape(-1)) if n_img else np.zeros(0),
        pos_off=pos_off,
        chi_off=chi_off,
        grid_pos=np.concatenate(pos_list).astype(np.int32) if n_img else np.zeros(0, dtype=np.int32),
        chi=np.ascontiguousarray(np.concatenate(chi_list)) if n_img else np.zeros(0),
    )


def fcc_positions(ncell: int) -> np.ndarray:
    """Fractional coordinates of an ncell^3 conventional fcc supercell (4 atoms / cell)."""
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=float)
    out = []
    for k in range(ncell):
        for j in range(ncell):
            for i in range(ncell):
                out.append((base + np.array([i, j, k])) / ncell)
    return np.concatenate(out, axis=0)


def chebyshev_bounds(grid: Grid, a0=-0.6, lambda_cutoff=0.5):
    """(a, b, a0) for the synthetic workload: b = 1.01 * MaxEig(-1/2 Lap) + 0.5."""
    return lambda_cutoff, 1.01 * grid.max_eig_mhalf_lap() + 0.5, a0


_MASK64 = (1 << 64) - 1


def _mix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def random_columns(n_per_col: int, ncol: int, first_col: int = 0, seed: int = 1) -> np.ndarray:
    """U(-0.5,0.5) start vectors, identical to ``chefsi_fill_random_device`` and
    ``oracle_fill_random`` (counter-based: value depends only on seed, column, index).
    Returns an array of shape (ncol, n_per_col) (i.e. column-major block)."""
    out = np.empty((ncol, n_per_col))
    idx = np.arange(n_per_col, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for n in range(ncol):
            col = np.uint64((first_col + n + 1) & _MASK64)
            key = _mix64(np.uint64(seed) + np.uint64(0x632BE59BD9B4E019) * col)
            hsh = _mix64(key + idx)
            out[n] = (hsh >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5
    return out
