"""Host-side mirror of the reference's operator interface for the CheFSI filter path.

The method names, argument meaning and in/out behaviour follow the reference functions they
stand in for, so the parity tests read like calls into SPARC:

    ChebyshevFiltering        src/eigenSolver.c:722     (X in/out -> p_{m-1}(H)X0, Y out -> p_m(H)X0)
    ChebyshevFiltering_kpt    src/eigenSolverKpt.c:458
    Hamiltonian_vectors_mult  src/hamiltonianVecRoutines.c:45    Hx = (-1/2 Lap + Veff + c) x + Vnl x
    Hamiltonian_vectors_mult_kpt                          :132

Blocks of orbitals are arrays of shape ``(ncol, ld)`` (column n at ``n*ld``: the reference's
column-major layout).  numpy arrays, pinned torch CPU tensors (host entry points) and torch CUDA
tensors / raw device addresses (``*_device`` entry points) are accepted; torch is only plumbing
for memory, the arithmetic is all inside ``libchefsi_b200.so``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .problem import Grid, Projectors


def _addr(obj) -> int:
    if obj is None:
        return 0
    if isinstance(obj, int):
        return obj
    if isinstance(obj, np.ndarray):
        if not obj.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return obj.ctypes.data
    if hasattr(obj, "data_ptr"):  # torch tensor
        if not obj.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return obj.data_ptr()
    raise TypeError(f"cannot take the address of {type(obj)}")


def _is_complex(obj) -> bool:
    if isinstance(obj, np.ndarray):
        return np.iscomplexobj(obj)
    if hasattr(obj, "is_complex"):
        return bool(obj.is_complex())
    raise TypeError("need an array to infer real/complex")


class ChefsiContext:
    """One CUDA device's instance of the filter path (``chefsi_ctx_t``)."""

    FLAG_NO_X_COPYBACK = 1
    FLAG_KEEP_Y = 2
    FLAG_NO_Y_COPYBACK = 4

    def __init__(self, device=0):
        """``device``: a CUDA ordinal, or a list of ordinals for one context that owns several GPUs
        (``chefsi_create_multi``: columns of host blocks are split like SPARC's NP_BAND_PARAL)."""
        self._lib = capi.load_library()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(v) for v in device])
            rc = self._lib.chefsi_create_multi(C.byref(h), devs, len(device))
            self.device = int(device[0])
        else:
            rc = self._lib.chefsi_create(C.byref(h), int(device))
            self.device = int(device)
        if rc != 0:
            raise capi.ChefsiError(self._lib.chefsi_last_error(None).decode())
        self._h = h
        self.grid = None
        self._keep = []

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.chefsi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise capi.ChefsiError(self._lib.chefsi_last_error(self._h).decode())

    # -- problem description ------------------------------------------------------------------
    def set_grid(self, grid: Grid):
        g = grid.to_c()
        self._check(self._lib.chefsi_set_grid(self._h, C.byref(g)))
        self.grid = grid

    def set_projectors(self, proj: Projectors | None):
        if proj is None:
            self._check(self._lib.chefsi_set_projectors(self._h, None))
            return
        s = proj.to_c()
        self._check(self._lib.chefsi_set_projectors(self._h, C.byref(s)))

    def set_veff(self, veff):
        if veff is not None:
            veff = np.ascontiguousarray(veff, dtype=np.float64)
            assert veff.size == self.grid.Nd
        self._check(self._lib.chefsi_set_veff(self._h, _addr(veff)))

    def set_kpoint(self, k):
        self._check(self._lib.chefsi_set_kpoint(self._h, float(k[0]), float(k[1]), float(k[2])))

    @property
    def device_ld(self) -> int:
        return int(self._lib.chefsi_device_ld(self._h))

    # -- reference-named host entry points ------------------------------------------------------
    def ChebyshevFiltering(self, X, Y, m, a, b, a0, copy_back_x=True, keep_y=False, copy_back_y=True):
        """X, Y: real arrays (ncol, ld).  X is overwritten with p_{m-1}(H)X0, Y with p_m(H)X0.
        keep_y: leave Y on the device for DP_Project_Hamiltonian / DP_Subspace_Rotation (after subspace_reserve)."""
        ncol, ldi = X.shape
        flags = 0 if copy_back_x else self.FLAG_NO_X_COPYBACK
        if keep_y:
            flags |= self.FLAG_KEEP_Y | (0 if copy_back_y else self.FLAG_NO_Y_COPYBACK)
        fn = self._lib.chefsi_chebyshev_filter_kpt if _is_complex(X) else self._lib.chefsi_chebyshev_filter
        self._check(fn(self._h, _addr(X), ldi, _addr(Y), Y.shape[1], ncol, int(m), float(a), float(b), float(a0), flags))

    ChebyshevFiltering_kpt = ChebyshevFiltering

    def Hamiltonian_vectors_mult(self, c, x, Hx):
        ncol, ldi = x.shape
        fn = self._lib.chefsi_hamiltonian_mult_kpt if _is_complex(x) else self._lib.chefsi_hamiltonian_mult
        self._check(fn(self._h, ncol, float(c), _addr(x), ldi, _addr(Hx), Hx.shape[1]))

    Hamiltonian_vectors_mult_kpt = Hamiltonian_vectors_mult

    def AAR(self, c, x, b, omega=0.6, beta=0.6, m=7, p=6, tol=1e-8, max_iter=1000):
        """Solve -(Lap + c) x = b in place (src/linearSolver.c:38 with poisson_residual + Jacobi_preconditioner);
        returns (iterations, ||r||)."""
        it, rn = C.c_int(0), C.c_double(0)
        self._check(self._lib.chefsi_poisson_aar(self._h, float(c), _addr(x), _addr(b), float(omega), float(beta), int(m), int(p),
                                                 float(tol), int(max_iter), C.byref(it), C.byref(rn)))
        return it.value, rn.value

    def Lanczos(self, x0, tol_min, tol_max, maxit=1000):
        """(eigmin, eigmax, iterations) of H by Lanczos from x0 (src/eigenSolver.c:1920; complex x0: Lanczos_kpt,
        src/eigenSolverKpt.c:1361, at the k-point set with set_kpoint), vectors resident on the device."""
        cplx = np.iscomplexobj(x0)
        x0 = np.ascontiguousarray(x0, dtype=np.complex128 if cplx else np.float64).reshape(-1)
        lo, hi, it = C.c_double(0), C.c_double(0), C.c_int(0)
        fn = self._lib.chefsi_lanczos_kpt if cplx else self._lib.chefsi_lanczos
        self._check(fn(self._h, _addr(x0), float(tol_min), float(tol_max), int(maxit), C.byref(lo), C.byref(hi), C.byref(it)))
        return lo.value, hi.value, it.value

    # -- Rayleigh-Ritz steps on the resident block (src/eigenSolver.c:939-1086, 1386-1443) ------------------------
    def subspace_reserve(self, ncol, is_complex=False):
        fn = self._lib.chefsi_subspace_reserve_kpt if is_complex else self._lib.chefsi_subspace_reserve
        self._check(fn(self._h, int(ncol)))

    def DP_Project_Hamiltonian(self, Y, Hp, Mp):
        """Hp = Y^H H Y, Mp = Y^H Y, column-major ncol x ncol: the numpy arrays hold Hp[n, m] = element (m, n)."""
        ncol, ldy = Y.shape
        fn = self._lib.chefsi_subspace_project_kpt if _is_complex(Y) else self._lib.chefsi_subspace_project
        self._check(fn(self._h, _addr(Y), ldy, ncol, _addr(Hp), _addr(Mp), Hp.shape[1]))

    DP_Project_Hamiltonian_kpt = DP_Project_Hamiltonian

    def DP_Subspace_Rotation(self, Q, X):
        """X = Y Q with the resident Y; Q given as the reference stores it: column-major ncol x ncol, i.e. the numpy
        array Q[n, m] holds element (m, n).  Q = None: the eigenvectors DP_Solve_Generalized_EigenProblem left on the
        device."""
        fn = self._lib.chefsi_subspace_rotate_kpt if _is_complex(X) else self._lib.chefsi_subspace_rotate
        if Q is None:
            self._check(fn(self._h, None, 0, X.shape[0], _addr(X), X.shape[1]))
        else:
            self._check(fn(self._h, _addr(Q), Q.shape[1], Q.shape[0], _addr(X), X.shape[1]))

    DP_Subspace_Rotation_kpt = DP_Subspace_Rotation

    def DP_Solve_Generalized_EigenProblem(self, ncol, Hp=None, Mp=None, is_complex=False, want_Q=True):
        """Hp q = lambda Mp q (src/eigenSolver.c:1262, LAPACKE_dsygvd; k-point: src/eigenSolverKpt.c:836, zhegvd) on the
        device.  Hp = Mp = None: the matrices the last DP_Project_Hamiltonian left there.  Returns (lambda, Q) with Q in
        the reference's column-major storage (numpy Q[n, :] = eigenvector n), or (lambda, None)."""
        if Hp is not None:
            is_complex = _is_complex(Hp)
        lam = np.zeros(ncol)
        Q = np.zeros((ncol, ncol), dtype=np.complex128 if is_complex else np.float64) if want_Q else None
        fn = self._lib.chefsi_subspace_eig_kpt if is_complex else self._lib.chefsi_subspace_eig
        self._check(fn(self._h, int(ncol), _addr(Hp) if Hp is not None else None, _addr(Mp) if Mp is not None else None,
                       Hp.shape[1] if Hp is not None else 0, _addr(lam), _addr(Q) if want_Q else None, ncol))
        return lam, Q

    def band_store(self, max_blocks):
        """Keep device copies of up to max_blocks rotated blocks for CalculateDensity_psi (0: off)."""
        self._check(self._lib.chefsi_band_store(self._h, int(max_blocks)))

    def CalculateDensity_psi(self, X, g, rho):
        """rho += sum_n g[n] |X[n]|^2 for one k-point / spin block (src/electronDensity.c:104-200, loop body)."""
        ncol, ldx = X.shape
        g = np.ascontiguousarray(g, dtype=np.float64)
        assert g.size == ncol and rho.dtype == np.float64 and rho.flags.c_contiguous
        fn = self._lib.chefsi_density_accumulate_kpt if _is_complex(X) else self._lib.chefsi_density_accumulate
        self._check(fn(self._h, _addr(X), ldx, ncol, _addr(g), _addr(rho)))

    def Lap_vec_mult(self, c, x, Lapx, a=1.0):
        """Lapx = (a Lap + c) x (src/lapVecRoutines.c:37: a = 1; no potential, no projectors)."""
        ncol, ldi = x.shape
        fn = self._lib.chefsi_laplacian_mult_kpt if _is_complex(x) else self._lib.chefsi_laplacian_mult
        self._check(fn(self._h, ncol, float(a), float(c), _addr(x), ldi, _addr(Lapx), Lapx.shape[1]))

    def Gradient_vectors_dir(self, c, x, Dx, dir, kdir=0.0):
        """Dx = (D_dir + c) x along lattice direction dir (src/gradVecRoutines.c:32; complex x: Gradient_vectors_dir_kpt,
        src/gradVecRoutinesKpt.c:35, kdir = the k-point component the reference passes as *kpt_vec)."""
        ncol, ldi = x.shape
        if _is_complex(x):
            self._check(self._lib.chefsi_gradient_mult_kpt(self._h, ncol, float(c), _addr(x), ldi, _addr(Dx), Dx.shape[1],
                                                           int(dir), float(kdir)))
        else:
            self._check(self._lib.chefsi_gradient_mult(self._h, ncol, float(c), _addr(x), ldi, _addr(Dx), Dx.shape[1], int(dir)))

    # -- device-resident entry points -------------------------------------------------------------
    def filter_device(self, bufA, bufB, bufC, ncol, m, a, b, a0, is_complex=False):
        """Enqueue one filter on device buffers; returns (y_slot, x_slot) in {0,1,2}."""
        ys, xs = C.c_int(-1), C.c_int(-1)
        fn = self._lib.chefsi_chebyshev_filter_kpt_device if is_complex else self._lib.chefsi_chebyshev_filter_device
        self._check(fn(self._h, _addr(bufA), _addr(bufB), _addr(bufC), int(ncol), int(m), float(a), float(b),
                       float(a0), C.byref(ys), C.byref(xs)))
        return ys.value, xs.value

    def hamiltonian_device(self, ncol, c, x, Hx, is_complex=False):
        fn = self._lib.chefsi_hamiltonian_mult_kpt_device if is_complex else self._lib.chefsi_hamiltonian_mult_device
        self._check(fn(self._h, int(ncol), float(c), _addr(x), _addr(Hx)))

    # -- domain-split building blocks (real data; see domain_split.py) -----------------------------
    def stencil_step_device(self, x, xprev, out, ncol, c, s1, s2):
        """out = s1 ((-1/2 Lap + Veff + c) x) - s2 xprev on device blocks, without the projector part."""
        self._check(self._lib.chefsi_stencil_step_device(self._h, _addr(x), _addr(xprev), _addr(out), int(ncol),
                                                         float(c), float(s1), float(s2)))

    def nloc_project_device(self, x, ncol, alpha_out):
        """Local projector inner products of x into the caller's device buffer (n_proj_total * ncol doubles,
        atom a at IP_displ[a] * ncol, [column][projector])."""
        self._check(self._lib.chefsi_nloc_project_device(self._h, _addr(x), int(ncol), _addr(alpha_out)))

    def nloc_expand_device(self, out, ncol, scale, alpha_in):
        """out += scale * Chi Gamma alpha_in (the all-reduced buffer)."""
        self._check(self._lib.chefsi_nloc_expand_device(self._h, _addr(out), int(ncol), float(scale), _addr(alpha_in)))

    def fill_random_device(self, buf, ncol, first_col=0, seed=1, is_complex=False):
        self._check(self._lib.chefsi_fill_random_device(self._h, _addr(buf), int(ncol), int(first_col), int(seed),
                                                        int(bool(is_complex))))

    def pack_device(self, dense, ld_dense, packed, ncol, is_complex=False):
        """dense device block (column n at n*ld_dense) -> internal layout."""
        self._check(self._lib.chefsi_pack_device(self._h, _addr(dense), int(ld_dense), _addr(packed), int(ncol),
                                                 int(bool(is_complex))))

    def unpack_device(self, packed, dense, ld_dense, ncol, is_complex=False):
        self._check(self._lib.chefsi_unpack_device(self._h, _addr(packed), _addr(dense), int(ld_dense), int(ncol),
                                                   int(bool(is_complex))))

    def synchronize(self):
        self._check(self._lib.chefsi_synchronize(self._h))

    def set_profiling(self, on: bool):
        self._check(self._lib.chefsi_set_profiling(self._h, int(bool(on))))

    def stats(self) -> dict:
        s = capi.ChefsiStats()
        self._check(self._lib.chefsi_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def multi_info(self) -> dict:
        nd, nccl = C.c_int(0), C.c_int(0)
        calls, nbytes = C.c_ulonglong(0), C.c_ulonglong(0)
        self._check(self._lib.chefsi_multi_info(self._h, C.byref(nd), C.byref(nccl), C.byref(calls), C.byref(nbytes)))
        return {"devices": nd.value, "nccl": bool(nccl.value), "broadcasts": calls.value, "broadcast_bytes": nbytes.value}

    @property
    def stream(self) -> int:
        return int(self._lib.chefsi_stream(self._h) or 0)
