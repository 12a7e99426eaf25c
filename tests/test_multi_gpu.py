"""One context that owns several GPUs (chefsi_create_multi): SPARC's band-parallel axis inside one process.
Columns of host blocks are split NB = ceil(ncol / ndev) per device (parallelization.c:403-428); Veff and the
projector tables are uploaded to the first device and replicated device-to-device (ncclBroadcast when NCCL is
loadable and the devices are distinct, cudaMemcpyPeer otherwise) -- Transfer_Veff_loc's MPI_Bcast
(electronicGroundState.c:1313-1385).  On a one-GPU box the device list names device 0 twice, which exercises the
split, the replication and the concurrent per-device pipelines; with two or more GPUs the NCCL path runs."""
import numpy as np
import pytest

from sparc_b200 import problem as P
from tests.cases import BOUNDS, KVEC, overlap_case, rel_fro, small_case

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _devices():
    import torch
    n = torch.cuda.device_count()
    return [0, 1] if n >= 2 else [0, 0]


@pytest.fixture(scope="module")
def mctx():
    from sparc_b200.chefsi import ChefsiContext
    c = ChefsiContext(_devices())
    yield c
    c.close()


def _setup(ctx, g, veff, proj, kvec=(0, 0, 0)):
    ctx.set_grid(g)
    ctx.set_veff(veff)
    ctx.set_projectors(proj)
    ctx.set_kpoint(kvec)


@pytest.mark.parametrize("complex_", [False, True])
@pytest.mark.parametrize("ncol", [1, 5, 8])
def test_multi_device_filter_matches_oracle(mctx, port, complex_, ncol):
    g, veff, proj, x = small_case(17, complex_=complex_, ncol=ncol)
    _setup(mctx, g, veff, proj, KVEC)
    a, b, a0 = BOUNDS
    X, Y = x.copy(), np.empty_like(x)
    mctx.ChebyshevFiltering(X, Y, 9, a, b, a0)
    Xw, Yw = port.chebyshev_filter(g, proj, veff, x, 9, a, b, a0, kvec=KVEC)
    assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL
    Hx = np.empty_like(x)
    mctx.Hamiltonian_vectors_mult(0.2, x, Hx)
    assert rel_fro(Hx, port.hamiltonian_mult(g, proj, veff, 0.2, x, kvec=KVEC)) < TOL
    info = mctx.multi_info()
    assert info["devices"] == 2 and info["broadcasts"] > 0 and info["broadcast_bytes"] > g.Nd * 8


def test_multi_device_streaming_and_overlap(mctx, port):
    """Streaming kernel + overlapping spheres on both devices, Veff replaced between calls (the per-SCF broadcast)."""
    import torch
    g, veff, proj, x = overlap_case("stream", ncol=7)
    _setup(mctx, g, veff, proj)
    a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
    for scale in (1.0, 0.7):
        v = veff * scale
        mctx.set_veff(v)
        X, Y = x.copy(), np.empty_like(x)
        mctx.ChebyshevFiltering(X, Y, 6, a, b, a0)
        Xw, Yw = port.chebyshev_filter(g, proj, v, x, 6, a, b, a0)
        assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL
    y = np.empty_like(x)
    mctx.Lap_vec_mult(-0.1, x, y)
    assert rel_fro(y, port.lap_plus_diag(g, 1.0, 0.0, -0.1, None, x)) < TOL
    assert mctx.multi_info()["nccl"] == (torch.cuda.device_count() >= 2)


@pytest.mark.parametrize("complex_", [False, True])
@pytest.mark.parametrize("ncol", [2, 9, 70])
def test_multi_device_subspace_steps(mctx, port, complex_, ncol):
    """Projection and rotation on a multi-device context: the block is split by columns, device I forms the column
    block I of Hp / Mp / Y Q and reads the other devices' Y_J through peer memory inside the GEMM kernels
    (the reference's BP2DP all-to-all + dgemm, eigenSolver.c:977-1030).  Uploaded Y (first call) and Y left resident by
    the filter (second call, host copy never written)."""
    from oracle import next_rows
    g, veff, proj, y = overlap_case("stream", ncol=ncol, complex_=complex_)
    kvec = KVEC if complex_ else (0, 0, 0)
    _setup(mctx, g, veff, proj, kvec)
    rng = np.random.default_rng(4)
    Q = rng.standard_normal((ncol, ncol)) + (1j * rng.standard_normal((ncol, ncol)) if complex_ else 0.0)
    Q = np.ascontiguousarray(Q.astype(y.dtype))

    def check(Y_host, Y_true):
        Hp, Mp = np.zeros((ncol, ncol), dtype=y.dtype), np.zeros((ncol, ncol), dtype=y.dtype)
        mctx.DP_Project_Hamiltonian(Y_host, Hp, Mp)
        Hp_w, Mp_w = next_rows.project(port, g, proj, veff, Y_true, kvec=kvec)
        assert rel_fro(Mp, Mp_w) < TOL and rel_fro(Hp, Hp_w) < TOL
        X = np.empty_like(y)
        mctx.DP_Subspace_Rotation(Q, X)
        assert rel_fro(X, next_rows.rotate(Y_true, Q)) < TOL

    check(y, y)
    a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
    mctx.subspace_reserve(ncol, is_complex=complex_)
    X0, Yh = y.copy(), np.full_like(y, np.nan)
    mctx.ChebyshevFiltering(X0, Yh, 5, a, b, a0, copy_back_x=False, keep_y=True, copy_back_y=False)
    assert np.isnan(Yh).all()
    _, Yw = port.chebyshev_filter(g, proj, veff, y, 5, a, b, a0, kvec=kvec)
    check(Yh, Yw)


def test_multi_device_single_vector_solvers(mctx, port):
    """Lanczos and the AAR solve on a multi-device context run on its first device (a single vector does not split)."""
    from oracle import next_rows
    g, veff, proj, x = small_case(17, ncol=1)
    _setup(mctx, g, veff, proj)
    lo, hi, it = mctx.Lanczos(x[0], 1e-2, 1e-2, maxit=300)
    emin, emax, j = next_rows.lanczos(port, g, proj, veff, x[0], 1e-2, 1e-2, 300)
    assert it == j and abs(lo - emin) < 1e-9 * abs(emax) and abs(hi - emax) < 1e-9 * abs(emax)
    b = np.random.default_rng(3).standard_normal(g.Nd)
    xs = np.zeros(g.Nd)
    it, rn = mctx.AAR(-0.35, xs, b, tol=1e-8, max_iter=600)
    r = b + port.lap_plus_diag(g, 1.0, 0.0, -0.35, None, xs[None, :].copy())[0]
    assert it < 600 and np.linalg.norm(r) <= 1.0001e-8 * np.linalg.norm(b)


def test_multi_device_context_refuses_device_entry_points(mctx):
    from sparc_b200 import capi
    g, veff, proj, x = small_case(0, ncol=2)
    _setup(mctx, g, veff, proj)
    with pytest.raises(capi.ChefsiError, match="single-device"):
        mctx.fill_random_device(0, 1)
