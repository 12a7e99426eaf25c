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


def test_multi_device_context_refuses_device_entry_points(mctx):
    from sparc_b200 import capi
    g, veff, proj, x = small_case(0, ncol=2)
    _setup(mctx, g, veff, proj)
    with pytest.raises(capi.ChefsiError, match="single-device"):
        mctx.fill_random_device(0, 1)
