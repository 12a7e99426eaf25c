"""Domain split (z slabs) of the filter: slab construction, halo exchange, alpha all-reduce and recurrence of
sparc_b200/domain_split.py against the single-domain oracle.  CPU: the oracle is the local engine, world_size 1
(self-exchange) and world_size 2 over gloo.  GPU: the library's building blocks (chefsi_stencil_step_device,
chefsi_nloc_project_device, chefsi_nloc_expand_device) as the local engine, two slabs on one device."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
    "periodic": dict(N=(14, 12, 26), L=(7.0, 6.0, 13.0), BC=(0, 0, 0)),
    "dirichlet_z": dict(N=(12, 14, 24), L=(6.0, 7.0, 12.0), BC=(0, 0, 1)),
}


def _problem(case):
    from sparc_b200 import problem as P
    kw = CASES[case]
    g = P.make_grid(kw["N"], kw["L"], BC=kw["BC"])
    veff = P.synthetic_veff(g)
    proj = P.make_projectors(g, np.array([[0.1, 0.2, 0.48], [0.6, 0.55, 0.98], [0.3, 0.7, 0.2]]), rc=[2.1, 1.7, 1.9],
                             nproj=[5, 9, 4])
    x = P.random_columns(g.Nd, 3, seed=4)
    return g, veff, proj, x


def _run_slab(engine_factory, case, rank, world, group, m=6):
    from sparc_b200 import domain_split as DS
    from tests.cases import BOUNDS
    g, veff, proj, x = _problem(case)
    slabs = DS.z_slabs(g.N[2], world)
    z0, z1 = slabs[rank]
    gl = DS.slab_grid(g, z0, z1)
    eng = engine_factory(gl, DS.slab_field(g, veff, z0, z1), DS.slab_projectors(g, proj, z0, z1))
    f = DS.DomainSplitFilter(eng, g, (z0, z1), rank, world, int(proj.IP_displ[-1]), group)
    X = eng.upload(DS.slab_field(g, x, z0, z1))
    Y, W = eng.block(x.shape[0]), eng.block(x.shape[0])
    a, b, a0 = BOUNDS
    Yf, Xf = f.ChebyshevFiltering(X, Y, W, m, a, b, a0)
    out = f.own_planes(eng.download(Yf)), f.own_planes(eng.download(Xf))
    eng.close()
    return out


def _want(case, m=6):
    from oracle.bindings import Port
    from tests.cases import BOUNDS
    g, veff, proj, x = _problem(case)
    a, b, a0 = BOUNDS
    Xw, Yw = Port().chebyshev_filter(g, proj, veff, x, m, a, b, a0)
    return Yw, Xw


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_slab_tables():
    from sparc_b200 import domain_split as DS
    g, veff, proj, x = _problem("periodic")
    slabs = DS.z_slabs(g.N[2], 3)
    assert slabs[0][0] == 0 and slabs[-1][1] == g.N[2] and all(b - a >= DS.FDN for a, b in slabs)
    with pytest.raises(ValueError):
        DS.z_slabs(20, 4)
    npts = 0
    for z0, z1 in slabs:
        pl = DS.slab_projectors(g, proj, z0, z1)
        npts += int(pl.img_ndc.sum())
        gl = DS.slab_grid(g, z0, z1)
        assert gl.N[2] == z1 - z0 + 12 and gl.BC[2] == 1 and pl.grid_pos.max() < gl.Nd
        v = DS.slab_field(g, veff, z0, z1).reshape(gl.N[2], -1)
        assert np.array_equal(v[6], veff.reshape(g.N[2], -1)[z0])
        assert np.array_equal(v[0], veff.reshape(g.N[2], -1)[(z0 - 6) % g.N[2]])
    assert npts == int(proj.img_ndc.sum())


@pytest.mark.parametrize("case", list(CASES))
def test_single_rank_self_exchange_matches_oracle(case):
    from oracle.bindings import build_port
    from tests.domain_engines import OracleSlabEngine
    build_port()
    Y, X = _run_slab(OracleSlabEngine, case, 0, 1, None)
    Yw, Xw = _want(case)
    assert _rel(Y, Yw) < 1e-12 and _rel(X, Xw) < 1e-12


def _worker(rank, world, port_no, out_dir, case):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from tests.domain_engines import OracleSlabEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Y, X = _run_slab(OracleSlabEngine, case, rank, world, None)
    np.save(os.path.join(out_dir, f"y{rank}.npy"), Y)
    np.save(os.path.join(out_dir, f"x{rank}.npy"), X)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,world", [("periodic", 2), ("dirichlet_z", 2), ("periodic", 3)])
def test_multi_rank_gloo_matches_oracle(tmp_path, case, world):
    from oracle.bindings import build_port
    build_port()
    port_no = 29900 + (os.getpid() % 90) + 10 * world
    mp.spawn(_worker, args=(world, port_no, str(tmp_path), case), nprocs=world, join=True)
    Yw, Xw = _want(case)
    Y = np.concatenate([np.load(tmp_path / f"y{r}.npy") for r in range(world)], axis=1)
    X = np.concatenate([np.load(tmp_path / f"x{r}.npy") for r in range(world)], axis=1)
    assert _rel(Y, Yw) < 1e-12 and _rel(X, Xw) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_gpu_building_blocks_two_slabs_one_device(case):
    """The library's domain-split entry points as the local engine: two slabs processed on cuda:0, the halo
    exchange and the alpha reduction done by hand between them, against the single-domain oracle."""
    import torch
    from sparc_b200 import domain_split as DS
    from tests.cases import BOUNDS
    g, veff, proj, x = _problem(case)
    world, m = 2, 6
    slabs = DS.z_slabs(g.N[2], world)
    ntot = int(proj.IP_displ[-1])
    eng, flt, blk = [], [], []
    for r, (z0, z1) in enumerate(slabs):
        e = DS.GpuSlabEngine(0, DS.slab_grid(g, z0, z1), DS.slab_field(g, veff, z0, z1), DS.slab_projectors(g, proj, z0, z1))
        eng.append(e)
        flt.append(DS.DomainSplitFilter(e, g, (z0, z1), r, world, ntot, None))
        blk.append([e.upload(DS.slab_field(g, x, z0, z1)), e.block(x.shape[0]), e.block(x.shape[0])])
    F = DS.FDN
    plane = g.N[0] * g.N[1]

    def exchange(which):
        views = []
        for r, (z0, z1) in enumerate(slabs):
            eng[r].sync()
            b = blk[r][which]
            views.append(b[:, :(z1 - z0 + 2 * F) * plane].view(b.shape[0], z1 - z0 + 2 * F, plane))
        per = g.BC[2] == 0
        for r in range(world):
            nzl = slabs[r][1] - slabs[r][0]
            up, dn = (r + 1) % world, (r - 1) % world
            nzu = slabs[up][1] - slabs[up][0]
            nzd = slabs[dn][1] - slabs[dn][0]
            if per or r < world - 1: views[r][:, F + nzl:] = views[up][:, F:2 * F]
            else: views[r][:, F + nzl:] = 0
            if per or r > 0: views[r][:, :F] = views[dn][:, nzd:nzd + F]
            else: views[r][:, :F] = 0
        torch.cuda.synchronize()

    a, b, a0 = BOUNDS
    e_, c = 0.5 * (b - a), 0.5 * (b + a)
    sigma = sigma1 = e_ / (a0 - c)
    gamma = 2.0 / sigma1
    al = [e.alpha_buffer(ntot, x.shape[0]) for e in eng]

    def reduce_alpha(src):
        for r in range(world):
            eng[r].project(blk[r][src], al[r])
            eng[r].sync()
        tot = al[0] + al[1]
        for r in range(world):
            al[r].copy_(tot)
        torch.cuda.synchronize()

    X, Y, W = 0, 1, 2
    torch.cuda.synchronize()
    exchange(X); reduce_alpha(X)
    for r in range(world):
        eng[r].stencil_step(blk[r][X], None, blk[r][Y], -c, sigma1 / e_, 0.0)
        eng[r].expand(blk[r][Y], sigma1 / e_, al[r])
    for _ in range(1, m):
        sigma2 = 1.0 / (gamma - sigma)
        exchange(Y); reduce_alpha(Y)
        for r in range(world):
            eng[r].stencil_step(blk[r][Y], blk[r][X], blk[r][W], -c, 2.0 * sigma2 / e_, sigma * sigma2)
            eng[r].expand(blk[r][W], 2.0 * sigma2 / e_, al[r])
        X, Y, W = Y, W, X
        sigma = sigma2
    Yg = np.concatenate([flt[r].own_planes(eng[r].download(blk[r][Y])) for r in range(world)], axis=1)
    Yw, _ = _want(case, m)
    for e in eng:
        e.close()
    assert _rel(Yg, Yw) < 1e-10


def _nccl_worker(rank, world, port_no, out_dir, case):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from sparc_b200 import domain_split as DS
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    Y, X = _run_slab(lambda gl, vl, pl: DS.GpuSlabEngine(rank, gl, vl, pl), case, rank, world, None)
    np.save(os.path.join(out_dir, f"y{rank}.npy"), Y)
    np.save(os.path.join(out_dir, f"x{rank}.npy"), X)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_gpu_two_ranks_nccl(tmp_path, case):
    """One process per GPU, NCCL send/recv halo exchange + all-reduce of alpha (needs two visible GPUs)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port_no = 29700 + (os.getpid() % 200)
    mp.spawn(_nccl_worker, args=(2, port_no, str(tmp_path), case), nprocs=2, join=True)
    Yw, Xw = _want(case)
    Y = np.concatenate([np.load(tmp_path / f"y{r}.npy") for r in range(2)], axis=1)
    X = np.concatenate([np.load(tmp_path / f"x{r}.npy") for r in range(2)], axis=1)
    assert _rel(Y, Yw) < 1e-10 and _rel(X, Xw) < 1e-10
