"""world_size-2 test (gloo, CPU) of the N>1 host logic: band partition of the columns, broadcast
of Veff + projector tables, and that filtering the two column slices independently (with the
oracle standing in for the device) reproduces the single-rank result -- the filter has no
inter-band communication."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_no, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle.bindings import Port
    from sparc_b200.partition import band_partition
    from sparc_b200.replicate import broadcast_problem
    from sparc_b200 import problem as P
    from tests.cases import BOUNDS, small_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, veff, proj, _ = small_case(0, N=(12, 10, 14), L=(6.0, 5.2, 7.0))
    if rank != 0:  # only rank 0 owns the tables before the broadcast
        veff, proj = None, None
    veff, proj = broadcast_problem(g, veff, proj, src=0)
    ns = 5
    first, n = band_partition(ns, world, rank)
    x = P.random_columns(g.Nd, n, first_col=first, seed=1)
    a, b, a0 = BOUNDS
    _, y = Port().chebyshev_filter(g, proj, veff, x, 6, a, b, a0)
    np.save(os.path.join(out_dir, f"y{rank}.npy"), y)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_band_split(tmp_path):
    from oracle.bindings import Port, build_port
    from sparc_b200 import problem as P
    from tests.cases import BOUNDS, small_case
    build_port()
    port_no = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port_no, str(tmp_path)), nprocs=2, join=True)
    g, veff, proj, _ = small_case(0, N=(12, 10, 14), L=(6.0, 5.2, 7.0))
    x = P.random_columns(g.Nd, 5, first_col=0, seed=1)
    a, b, a0 = BOUNDS
    _, want = Port().chebyshev_filter(g, proj, veff, x, 6, a, b, a0)
    got = np.concatenate([np.load(tmp_path / "y0.npy"), np.load(tmp_path / "y1.npy")])
    assert got.shape == want.shape
    assert np.array_equal(got, want)
