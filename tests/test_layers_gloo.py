"""world_size-2 gloo test of the N>1 path's host logic (eq_b200/layers.py): layer->rank map, per-cell
all-gather across layers, deposits routed to the owning rank, max-over-ranks timing.  The GPU solver is
replaced by an oracle-backed stand-in so this runs on CPU; results must equal a single-process run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NW, NH, NPM = 61, 31, 2.0
D_HSL = [1200.0, 640.0, 300.0]   # C4, C14 (src/eQinit.h:12-13) + one more to exercise 3 layers on 2 ranks


class OracleLayer:
    """Stand-in with GpuHSL's surface, computing with the CPU oracle."""

    def __init__(self, layer, D):
        from oracle import oracle as O
        self.O = O
        self.p = O.Problem(nW=NW, nH=NH, D=D, bc_type=(2, 2, 0, 0), bc_value=(120.0, 120.0, 0, 0))
        self.s = O.new_state(self.p)
        self.totalBoundaryFlux = 0.0

    def upload_cells(self, rec, npm):
        self.cells, self.npm = rec, npm

    def gather(self):
        return self.O.gather(self.cells, self.npm, NH, NW, self.s.u)

    def scatter(self, amount):
        self.s.u = self.O.scatter(self.cells, self.npm, NH, NW, amount, self.s.u)

    def step(self):
        self.s = self.O.step(self.p, self.s)
        self.totalBoundaryFlux = self.s.total_boundary_flux


def colony():
    from oracle import oracle as O
    return O.synthetic_colony(25, (NW - 1) / NPM, (NH - 1) / NPM, seed=5)


def couple(g):
    """Toy gene circuit: each species' secretion depends on the other species' local level."""
    return 100.0 + 0.3 * np.roll(g, 1, axis=0) - 0.1 * g


def run(group, steps=3):
    group.upload_cells(colony(), NPM)
    hist = []
    for _ in range(steps):
        g = group.gather_all()
        group.scatter_all(couple(g))
        group.step()
        hist.append((g.copy(), group.boundary_flux().copy()))
    return hist


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eq_b200.layers import LayerGroup, max_over_ranks_ms
    grp = LayerGroup(D_HSL, OracleLayer)
    owned = sorted(grp.local)
    hist = run(grp)
    tmax = max_over_ranks_ms(10.0 + rank)
    q.put((rank, owned, hist, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_layer_group_equals_single_process():
    sys.path.insert(0, ROOT)
    from eq_b200.layers import LayerGroup, layer_owner
    ref = run(LayerGroup(D_HSL, OracleLayer, rank=0, world=1))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2] and res[1][1] == [1]                  # layer l lives on rank l % world
    assert [layer_owner(l, 2) for l in range(3)] == [0, 1, 0]
    for rank, owned, hist, tmax in res:
        assert tmax == 11.0                                          # max over ranks, on every rank
        for (g, f), (g0, f0) in zip(hist, ref):
            assert np.array_equal(g, g0) and np.array_equal(f, f0)   # identical to the single-process run


def test_layout_validation():
    sys.path.insert(0, ROOT)
    from eq_b200.layers import check_layout
    with pytest.raises(ValueError):
        check_layout(0, 2)
    check_layout(2, 8)
