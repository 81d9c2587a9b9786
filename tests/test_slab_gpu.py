"""Row-slab decomposition over 2 GPUs (eqgpu_create_slab): halo exchange + CG all-reduce over NCCL.
The decomposed solve must equal the single-GPU solve and the oracle.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NW, NH, NPM = 384, 256, 2.0
BC = dict(bc_type=(2, 1, 0, 1), bc_value=(120.0, 2.0, 0.0, 0.5))
# D = 50: the hierarchy ends mass-dominated while every rank still has >= 12 rows -> fused tile kernels on
# slabs (6-row halos).  D = 1200: the coarse solve needs a high Chebyshev degree -> unfused fallback (1-row halos).
CASES = {"fused": 50.0, "unfused": 1200.0}


def problem(D, nh=NH):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    p = O.Problem(nW=NW, nH=nh, D=D, **BC)
    cells = O.synthetic_colony(300, p.W, p.H, seed=11)
    return O, p, cells


def worker(rank, world, q_id, q_out, kase, transport="peer", nh=NH):
    sys.path.insert(0, ROOT)
    os.environ["EQGPU_SLAB_PEER"] = "1" if transport == "peer" else "0"
    os.environ.setdefault("EQGPU_PEER_TIMEOUT_MS", "5000")   # a lost flag fails the test in seconds, it never hangs the GPU
    import torch
    torch.cuda.set_device(rank)
    import eq_b200 as E
    O, p, cells = problem(CASES[kase], nh)
    if rank == 0:
        uid = E.nccl_unique_id()
        for _ in range(world - 1):
            q_id.put(uid)
    else:
        uid = q_id.get(timeout=120)
    g = E.GpuHSL(NW, nh, D=CASES[kase], device=rank, slab=(rank, world, uid), **BC)
    assert g.path()["slab"] and g.path()["slab_fused"] == (kase == "fused"), g.path()
    g0, g1 = g.slab_rows()
    assert (g0, g1) == E.slab_plan(nh, world, rank)[0][:2]
    g.upload_cells(cells, NPM)
    rng = np.random.default_rng(3)
    u = rng.uniform(0, 5, NW * nh)
    g.set_field(u)
    hist = []
    for _ in range(3):
        s = g.gather()
        g.scatter(100.0 + 0.5 * s)
        g.step()
        hist.append((s, g.totalBoundaryFlux, g.stats().iterations, g.last_guess()))
    out = np.zeros(NW * nh)
    g.get_field(out)
    q_out.put((rank, g0, g1, out[g0 * NW:g1 * NW].copy(), hist, g.comm_stats()))
    g.close()


@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("kase", list(CASES))
def test_two_gpu_slab_equals_single_gpu_and_oracle(kase, transport):
    """transport "peer": halo rows pulled from the neighbour's memory (CUDA IPC) and scalar all-reduces by one warp over
    peer memory, no NCCL call on the data path; "nccl": ncclSend/Recv + ncclAllReduce (EQGPU_SLAB_PEER=0)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import eq_b200 as E
    O, p, cells = problem(CASES[kase])
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, q_id, q_out, kase, transport)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted([q_out.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == NH   # contiguous partition
    for r in range(2):   # the transport asked for is the one that ran
        cs = res[r][5]
        if transport == "peer":
            assert cs["peer_exchange_kernels"] > 0 and cs["peer_allreduces"] > 0, cs
        else:
            assert cs["peer_exchange_kernels"] == -1, cs
        assert cs["halo_bytes_sent"] > 0 and cs["allreduce_calls"] > 0, cs
    field = np.concatenate([res[0][3], res[1][3]])
    # oracle with the same coupling
    rng = np.random.default_rng(3)
    s = O.new_state(p)
    s.u = rng.uniform(0, 5, NW * NH)
    hist = []
    for _ in range(3):
        smp = O.gather(cells, NPM, NH, NW, s.u)
        s.u = O.scatter(cells, NPM, NH, NW, 100.0 + 0.5 * smp, s.u)
        s = O.step(p, s)
        hist.append((smp, s.total_boundary_flux))
    rel = np.linalg.norm(field - s.u) / np.linalg.norm(s.u)
    assert rel < 1e-8, rel
    for r in range(2):
        for (smp, flux, it, guess), (smp0, flux0) in zip(res[r][4], hist):
            assert np.allclose(smp, smp0, rtol=1e-9, atol=1e-12)        # rank-summed samples, same on both ranks
            assert abs(flux - flux0) <= 1e-7 * max(abs(flux0), 1.0)
            assert 0 < it < 40
        # warm start on slabs (fused path): steps 2 and 3 start from the previous solution or an extrapolation
        # of it (history copied on the device, halo rows exchanged); both ranks take the same rank-summed decision
        guesses = [h[3] for h in res[r][4]]
        assert guesses[0] in (0, 1)
        if kase == "fused":
            assert all(q >= 2 for q in guesses[1:]), guesses
        assert guesses == [h[3] for h in res[0][4]]


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_four_gpu_slab_interior_ranks(transport):
    """Four ranks: ranks 1 and 2 have a neighbour on BOTH sides (two staging buffers in, two out per exchange; the scalar
    all-reduce sums four slots in rank order).  384 x 768 nodes, fused tile kernels on slabs; field against the oracle."""
    import torch
    world, nh, kase = 4, 768, "fused"
    if torch.cuda.device_count() < world:
        pytest.skip("needs 4 GPUs")
    import torch.multiprocessing as mp
    O, p, cells = problem(CASES[kase], nh)
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, q_id, q_out, kase, transport, nh)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted([q_out.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert res[0][1] == 0 and res[-1][2] == nh and all(res[r][2] == res[r + 1][1] for r in range(world - 1))
    for r in range(world):
        cs = res[r][5]
        assert (cs["peer_exchange_kernels"] > 0) == (transport == "peer"), cs
    field = np.concatenate([res[r][3] for r in range(world)])
    rng = np.random.default_rng(3)
    s = O.new_state(p)
    s.u = rng.uniform(0, 5, NW * nh)
    for _ in range(3):
        smp = O.gather(cells, NPM, nh, NW, s.u)
        s.u = O.scatter(cells, NPM, nh, NW, 100.0 + 0.5 * smp, s.u)
        s = O.step(p, s)
    rel = np.linalg.norm(field - s.u) / np.linalg.norm(s.u)
    assert rel < 1e-8, rel
    its = [[h[2] for h in res[r][4]] for r in range(world)]
    assert all(i == its[0] for i in its), its   # bit-identical rank-ordered sums: every rank stops at the same iteration
