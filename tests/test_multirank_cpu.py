"""N > 1 host logic on CPU: two gloo ranks partition the axis with pdgpu_partition /
pdgpu_slab_layout (the same functions the CUDA halo exchange uses), exchange their
boundary planes with torch.distributed and must reconstruct exactly the ghost planes a
single-domain array holds."""
import ctypes as C
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_axial, plane, reach, q):
    import torch
    import torch.distributed as dist
    from pd_mg_pin_corrosion_b200 import lib as L_
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = L_.load()
    lay = (C.c_longlong * 10)()
    L_.check(L.pdgpu_slab_layout(n_axial, plane, reach, world, rank, lay))
    a0, a1, nlp, NL, own_lo, own_hi, send_lo, recv_lo, send_hi, recv_hi = [int(v) for v in lay]
    hp = reach * plane
    glob = np.arange(n_axial * plane, dtype=np.float64) * 0.5 + 1.0       # the single-domain field
    loc = np.full(NL, -1.0)
    loc[own_lo:own_hi] = glob[a0 * plane:a1 * plane]                      # owned part only
    t = torch.from_numpy(loc)
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(t[send_lo:send_lo + hp].clone(), rank - 1))
        reqs.append(dist.irecv(t[recv_lo:recv_lo + hp], rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(t[send_hi:send_hi + hp].clone(), rank + 1))
        reqs.append(dist.irecv(t[recv_hi:recv_hi + hp], rank + 1))
    for r in reqs:
        r.wait()
    # expected: global planes [a0-reach, a1+reach) clipped to the domain, -1 padding outside
    exp = np.full(NL, -1.0)
    ga, gb = max(a0 - reach, 0), min(a1 + reach, n_axial)
    lo = (ga - (a0 - reach)) * plane
    exp[lo:lo + (gb - ga) * plane] = glob[ga * plane:gb * plane]
    ok = bool(np.array_equal(loc, exp))
    # scalar reductions of the convergence poll: sum / max over ranks
    s = torch.tensor([float(loc[own_lo:own_hi].sum()), float(loc[own_lo:own_hi].max())], dtype=torch.float64)
    ssum = s[:1].clone(); smax = s[1:].clone()
    dist.all_reduce(ssum, op=dist.ReduceOp.SUM)
    dist.all_reduce(smax, op=dist.ReduceOp.MAX)
    ok = ok and abs(ssum.item() - glob.sum()) <= 1e-9 * glob.sum() and smax.item() == glob.max()
    q.put((rank, ok, a0, a1))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_axial,plane,reach", [(37, 11, 3), (64, 157, 3)])
def test_two_rank_halo_exchange_gloo(n_axial, plane, reach):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_axial, plane, reach, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == n_axial


def _worker_balanced(rank, world, port, q):
    """what pdgpu_create_slab does: pdgpu_partition_balanced + pdgpu_slab_layout_range, then the same exchange"""
    import torch
    import torch.distributed as dist
    import helpers as H
    from pd_mg_pin_corrosion_b200 import lib as L_
    from pd_mg_pin_corrosion_b200.config import Config
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = L_.load()
    cfg = Config.load(os.path.join(H.CONFIG_DIR, "params.cfg"), {"use_implicit": 0}, quiet=True)
    s = cfg.to_struct()
    Nx, Ny, Nz = C.c_int(), C.c_int(), C.c_int()
    org = (C.c_double * 3)()
    L_.check(L.pdgpu_grid_extents(C.byref(s), 3, C.byref(Nx), C.byref(Ny), C.byref(Nz), org))
    n_axial, plane, reach = Nz.value, 7, cfg.m_ratio            # a thin stand-in plane keeps the arrays small
    a0c, a1c = C.c_int(), C.c_int()
    L_.check(L.pdgpu_partition_balanced(C.byref(s), 3, world, rank, C.byref(a0c), C.byref(a1c)))
    lay = (C.c_longlong * 10)()
    L_.check(L.pdgpu_slab_layout_range(a0c.value, a1c.value, plane, reach, lay))
    a0, a1, nlp, NL, own_lo, own_hi, send_lo, recv_lo, send_hi, recv_hi = [int(v) for v in lay]
    hp = reach * plane
    glob = np.arange(n_axial * plane, dtype=np.float64) * 0.25 - 3.0
    loc = np.full(NL, -1.0)
    loc[own_lo:own_hi] = glob[a0 * plane:a1 * plane]
    t = torch.from_numpy(loc)
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(t[send_lo:send_lo + hp].clone(), rank - 1))
        reqs.append(dist.irecv(t[recv_lo:recv_lo + hp], rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(t[send_hi:send_hi + hp].clone(), rank + 1))
        reqs.append(dist.irecv(t[recv_hi:recv_hi + hp], rank + 1))
    for r in reqs:
        r.wait()
    exp = np.full(NL, -1.0)
    ga, gb = max(a0 - reach, 0), min(a1 + reach, n_axial)
    lo = (ga - (a0 - reach)) * plane
    exp[lo:lo + (gb - ga) * plane] = glob[ga * plane:gb * plane]
    q.put((rank, bool(np.array_equal(loc, exp)), a0, a1, n_axial))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_balanced_slabs_halo_exchange_gloo(world):
    """cost-balanced slab boundaries: every rank derives its own range from the configuration alone, the ranges tile
    the axis, and the halo exchange over them reconstructs the single-domain ghost planes"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_balanced, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _, _ in res), res
    assert res[0][2] == 0 and res[-1][3] == res[0][4]
    for a, b in zip(res, res[1:]):
        assert a[3] == b[2]
