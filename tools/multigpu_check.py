#!/usr/bin/env python
"""Multi-GPU equivalence check: the z-slab run over N ranks must reproduce the
single-GPU fields BITWISE (same per-node arithmetic and order; only all-reduced scalars may
differ in the last bits).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_check.py [case] [ns_iters] [ard_steps] [dissolve_cycles] [host_chunks]

host_chunks > 0 additionally runs one pdgpu_step_host pass (host arrays in/out, chunked pipeline,
slab exchanges inside) on every rank's copy of the state and compares the owned parts.

Every rank runs its slab; rank 0 additionally runs the whole domain on its own GPU and
compares. Exit code 0 = identical.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from pd_mg_pin_corrosion_b200 import lib as L_, solver as S  # noqa: E402


def run(grid, cfg, iters, steps, dissolve_cycles, host_chunks=0):
    fields = S.Fields()
    fields.bind(grid)
    L = L_.load()
    from pd_mg_pin_corrosion_b200.grains import GrainStructure
    nt = np.zeros(grid.N_total, np.uint8)
    full = S.Grid(grid.dim, device=grid.device)   # whole-domain types for the host grain generator
    full.build(cfg)
    gs = GrainStructure().generate(full.node_type, cfg, grid.dim)
    full.close()
    L_.check(L.pdgpu_fields_init(grid.ctx, gs.is_grain_boundary.ctypes.data_as(C.c_void_p),
                                 gs.is_precipitate.ctypes.data_as(C.c_void_p)))
    ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
    ns.init(grid, cfg)
    ard.init(grid, cfg)
    dissolved = []
    for _ in range(max(1, dissolve_cycles)):
        dt = ns.compute_dt(fields, grid, cfg)
        ns.iterate(fields, grid, cfg, iters, dt)
        dtc = ard.compute_dt(fields, grid, cfg)
        ard.iterate(fields, grid, cfg, steps, dtc)
        if dissolve_cycles:
            n = ard.apply_phase_change(fields, grid, cfg)
            dissolved.append(ard.last_dissolved.copy())
    res = ns.residual(grid)
    out = {n: fields.get(n) for n in ("rho", "vel", "C")}
    if host_chunks:
        used, why = S.step_host_chunks(grid, host_chunks)
        print(f"[multigpu_check] rank {grid.rank}/{grid.nranks}: step_host chunks {used} {why}", flush=True)
        S.step_host(grid, dt, dtc, out["rho"], out["vel"], out["C"], host_chunks)
    out["node_type"] = grid.node_type
    out["scalars"] = np.array([dt, dtc, res.num, res.den, res.v_max])
    out["dissolved"] = np.concatenate(dissolved) if dissolved else np.zeros(0, np.int32)
    return out


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "3d_small"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    cycles = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    host_chunks = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = L_.load()
    dim, cfg, _ = H.load_cfg(case)
    grid = S.Grid(dim, device=local, rank=rank, nranks=world)
    grid.build(cfg)
    nb = L.pdgpu_comm_uid_bytes()
    uid = torch.zeros(nb, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * nb)()
        L_.check(L.pdgpu_comm_get_uid(buf))
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    L_.check(L.pdgpu_comm_init(grid.ctx, bytes(uid.cpu().tolist()), rank, world))
    for opt in os.environ.get("PDGPU_OPTIONS", "").split(","):   # e.g. PDGPU_OPTIONS=graph=2
        if "=" in opt:
            k, v = opt.split("=")
            grid.set_option(k.strip(), int(v))
    mine = run(grid, cfg, iters, steps, cycles, host_chunks)
    a0, a1, P = grid.a0, grid.a1, grid.plane
    # gather the owned parts on rank 0
    ok = True
    if rank == 0:
        full = S.Grid(dim, device=local)
        full.build(cfg)
        want = run(full, cfg, iters, steps, cycles, 1 if host_chunks else 0)   # unchunked single-GPU pass
        got = {n: np.array(mine[n], copy=True) for n in ("rho", "vel", "C", "node_type")}
    for n in ("rho", "vel", "C", "node_type"):
        for r in range(1, world):
            if rank == r:
                t = torch.from_numpy(np.ascontiguousarray(mine[n][a0 * P:a1 * P])).cuda()
                meta = torch.tensor([a0, a1], dtype=torch.int64).cuda()
                dist.send(meta, 0)
                dist.send(t.view(torch.uint8).flatten(), 0)
            elif rank == 0:
                meta = torch.zeros(2, dtype=torch.int64).cuda()
                dist.recv(meta, r)
                b0, b1 = [int(v) for v in meta.cpu()]
                tmpl = got[n][b0 * P:b1 * P]
                t = torch.zeros(tmpl.nbytes, dtype=torch.uint8).cuda()
                dist.recv(t, r)
                got[n][b0 * P:b1 * P] = np.frombuffer(t.cpu().numpy().tobytes(), dtype=tmpl.dtype).reshape(tmpl.shape)
    if rank == 0:
        for n in ("node_type", "rho", "vel", "C"):
            same = got[n].tobytes() == want[n].tobytes()
            err = H.rel_err(got[n].astype(np.float64), want[n].astype(np.float64))
            print(f"[multigpu_check] {case} N={world} {n:9s} bitwise={'yes' if same else 'NO'} rel_err={err:.2e}")
            ok = ok and same
        sc = np.abs(mine["scalars"] - want["scalars"]) / np.maximum(np.abs(want["scalars"]), 1e-300)
        print(f"[multigpu_check] scalars (dt_ns, dt_ard, num, den, vmax) rel diff: {sc}")
        ok = ok and bool((sc <= 1e-12).all())
        if cycles:
            print(f"[multigpu_check] dissolved (rank-0 slab) {mine['dissolved'].size} vs whole {want['dissolved'].size}")
    flag = torch.tensor([1 if ok else 0]).cuda()
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
