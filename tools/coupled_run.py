#!/usr/bin/env python
"""CoupledSolver::run (src/coupling.cpp:82-302, explicit branch) over z-slabs: one rank per GPU.

    python tools/coupled_run.py <case> <out_dir>                                    # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/coupled_run.py <case> <out_dir>                       # N slabs

Every rank runs the coupling loop on its slab; convergence polls, dt, diagnostics and the
dissolved-node counts are reduced over the ranks inside the library, the ordered sum of C over the
initial solid nodes (src/coupling.cpp:32-38) is formed identically on every rank from a collective
gather, rank 0 writes diagnostics.csv / mass_loss.csv.  Rank 0 finally prints a SHA-256 of the
global rho / vel / C arrays, so that runs at different N can be compared bit for bit.
`case` is a name from tests/helpers.py CASES or a .cfg path (then --dim 2|3)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    case, out_dir = args[0], args[1]
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    from pd_mg_pin_corrosion_b200 import solver as S
    from pd_mg_pin_corrosion_b200.config import Config
    from pd_mg_pin_corrosion_b200.grains import GrainStructure
    if case.endswith(".cfg"):
        dim = int(sys.argv[sys.argv.index("--dim") + 1]) if "--dim" in sys.argv else 3
        cfg = Config.load(case, {"use_implicit": 0, "output_dir": out_dir}, quiet=True)
    else:
        import helpers as H
        dim, cfg, _ = H.load_cfg(case, {"output_dir": out_dir})
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid = S.Grid(dim, device=local, rank=rank, nranks=world)
    grid.build(cfg)
    if world > 1:
        grid.comm_init_torch()
    for opt in os.environ.get("PDGPU_OPTIONS", "").split(","):
        if "=" in opt:
            k, v = opt.split("=")
            grid.set_option(k.strip(), int(v))
    grains = GrainStructure().generate(grid.node_type_all, cfg, dim)
    fields = S.Fields()
    fields.allocate(grid.N_total, grid)
    S.initialize_fields(fields, grid, grains, cfg)
    cs = S.CoupledSolver()
    if rank != 0 or "--quiet" in sys.argv:
        cs.log = lambda *a, **k: None
    t_end = cs.run(grid, fields, cfg)
    h = hashlib.sha256()
    nt = grid.node_type_all
    parts = []
    for name in ("rho", "vel", "C"):
        a = np.ascontiguousarray(fields.get_all(name))
        h.update(a.tobytes())
        if rank == 0 and os.environ.get("PD_DUMP_DIR"):
            np.save(os.path.join(os.environ["PD_DUMP_DIR"], f"{name}_n{world}.npy"), a)
            np.save(os.path.join(os.environ["PD_DUMP_DIR"], f"type_n{world}.npy"), nt)
        per_type = [hashlib.sha256(np.ascontiguousarray(a[nt == t]).tobytes()).hexdigest()[:8] for t in range(6)]
        parts.append(f"{name}:" + "/".join(per_type))
    if rank == 0 and "--quiet" not in sys.argv or os.environ.get("PD_HASH_PARTS"):
        if rank == 0:
            print("[coupled_run] per-type hashes (FLUID/SOLID/WALL/INLET/OUTLET/OUTSIDE) " + " ".join(parts), flush=True)
    if rank == 0:
        print(f"[coupled_run] case={case} ranks={world} t_end={t_end:.9e} dissolved={cs.total_dissolved} "
              f"solid={int((nt == 1).sum())} fields_sha256={h.hexdigest()}", flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
