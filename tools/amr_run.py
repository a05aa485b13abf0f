"""Whole coupled run on the two-level AMR grid (explicit or implicit ARD branch, as the config says), standalone: config -> Grid::build_amr ->
cell-list neighbours -> grains -> initialize_fields -> CoupledSolver::run; writes <output_dir>/diagnostics.csv and the
state_/flow_/corr_/final_ VTU series with simulation.pvd / flow.pvd (as the reference does; --no-vti: none).
usage: python tools/amr_run.py configs/params_amr.cfg [key=value ...] [--no-vti]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pd_mg_pin_corrosion_b200 import amr as A            # noqa: E402
from pd_mg_pin_corrosion_b200.config import Config       # noqa: E402

path = sys.argv[1]
ov = {}                                                   # use_implicit as the file says (explicit: use_implicit=0 on the command line)
for kv in [a for a in sys.argv[2:] if a != "--no-vti"]:
    k, v = kv.split("=", 1)
    ov[k] = type(getattr(Config(), k))(float(v)) if not isinstance(getattr(Config(), k), str) else v
cfg = Config.load(path, ov, quiet=False)
t0 = time.perf_counter()
g = A.AmrGrid(cfg)
g.build_amr()
g.build_neighbors_celllist()
i = g.info
print(f"AMR total: {i.N_total} nodes (fine={i.n_fine}, coarse={i.n_coarse}, fict={i.n_fict}); {i.nnz} neighbour entries")
gid, gb, pr, n = A.generate_grains(g)
print(f"Grain generation: {n} grains, {int(gb.sum())} boundary nodes, {int(pr.sum())} precipitate nodes")
g.device_init(0)
A.initialize_fields(g, gb, pr)
rows = A.AmrCoupledSolver(log=print).run(g, cfg.output_dir, grain_id=None if "--no-vti" in sys.argv else gid)   # VTU series + PVD
print(f"{len(rows)} diagnostics rows -> {os.path.join(cfg.output_dir, 'diagnostics.csv')}; total {time.perf_counter() - t0:.2f} s")
