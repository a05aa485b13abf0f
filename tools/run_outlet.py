"""Runs apply_outlet_bc a few times on the bench workload (profiling target for the sweep kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pd_mg_pin_corrosion_b200 import lib as L_, solver as S   # noqa: E402
from pd_mg_pin_corrosion_b200.config import Config            # noqa: E402

cfg = Config.load(os.path.join(ROOT, "configs", "params_fine.cfg"), {"use_implicit": 0}, quiet=True)
L = L_.load()
grid = S.Grid(3)
grid.build(cfg)
fields = S.Fields(); fields.bind(grid)
L_.check(L.pdgpu_fields_init(grid.ctx, None, None))
for _ in range(3):
    L_.check(L.pdgpu_bc_outlet(grid.ctx))
print("ok")
