import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import helpers as H
from test_gpu_parity import gpu_side
case, nch = sys.argv[1], int(sys.argv[2])
ref = H.make_ref(case)
S, cfg, grid, fields = gpu_side(case, None, ref=ref, upload=False)
ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
ns.init(grid, cfg); ard.init(grid, cfg)
dt = ref.ns_compute_dt()
ns.iterate(fields, grid, cfg, 12, dt)
dtc = ard.compute_dt(fields, grid, cfg)
ard.iterate(fields, grid, cfg, 3, dtc)
state0 = {n: fields.get(n) for n in ("rho", "vel", "C")}
print(S.step_host_chunks(grid, nch))
outs = []
for nc in (1, nch):
    st = {n: state0[n].copy() for n in state0}
    S.step_host(grid, dt, dtc, st["rho"], st["vel"], st["C"], nc)
    outs.append(st)
nt = grid.node_type
P = grid.Nx * grid.Ny
for n in ("rho", "vel", "C"):
    a, b = outs[0][n], outs[1][n]
    if a.ndim == 2: a = a[:, 2]; b = b[:, 2]
    bad = np.nonzero(a != b)[0]
    print(n, "mismatch", bad.size)
    if bad.size:
        print(" planes", np.unique(bad // P), "types", np.bincount(nt[bad], minlength=6))
        for i in bad[:10]:
            print("  ", i, i // P, (i % P) // grid.Nx, i % grid.Nx, nt[i], a[i], b[i])
