"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path through the C ABI
(libpdgpu.so via pd_mg_pin_corrosion_b200.solver) against the oracle -- oracle/_ref (the
unmodified reference, compiled from its sources) when it was built, else the plain-C port.

Bars (BASELINE.json north_star / SURVEY.md 7.3):
  bit-exact : node classification, CSR neighbour list, wall-mirror table, dissolved sets
  1e-12 rel : per-step fields (rho, u, C), max-abs error over max-abs value
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

TOL = 1e-12   # north_star: per-step fields within 1e-12 relative max-error in FP64


def gpu_side(case, extra=None, ref=None, upload=True):
    from pd_mg_pin_corrosion_b200 import solver as S
    dim, cfg, _ = H.load_cfg(case, extra)
    grid = S.Grid(dim)
    grid.build(cfg)
    fields = S.Fields()
    fields.allocate(grid.N_total, grid)

    class G:   # grains as produced by the host generator (here: taken from the oracle)
        is_grain_boundary = ref.get("is_gb") if ref is not None else np.zeros(grid.N_total, np.uint8)
        is_precipitate = ref.get("is_precip") if ref is not None else np.zeros(grid.N_total, np.uint8)
        grain_id = np.full(grid.N_total, -1, np.int32)

    S.initialize_fields(fields, grid, G, cfg)
    if ref is not None and upload:
        for n in ("rho", "vel", "C", "rho_new", "vel_new", "C_new", "phase"):
            fields.set(n, ref.get(n))
    return S, cfg, grid, fields


def assert_fields(fields, ref, names, tol=TOL, where=None):
    for n in names:
        a, b = fields.get(n), ref.get(n)
        if where is not None:
            a, b = a[where], b[where]
        e = H.rel_err(a, b)
        assert e <= tol, f"{n}: rel err {e:.3e} > {tol:.1e}"


GEOM_CASES = ["2d_default", "2d_poiseuille", "2d_offgrid", "3d_small", "3d_offgrid", "3d_default"]


@pytest.mark.parametrize("case", GEOM_CASES)
def test_classification_and_tables_bit_exact(case):
    ref = H.make_ref(case)
    S, cfg, grid, fields = gpu_side(case, ref=None)
    assert (grid.Nx, grid.Ny, grid.Nz, grid.N_total) == (ref.Nx, ref.Ny, ref.Nz, ref.N)
    nt_ref = ref.get("node_type")
    assert np.array_equal(grid.node_type, nt_ref)
    assert [int(c) for c in grid.info.counts] == np.bincount(nt_ref, minlength=6).tolist()
    # wall-mirror table: index trick on the reference, direct download here
    N = ref.N
    ref.set("rho", np.arange(N) + 0.25)
    ref.wall_bc()
    rr = ref.get("rho")
    w = nt_ref == 2
    mir_ref = np.where(np.modf(rr[w])[0] == 0.25, (rr[w] - 0.25).astype(np.int64), -1)
    assert np.array_equal(grid.wall_mirror[w], mir_ref)
    assert np.all(grid.wall_mirror[~w] == -1)


@pytest.mark.parametrize("case", ["2d_default", "2d_offgrid", "3d_small", "3d_default"])
def test_csr_bit_exact(case):
    ref = H.make_ref(case)
    S, cfg, grid, fields = gpu_side(case, ref=None)
    nnz = grid.build_neighbors()
    off, idx, dist, evec, vol = grid.csr()
    if hasattr(ref, "csr"):
        roff, ridx, rdist, revec, rvol = ref.csr()
    else:
        roff, ridx, rdist, revec, rvol = (ref.get(n) for n in ("nbr_offset", "nbr_index", "nbr_dist", "nbr_evec", "nbr_vol"))
    assert nnz == int(roff[-1]) == grid.info.nnz
    assert np.array_equal(off, roff.astype(np.int64))
    assert np.array_equal(idx, ridx)
    assert dist.tobytes() == rdist.tobytes()
    assert evec.tobytes() == revec.tobytes()
    assert vol.tobytes() == rvol.tobytes()
    # bond-update counts of the metric
    nt = ref.get("node_type")
    rowlen = np.diff(roff.astype(np.int64))
    assert grid.info.ns_bonds == int(rowlen[nt == 0].sum())
    assert grid.info.ard_bonds == int(rowlen[(nt == 0) | (nt == 1)].sum())
    grid.free_neighbors()


@pytest.mark.parametrize("case", ["2d_default", "2d_poiseuille", "2d_offgrid", "3d_small", "3d_default"])
def test_boundary_operators(case):
    ref = H.make_ref(case)
    H.perturbed_state(ref, seed=1)
    S, cfg, grid, fields = gpu_side(case, ref=ref)
    ops = [("inlet_bc", lambda: S.apply_inlet_bc(fields, grid, cfg)),
           ("outlet_bc", lambda: S.apply_outlet_bc(fields, grid, cfg)),
           ("wall_bc", lambda: S.apply_wall_bc(fields, grid, cfg)),
           ("solid_bc", lambda: S.apply_solid_surface_bc(fields, grid)),
           ("wall_conc_bc", lambda: S.apply_wall_concentration_bc(fields, grid, cfg)),
           ("wall_bc_new", lambda: S.apply_wall_bc_new(fields, grid, cfg))]
    for name, gpu_op in ops:
        getattr(ref, name)()
        gpu_op()
        assert_fields(fields, ref, ("rho", "vel", "C", "rho_new", "vel_new"))


@pytest.mark.parametrize("case,extra", [("3d_default", None), ("2d_default", {"gb_width_cells": 2, "precip_cluster_cells": 2}),
                                        ("3d_small", {"gb_width_cells": 1, "precip_cluster_cells": 1, "precip_fraction": 0.2}),
                                        ("2d_offgrid", None), ("3d_offgrid", {"gb_width_cells": 2})])
def test_device_grain_generation_bit_exact(case, extra):
    """GrainStructure::generate with the Voronoi / boundary / dilation / cluster passes on the device
    (src/grains.cpp:55-107,152-166) against the reference: grain ids and both flag arrays
    bit for bit (nearest-seed ties resolve like the reference's strict `<`)."""
    from pd_mg_pin_corrosion_b200.grains import GrainStructure
    ref = H.make_ref(case, extra)
    dim, cfg, _ = H.load_cfg(case, extra)
    S, cfg2, grid, fields = gpu_side(case, extra, ref=ref, upload=False)
    g = GrainStructure().generate(ref.get("node_type"), cfg, dim, grid=grid)
    h = GrainStructure().generate(ref.get("node_type"), cfg, dim)          # host path
    for name, want in (("grain_id", ref.get("grain_id")), ("is_grain_boundary", ref.get("is_gb")),
                       ("is_precipitate", ref.get("is_precip"))):
        assert np.array_equal(getattr(g, name), want), name
        assert np.array_equal(getattr(h, name), want), name
    assert g.n_grains == h.n_grains and int(g.is_grain_boundary.sum()) > 0


@pytest.mark.parametrize("case", ["2d_default", "2d_offgrid", "3d_small", "3d_default"])
def test_smooth_boundary_concentration_bit_exact(case):
    """smooth_boundary_concentration (src/boundary.cpp:332-376): in-place sweep in the
    reference's index order; sums run in CSR order, so C must be reproduced bit for bit. Applied twice:
    the second pass reads the first pass's output (order dependence at the outlet end)."""
    ref = H.make_ref(case)
    H.perturbed_state(ref, seed=7)
    S, cfg, grid, fields = gpu_side(case, ref=ref)
    before = fields.get("C")
    for _ in range(2):
        ref.smooth_conc()
        S.smooth_boundary_concentration(fields, grid, cfg)
        got, want = fields.get("C"), ref.get("C")
        assert np.array_equal(got, want), float(np.abs(got - want).max())
    assert not np.array_equal(before, fields.get("C"))   # the operator did something


@pytest.mark.parametrize("case", ["2d_default", "2d_poiseuille", "2d_offgrid", "3d_small", "3d_default"])
def test_ns_step_and_dt(case):
    ref = H.make_ref(case)
    H.perturbed_state(ref, seed=2)
    S, cfg, grid, fields = gpu_side(case, ref=ref)
    ns = S.PD_NS_Solver()
    ns.init(grid, cfg)
    dt_ref = ref.ns_compute_dt()
    dt = ns.compute_dt(fields, grid, cfg)
    assert abs(dt - dt_ref) <= 1e-15 * dt_ref
    ref.ns_step(dt_ref)
    ns.step(fields, grid, cfg, dt_ref)
    assert_fields(fields, ref, ("rho_new", "vel_new"))
    nt = ref.get("node_type")
    assert_fields(fields, ref, ("pressure",), where=nt != 5)


@pytest.mark.parametrize("case,iters", [("2d_default", 200), ("2d_poiseuille", 200), ("3d_small", 40)])
def test_ns_iterate(case, iters):
    """K full loop bodies (BCs + step + wall_new + swap) from the initial state."""
    ref = H.make_ref(case)
    S, cfg, grid, fields = gpu_side(case, ref=ref)
    ns = S.PD_NS_Solver()
    ns.init(grid, cfg)
    dt = ref.ns_compute_dt()
    ref.ns_iterate(iters, dt)
    ns.iterate(fields, grid, cfg, iters, dt)
    assert_fields(fields, ref, ("rho", "vel", "C", "rho_new", "vel_new"))
    # convergence-block scalars (src/pd_ns.cpp:273-301) of one more un-swapped step
    for bc in ("inlet_bc", "outlet_bc", "wall_bc", "solid_bc"):
        getattr(ref, bc)()
    S.apply_inlet_bc(fields, grid, cfg); S.apply_outlet_bc(fields, grid, cfg)
    S.apply_wall_bc(fields, grid, cfg); S.apply_solid_surface_bc(fields, grid)
    ref.ns_step(dt); ref.wall_bc_new()
    ns.step(fields, grid, cfg, dt); S.apply_wall_bc_new(fields, grid, cfg)
    r = ns.residual(grid)
    nt = ref.get("node_type")
    fl = nt == 0
    v, vn, rn = ref.get("vel")[fl], ref.get("vel_new")[fl], ref.get("rho_new")[fl]
    num, den = float(((vn - v) ** 2).sum()), float((v ** 2).sum())
    assert abs(r.num - num) <= 1e-9 * num and abs(r.den - den) <= 1e-12 * den
    assert abs(r.v_max - np.sqrt((vn ** 2).sum(1)).max()) <= 1e-12 * r.v_max
    assert abs(r.rho_min - rn.min()) <= 1e-12 * rn.min() and abs(r.rho_max - rn.max()) <= 1e-12 * rn.max()
    assert r.has_nan == 0


@pytest.mark.parametrize("case", ["2d_default", "2d_dissolve", "3d_small", "3d_default"])
def test_ard_step_and_dt(case):
    ref = H.make_ref(case)
    H.perturbed_state(ref, seed=3)
    # exercise the salt layer: saturate some fluid next to the wire
    C = ref.get("C")
    nt = ref.get("node_type")
    rng = np.random.default_rng(5)
    C[(nt == 0) & (rng.random(C.size) < 0.02)] = 0.95
    ref.set("C", C)
    S, cfg, grid, fields = gpu_side(case, ref=ref)
    ard = S.PD_ARD_Solver()
    ard.init(grid, cfg)
    dt_ref = ref.ard_compute_dt()
    dt = ard.compute_dt(fields, grid, cfg)
    assert abs(dt - dt_ref) <= 1e-15 * dt_ref
    ref.ard_step(dt_ref)
    ard.step(fields, grid, cfg, dt_ref)
    assert_fields(fields, ref, ("C_new",))


@pytest.mark.parametrize("case,flow_iters,steps", [("2d_default", 300, 100), ("3d_small", 50, 20)])
def test_ard_iterate(case, flow_iters, steps):
    ref = H.make_ref(case)
    S, cfg, grid, fields = gpu_side(case, ref=ref)
    ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
    ns.init(grid, cfg); ard.init(grid, cfg)
    dt = ref.ns_compute_dt()
    ref.ns_iterate(flow_iters, dt)
    ns.iterate(fields, grid, cfg, flow_iters, dt)
    dtc = ref.ard_compute_dt()
    assert abs(ard.compute_dt(fields, grid, cfg) - dtc) <= 1e-12 * dtc
    ref.ard_iterate(steps, dtc)
    ard.iterate(fields, grid, cfg, steps, dtc)
    assert_fields(fields, ref, ("rho", "vel", "C"))


def test_phase_change_sets_bit_exact():
    """Coupling cycles of the explicit branch with dissolution: the set of
    dissolved nodes per check must be identical, fields within tolerance."""
    case = "2d_dissolve"
    ref = H.make_ref(case)
    S, cfg, grid, fields = gpu_side(case, ref=ref)
    ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
    ns.init(grid, cfg); ard.init(grid, cfg)
    dt = ref.ns_compute_dt()
    total = 0
    for cycle in range(3):
        ref.ns_iterate(300, dt)
        ns.iterate(fields, grid, cfg, 300, dt)
        dtc = ref.ard_compute_dt()
        ref.ard_iterate(50, dtc)
        ard.iterate(fields, grid, cfg, 50, dtc)
        assert_fields(fields, ref, ("C",), tol=1e-11)
        before = ref.get("node_type")
        n_ref = ref.phase_change()
        after = ref.get("node_type")
        dissolved_ref = np.nonzero(before != after)[0]
        n_gpu = ard.apply_phase_change(fields, grid, cfg)
        assert n_gpu == n_ref == dissolved_ref.size
        assert np.array_equal(ard.last_dissolved, dissolved_ref)
        assert np.array_equal(grid.node_type, after)
        assert np.array_equal(fields.get("phase"), ref.get("phase"))
        ref.rebuild_neighbors()
        assert_fields(fields, ref, ("rho", "vel", "C"), tol=1e-11)
        total += n_ref
    assert total > 0, "synthetic config must dissolve nodes"
    d = S.diagnostics(grid)
    assert d.solid_count == int((ref.get("node_type") == 1).sum())


def test_errors_are_loud():
    from pd_mg_pin_corrosion_b200 import lib as L, solver as S
    dim, cfg, _ = H.load_cfg("2d_default", {"use_amr": 1})
    with pytest.raises(ValueError, match="use_amr"):
        S.Grid(2).build(cfg)
    cfg.use_amr = 0
    s = cfg.to_struct()
    import ctypes as C
    ctx = C.c_void_p()
    assert L.load().pdgpu_create(C.byref(s), 4, 0, C.byref(ctx)) != 0          # dim must be 2 or 3
    assert b"dim" in L.load().pdgpu_last_error()


@pytest.mark.parametrize("case", ["3d_small", "3d_default"])
def test_kernel_variants_agree(case):
    """Every fast path (tiled bond kernels, ring-buffer outlet sweeps, side-stream overlap, CUDA
    graphs) against the generic one-thread-per-node kernels and against the oracle."""
    ref = H.make_ref(case)
    iters, steps = (30, 10) if case == "3d_small" else (6, 3)
    variants = [
        dict(ns_kernel=0, ard_kernel=0, outlet_kernel=0, overlap=0, graph=0),   # all generic
        dict(ns_kernel=1, ard_kernel=1, outlet_kernel=1, overlap=0, graph=0),
        dict(ns_kernel=1, ard_kernel=1, outlet_kernel=2, overlap=0, graph=1),
        dict(ns_kernel=1, ard_kernel=1, outlet_kernel=2, overlap=1, graph=0, lazy_wallc=0),
        dict(ns_kernel=1, ard_kernel=1, outlet_kernel=2, overlap=1, graph=1),
        dict(ns_kernel=1, ard_kernel=1, outlet_kernel=3, overlap=1, graph=1),   # block-tile kernels
        dict(ns_kernel=1, ard_kernel=1, outlet_kernel=3, overlap=0, graph=0),
        dict(ns_kernel=1, ard_kernel=1, outlet_kernel=3, overlap=1, graph=1, outlet_single_rows=1),
        dict(ns_kernel=0, ard_kernel=0, outlet_kernel=2, overlap=1, graph=1),
        dict(ns_kernel=2, ard_kernel=1, outlet_kernel=3, overlap=1, graph=1),   # z-streaming NS kernel (default)
        dict(ns_kernel=2, ard_kernel=1, outlet_kernel=2, overlap=0, graph=0),
        dict(ns_kernel=2, ard_kernel=1, outlet_kernel=3, overlap=1, graph=1, stream_chunk=8),
        dict(ns_kernel=2, ard_kernel=1, outlet_kernel=3, overlap=0, graph=0, stream_chunk=12),
        dict(ns_kernel=3, ard_kernel=3, outlet_kernel=2, overlap=1, graph=1),   # materialised-CSR path
    ]
    dt = ref.ns_compute_dt()
    results = []
    for opts in variants:
        S, cfg, grid, fields = gpu_side(case, ref=ref, upload=False)
        for k, v in opts.items():
            grid.set_option(k, v)
        if opts["ns_kernel"] == 3:
            grid.build_neighbors()
        ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
        ns.init(grid, cfg); ard.init(grid, cfg)
        ns.iterate(fields, grid, cfg, iters, dt)
        dtc = ard.compute_dt(fields, grid, cfg)
        ard.iterate(fields, grid, cfg, steps, dtc)
        results.append({n: fields.get(n) for n in ("rho", "vel", "C")})
        grid.close()
    ref.ns_iterate(iters, dt)
    ref.ard_iterate(steps, ref.ard_compute_dt())
    for opts, res in zip(variants, results):
        for n in ("rho", "vel", "C"):
            assert H.rel_err(res[n], results[0][n]) <= 1e-13, (opts, n, "vs generic")
            assert H.rel_err(res[n], ref.get(n)) <= TOL, (opts, n, "vs oracle")


@pytest.mark.parametrize("case,extra", [
    ("2d_default", None), ("2d_offgrid", None), ("2d_poiseuille", {"channel_flow_corrections": 1}),
    ("2d_default", {"m_ratio": 2}), ("2d_default", {"m_ratio": 4}), ("2d_dissolve", None)])
def test_2d_persistent_flow_loop(case, extra):
    """2D: batches of NS loop bodies as one persistent cooperative kernel (csrc/ns2d.cu: BCs, Gauss-Seidel
    outlet recurrence, wall mirror folded into the staging, bond sums, channel corrections, two grid barriers
    per iteration) against one launch per operator (ns2d = 0) and against the oracle -- after 1, 2 and 57 loop
    bodies from a perturbed state (both buffers: the state a later swap exposes must match too)."""
    ref = H.make_ref(case, extra)
    H.perturbed_state(ref, seed=7)
    dt = ref.ns_compute_dt()
    names = ("rho", "vel", "C", "rho_new", "vel_new")
    done = 0
    sides = []
    for ns2d in (1, 0):
        S, cfg, grid, fields = gpu_side(case, extra, ref=ref, upload=True)
        grid.set_option("ns2d", ns2d)
        ns = S.PD_NS_Solver(); ns.init(grid, cfg)
        sides.append((S, cfg, grid, fields, ns))
    for iters in (1, 2, 57):
        out = []
        for S, cfg, grid, fields, ns in sides:
            n0 = grid.launch_count()
            ns.iterate(fields, grid, cfg, iters, dt)
            out.append(({n: fields.get(n) for n in names}, grid.launch_count() - n0))
        ref.ns_iterate(iters, dt)
        done += iters
        assert out[0][1] <= 3 and out[1][1] >= 5 * iters, (out[0][1], out[1][1])   # dt upload + one kernel per batch
        for n in names:
            assert H.rel_err(out[0][0][n], out[1][0][n]) <= 1e-13, (case, done, n, "persistent vs per-operator")
        if not (extra and "channel_flow_corrections" in extra):   # (the oracle applies those in solve_steady only)
            for n in ("rho", "vel", "C"):
                assert H.rel_err(out[0][0][n], ref.get(n)) <= TOL, (case, done, n, "vs oracle")
    # the corrosion loop body (src/coupling.cpp:232-240) through k_ard2d_loop: 1, 2 and 31 steps
    dtc = ref.ard_compute_dt()
    ards = []
    for S, cfg, grid, fields, ns in sides:
        ard = S.PD_ARD_Solver(); ard.init(grid, cfg)
        ards.append(ard)
    for steps in (1, 2, 31):
        out = []
        for (S, cfg, grid, fields, ns), ard in zip(sides, ards):
            n0 = grid.launch_count()
            ard.iterate(fields, grid, cfg, steps, dtc)
            out.append(({n: fields.get(n) for n in ("rho", "vel", "C", "C_new")}, grid.launch_count() - n0))
        ref.ard_iterate(steps, dtc)
        assert out[0][1] <= 6 and out[1][1] >= 4 * steps, (out[0][1], out[1][1])   # dt, |v| table, one kernel per batch
        for n in ("rho", "vel", "C", "C_new"):
            assert H.rel_err(out[0][0][n], out[1][0][n]) <= 1e-13, (case, steps, n, "persistent vs per-operator")
        if not (extra and "channel_flow_corrections" in extra):
            for n in ("rho", "vel", "C"):
                assert H.rel_err(out[0][0][n], ref.get(n)) <= TOL, (case, steps, n, "vs oracle")
    for _, _, grid, _, _ in sides:
        grid.close()


@pytest.mark.parametrize("case", ["2d_default", "2d_poiseuille"])
def test_2d_solve_steady_persistent(case):
    """solve_steady through the persistent kernel: same iteration count, same residual and the same fields as
    the per-operator path (flow_max_iters capped so the test stays short)."""
    extra = {"flow_max_iters": 700}
    res = []
    for ns2d in (1, 0):
        S, cfg, grid, fields = gpu_side(case, extra, ref=None)
        grid.set_option("ns2d", ns2d)
        ns = S.PD_NS_Solver(); ns.init(grid, cfg)
        it = ns.solve_steady(fields, grid, cfg, verbose=False)
        res.append((it, ns.last.eps, ns.last.status, {n: fields.get(n) for n in ("rho", "vel", "C", "rho_new", "vel_new")}))
        grid.close()
    assert res[0][0] == res[1][0] and res[0][2] == res[1][2]
    assert abs(res[0][1] - res[1][1]) <= 1e-9 * abs(res[1][1])
    for n, a in res[0][3].items():
        assert H.rel_err(a, res[1][3][n]) <= 1e-12, n


@pytest.mark.parametrize("case", ["3d_small", "2d_default"])
def test_step_iterate_equals_alternating_calls(case):
    """pdgpu_step_iterate(n) (no host synchronisation between the bodies) leaves exactly the state of
    n x { pdgpu_ns_iterate(1) ; pdgpu_ard_iterate(1) }."""
    from pd_mg_pin_corrosion_b200 import lib as L_
    import ctypes as C
    ref = H.make_ref(case)
    dt, dtc = ref.ns_compute_dt(), ref.ard_compute_dt()
    out = []
    for fused in (True, False):
        S, cfg, grid, fields = gpu_side(case, ref=ref, upload=False)
        if fused:
            L_.check(L_.load().pdgpu_step_iterate(grid.ctx, 7, dt, dtc))
        else:
            for _ in range(7):
                L_.check(L_.load().pdgpu_ns_iterate(grid.ctx, 1, dt))
                L_.check(L_.load().pdgpu_ard_iterate(grid.ctx, 1, dtc))
        out.append({n: fields.get(n) for n in ("rho", "vel", "C", "rho_new", "vel_new", "C_new")})
        grid.close()
    for n, a in out[0].items():
        assert np.array_equal(a, out[1][n]), (case, n)


@pytest.mark.parametrize("case,extra,n_chunks,expect", [
    ("3d_small", None, 3, 2), ("3d_default", None, 8, 8), ("3d_default", None, 16, 12),
    ("2d_default", None, 6, 4), ("2d_dissolve", None, 4, 3),
    ("2d_poiseuille", {"channel_flow_corrections": 1}, 4, 1)])   # falls back to one chunk
def test_step_host_chunked_is_bit_identical(case, extra, n_chunks, expect):
    """pdgpu_step_host (host arrays in/out, chunked upload/compute/download pipeline) against
    upload + ns_iterate(1) + ard_iterate(1) + download, bit for bit, and against the oracle."""
    ref = H.make_ref(case, extra)
    S, cfg, grid, fields = gpu_side(case, extra, ref=ref, upload=False)
    # the chunked pipeline runs the per-operator kernels; the 2D persistent loop (csrc/ns2d.cu) sums in another
    # order (1e-13, test_2d_persistent_flow_loop), so the bit-for-bit comparison is made against the former
    grid.set_option("ns2d", 0)
    ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
    ns.init(grid, cfg); ard.init(grid, cfg)
    dt = ref.ns_compute_dt()
    ns.iterate(fields, grid, cfg, 12, dt)            # some flow so that every term is exercised
    dtc = ard.compute_dt(fields, grid, cfg)
    ard.iterate(fields, grid, cfg, 3, dtc)
    state0 = {n: fields.get(n) for n in ("rho", "vel", "C")}
    used, why = S.step_host_chunks(grid, n_chunks)
    assert (used == 1 if expect == 1 else used >= expect), (used, why)
    outs = []
    for nc in (1, n_chunks):
        st = {n: state0[n].copy() for n in state0}
        for _ in range(3):                           # buffer parity flips between calls
            S.step_host(grid, dt, dtc, st["rho"], st["vel"], st["C"], nc)
        outs.append(st)
        # the device state after the call is the downloaded one
        for n in ("rho", "vel", "C"):
            assert np.array_equal(fields.get(n), st[n]), (nc, n)
    for n in ("rho", "vel", "C"):
        assert np.array_equal(outs[0][n], outs[1][n]), n
    grid.close()
    if extra:      # (the shim's ns_iterate has no channel corrections: solve_steady-only in the reference)
        return
    # and the unchunked call is the reference's sequence
    for n in ("rho", "vel", "C"):
        ref.set(n, state0[n])
    for _ in range(3):
        ref.ns_iterate(1, dt)
        ref.ard_iterate(1, dtc)
    for n in ("rho", "vel", "C"):
        assert H.rel_err(outs[1][n], ref.get(n)) <= TOL, n
    # OUTLET concentrations are many orders below C_solid: compare them on their own scale (they see
    # the outlet sweeps of BOTH loop bodies, which must hit the current C buffer whatever its parity)
    outlet = ref.get("node_type") == 4
    for st in outs:
        assert H.rel_err(st["C"][outlet], ref.get("C")[outlet]) <= 1e-11


@pytest.mark.parametrize("case", ["2d_dissolve", "3d_dissolve"])
def test_host_driver_whole_run_diagnostics(case, tmp_path):
    """host/pd_corrosion_gpu (C++17 driver over the C ABI) on the dissolving synthetic configs:
    every numeric column of diagnostics.csv within 1e-6 relative of the reference's own main()
    (tests/golden/diagnostics_<case>.csv); solid-node counts exact (the 3D case dissolves 61 nodes in two
    checks and re-solves the flow in between)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "host", "pd_corrosion_gpu")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "host")])
    dim, base, ov = H.CASES[case]
    ov = dict(ov, use_implicit=0, output_dir=str(tmp_path / "out"))
    from oracle import refapi
    cfg_path = refapi.write_cfg(base, ov, str(tmp_path / "run.cfg"))
    r = subprocess.run([exe, cfg_path, "--dim", str(dim), "--no-vti"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = np.loadtxt(str(tmp_path / "out" / "diagnostics.csv"), delimiter=",", skiprows=1)
    gold = np.loadtxt(os.path.join(root, "tests", "golden", f"diagnostics_{case}.csv"), delimiter=",", skiprows=1)
    assert got.shape == gold.shape
    assert np.array_equal(got[:, 3], gold[:, 3])
    for col in (0, 1, 2, 4, 5):
        rel = np.abs(got[:, col] - gold[:, col]) / np.maximum(np.abs(gold[:, col]), 1e-300)
        assert rel.max() <= 1e-6, (col, float(rel.max()))


def test_host_driver_output_files_match_reference(tmp_path):
    """Same run as above through the reference's own main() (oracle/_ref): identical set of
    state_/flow_/corr_/final_ VTI names, byte-identical PVD collections, and VTI snapshots that
    agree line by line (6 printed digits; fields agree to 1e-12, so a last printed digit may flip)."""
    import glob
    import os
    import subprocess
    from oracle import refapi
    if not refapi.have_ref(2):
        pytest.skip("oracle/_ref not built")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "host", "pd_corrosion_gpu")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "host")])
    dim, base, ov = H.CASES["2d_dissolve"]
    outs = {}
    for who in ("ref", "gpu"):
        o = dict(ov, use_implicit=0, output_dir=str(tmp_path / who))
        cfg_path = refapi.write_cfg(base, o, str(tmp_path / f"{who}.cfg"))
        if who == "ref":
            assert refapi.run_reference_main(2, cfg_path) == 0
        else:
            r = subprocess.run([exe, cfg_path, "--dim", "2"], capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[who] = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / who / "*")))
    assert outs["ref"] == outs["gpu"]
    assert any(f.startswith("corr_") for f in outs["gpu"]) and any(f.startswith("final_") for f in outs["gpu"])
    for f in outs["gpu"]:
        a, b = (tmp_path / "ref" / f).read_bytes(), (tmp_path / "gpu" / f).read_bytes()
        if f.endswith(".pvd"):
            assert a == b, f
        elif f.endswith(".vti"):
            la, lb = a.split(b"\n"), b.split(b"\n")
            assert len(la) == len(lb), f
            bad = [(x, y) for x, y in zip(la, lb) if x != y]
            assert len(bad) <= 5e-3 * len(la), (f, len(bad), bad[:3])   # (lines of round-off sized values, |v| < 1e-12 max|v|)
            for x, y in bad:
                xs, ys = [float(t) for t in x.split()], [float(t) for t in y.split()]
                assert np.allclose(xs, ys, rtol=2e-5, atol=1e-10), (f, x, y)   # fields agree to 1e-12 of their maximum


@pytest.mark.parametrize("case", ["2d_dissolve", "3d_dissolve"])
def test_python_coupled_solver_whole_run(case, tmp_path):
    """solver.CoupledSolver.run (Python mirror of the coupling loop) against the same golden CSV; with
    write_vti it produces one state_, one final_ and a corr_ snapshot per diagnostics row."""
    import glob
    import os
    from pd_mg_pin_corrosion_b200 import solver as S
    from pd_mg_pin_corrosion_b200.grains import GrainStructure
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dim, cfg, _ = H.load_cfg(case, {"output_dir": str(tmp_path / "out")})
    grid = S.Grid(dim)
    grid.build(cfg)
    grains = GrainStructure().generate(grid.node_type, cfg, dim)
    fields = S.Fields()
    fields.allocate(grid.N_total, grid)
    S.initialize_fields(fields, grid, grains, cfg)
    cs = S.CoupledSolver()
    cs.log = lambda *a, **k: None
    cs.write_vti = (case == "2d_dissolve")
    cs.run(grid, fields, cfg)
    got = np.loadtxt(str(tmp_path / "out" / "diagnostics.csv"), delimiter=",", skiprows=1)
    if cs.write_vti:
        names = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "out" / "*.vti")))
        assert sum(n.startswith("corr_") for n in names) == got.shape[0]
        assert sum(n.startswith("state_") for n in names) == 1 and sum(n.startswith("final_") for n in names) == 1
        assert sum(n.startswith("flow_") for n in names) >= 1
        pvd = (tmp_path / "out" / "simulation.pvd").read_text()
        assert pvd.count("<DataSet") == got.shape[0] + 2
    gold = np.loadtxt(os.path.join(root, "tests", "golden", f"diagnostics_{case}.csv"), delimiter=",", skiprows=1)
    assert got.shape == gold.shape and np.array_equal(got[:, 3], gold[:, 3])
    for col in (0, 1, 2, 4, 5):
        rel = np.abs(got[:, col] - gold[:, col]) / np.maximum(np.abs(gold[:, col]), 1e-300)
        assert rel.max() <= 1e-6, (col, float(rel.max()))


@pytest.mark.parametrize("extra,expect_iters", [({"flow_conv_tol": 7e-5, "flow_max_iters": 3000}, 700),
                                                ({"flow_max_iters": 250, "channel_flow_corrections": 1}, 251)])
def test_solve_steady_matches_reference(extra, expect_iters):
    """PD_NS_Solver::solve_steady (src/pd_ns.cpp:182-372): same iteration count, eps and final
    state (on convergence the un-swapped pre-step state); second case exercises the channel-flow
    corrections (:209-270)."""
    case = "2d_poiseuille"
    ref = H.make_ref(case, extra)
    S, cfg, grid, fields = gpu_side(case, extra, ref=ref)
    ns = S.PD_NS_Solver()
    ns.init(grid, cfg)
    it_ref = ref.ns_solve_steady()
    it = ns.solve_steady(fields, grid, cfg, verbose=False)
    assert it == it_ref == expect_iters
    assert_fields(fields, ref, ("rho", "vel", "rho_new", "vel_new"))
    if expect_iters == 700:
        assert ns.last.status == 0 and abs(ns.last.eps - 6.438e-5) < 5e-8
        assert 0 < ns.last.poiseuille_l2 < 0.1 and ns.last.poiseuille_nodes > 0


@pytest.mark.parametrize("case,iters,eps,l2", [("2d_poiseuille", 19000, 8.612e-07, 6.427e-04),
                                               ("2d_default", 25300, 4.945e-06, 1.132e-02)])
def test_steady_state_known_answers(case, iters, eps, l2):
    """Full flow solves of BASELINE configs 1 and 2 to convergence: iteration count, epsilon and the
    Poiseuille L2 the reference prints, and the converged velocity field
    against tests/golden/steady.json (generated from oracle/_ref)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gold = json.load(open(os.path.join(root, "tests", "golden", "steady.json")))[case]
    from pd_mg_pin_corrosion_b200 import solver as S
    from pd_mg_pin_corrosion_b200.grains import GrainStructure
    dim, cfg, _ = H.load_cfg(case)
    grid = S.Grid(dim)
    grid.build(cfg)
    grains = GrainStructure().generate(grid.node_type, cfg, dim)
    fields = S.Fields()
    fields.allocate(grid.N_total, grid)
    S.initialize_fields(fields, grid, grains, cfg)
    ns = S.PD_NS_Solver()
    ns.init(grid, cfg)
    it = ns.solve_steady(fields, grid, cfg, verbose=False)
    assert it == iters == gold["iters"]
    assert abs(ns.last.eps - eps) <= 1e-3 * eps          # the reference prints 4 significant digits
    assert abs(ns.last.poiseuille_l2 - l2) <= 1e-3 * l2
    v = fields.get("vel")
    assert H.rel_err(v[::97], np.array(gold["vel_sample"])) <= 1e-10   # ~2e4 iterations of rounding drift
    assert H.rel_err(fields.get("rho")[::97], np.array(gold["rho_sample"])) <= 1e-12


def test_full_size_params_fine_vs_port_oracle():
    """BASELINE config 4 at full size (3D params_fine, dx = 2 um, 157 x 157 x 707 = 17.4 M nodes,
    2.39 G CSR-equivalent bonds -- beyond the reference's int32 CSR): classification,
    wall-mirror table and FIVE full NS + ARD loop-body pairs against the stencil-implicit plain-C oracle."""
    import os
    from oracle.portapi import PortSim
    from pd_mg_pin_corrosion_b200 import solver as S
    from pd_mg_pin_corrosion_b200.config import Config
    from pd_mg_pin_corrosion_b200.grains import GrainStructure
    cfg = Config.load(os.path.join(H.CONFIG_DIR, "params_fine.cfg"), {"use_implicit": 0}, quiet=True)
    port = PortSim(3, cfg, threads=os.cpu_count() or 4)
    grid = S.Grid(3)
    grid.build(cfg)
    assert (grid.Nx, grid.Ny, grid.Nz) == (157, 157, 707) == (port.Nx, port.Ny, port.Nz)
    nt = grid.node_type
    assert np.array_equal(nt, port.node_type)
    assert np.array_equal(grid.wall_mirror, port.wall_mirror)
    assert grid.info.ns_bonds == 178 * int((nt == 0).sum())          # every FLUID row is full (appendix A)
    grains = GrainStructure().generate(nt, cfg, 3)
    port.init_fields(grains.is_grain_boundary, grains.is_precipitate)
    # non-trivial state: perturb, then let both sides start from the identical arrays
    rng = np.random.default_rng(11)
    fl = nt != 5
    port.rho *= 1.0 + 1e-4 * rng.standard_normal(port.N) * fl
    port.vel += 1e-3 * rng.standard_normal(port.vel.shape) * (nt == 0)[:, None]
    port.C[:] = np.abs(port.C + 0.02 * rng.standard_normal(port.N)) * fl
    port.rho_new[:] = port.rho; port.vel_new[:] = port.vel; port.C_new[:] = port.C
    fields = S.Fields()
    fields.allocate(grid.N_total, grid)
    S.initialize_fields(fields, grid, grains, cfg)
    for n in ("rho", "vel", "C", "rho_new", "vel_new", "C_new"):
        fields.set(n, getattr(port, n))
    ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
    ns.init(grid, cfg); ard.init(grid, cfg)
    dt = port.ns_compute_dt()
    assert abs(ns.compute_dt(fields, grid, cfg) - dt) <= 1e-15 * dt
    dtc = None
    for rep in range(5):
        port.ns_iterate(1, dt)
        ns.iterate(fields, grid, cfg, 1, dt)
        if rep in (0, 4):
            for n in ("rho", "vel"):
                assert H.rel_err(fields.get(n), getattr(port, n)) <= TOL, (rep, n)
        if dtc is None:
            dtc = port.ard_compute_dt()
        port.ard_iterate(1, dtc)
        ard.iterate(fields, grid, cfg, 1, dtc)
        if rep in (0, 4):
            for n in ("rho", "vel", "C"):
                assert H.rel_err(fields.get(n), getattr(port, n)) <= TOL, (rep, n)


def test_dx1um_cross_section_vs_port_oracle():
    """BASELINE config 5 cross-section (params_fine geometry at dx = 1 um: 307 x 307 nodes per plane, where
    the outlet sweep needs the 128-row single-row ring), tube shortened to 63 planes so that the plain-C
    oracle finishes in seconds: classification, wall mirror and 3 NS + 2 ARD loop bodies."""
    import os
    from oracle.portapi import PortSim
    from pd_mg_pin_corrosion_b200 import solver as S
    from pd_mg_pin_corrosion_b200.config import Config
    cfg = Config.load(os.path.join(H.CONFIG_DIR, "params_fine.cfg"),
                      {"use_implicit": 0, "dx": 1.0e-6, "L_wire": 24e-6, "L_upstream": 16e-6, "L_downstream": 16e-6},
                      quiet=True)
    port = PortSim(3, cfg, threads=os.cpu_count() or 4)
    grid = S.Grid(3)
    grid.build(cfg)
    assert (grid.Nx, grid.Ny, grid.Nz) == (port.Nx, port.Ny, port.Nz) and grid.Nx == 307
    nt = grid.node_type
    assert np.array_equal(nt, port.node_type)
    assert np.array_equal(grid.wall_mirror, port.wall_mirror)
    port.init_fields()
    rng = np.random.default_rng(5)
    fl = nt != 5
    port.rho *= 1.0 + 1e-4 * rng.standard_normal(port.N) * fl
    port.vel += 1e-3 * rng.standard_normal(port.vel.shape) * (nt == 0)[:, None]
    port.C[:] = np.abs(port.C + 0.02 * rng.standard_normal(port.N)) * fl
    port.rho_new[:] = port.rho; port.vel_new[:] = port.vel; port.C_new[:] = port.C
    fields = S.Fields()
    fields.allocate(grid.N_total, grid)
    S.initialize_fields(fields, grid, None, cfg)
    for n in ("rho", "vel", "C", "rho_new", "vel_new", "C_new"):
        fields.set(n, getattr(port, n))
    ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
    ns.init(grid, cfg); ard.init(grid, cfg)
    dt = port.ns_compute_dt()
    port.ns_iterate(3, dt)
    ns.iterate(fields, grid, cfg, 3, dt)
    for n in ("rho", "vel"):
        assert H.rel_err(fields.get(n), getattr(port, n)) <= TOL, n
    dtc = port.ard_compute_dt()
    port.ard_iterate(2, dtc)
    ard.iterate(fields, grid, cfg, 2, dtc)
    for n in ("rho", "vel", "C"):
        assert H.rel_err(fields.get(n), getattr(port, n)) <= TOL, n
