// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin C entry points over the UNMODIFIED reference objects, compiled from the
// reference sources where they lie under /root/reference/src by oracle/Makefile
// into oracle/_ref/libpdref{2,3}d.so.  Nothing from the reference is copied into
// this repository: this file only #includes the reference's own translation
// units / headers at build time (-I/root/reference/src).
//
// What it wraps (reference file:line):
//   Config::load / compute_derived            src/config.cpp:16,98
//   Grid::build / build_neighbors             src/grid.cpp:29,157
//   Grid::build_amr / build_neighbors_celllist / update_fictitious   src/grid.cpp:352,660,802
//   GrainStructure::generate                  src/grains.cpp:9   (forced 1 thread)
//   initialize_fields (static in main.cpp)    src/main.cpp:9
//   apply_*_bc                                src/boundary.cpp:31,88,288,292,302,381
//   PD_NS_Solver::{init,compute_dt,step,solve_steady}   src/pd_ns.cpp:7,52,78,182
//   PD_ARD_Solver::{init,compute_dt,step,apply_phase_change} src/pd_ard.cpp:6,34,55,193
//   main()                                    src/main.cpp:129  (whole-run diagnostics.csv)
//
// The implicit solver (src/pd_ard_implicit.cpp) needs Eigen 3.4.0, which the
// reference fetches from the network.  Two builds of this file (oracle/Makefile):
//   libpdref{2,3}d.so     explicit path only: -Ieigen_stub (empty declarations), the solver's five public
//                         methods are aborting stubs defined below.
//   libpdrefimp{2,3}d.so  -DPD_REF_IMPLICIT -Ieigen_min: the UNMODIFIED src/pd_ard_implicit.cpp compiled against
//                         an independently written work-alike of the few Eigen types it uses (oracle/eigen_min/).
//                         Assembly, right-hand side, clamp, adaptive step and the implicit coupling loop are then
//                         the reference's own code; the linear solve meets the reference's tolerance (1e-10) with
//                         a different Krylov implementation.  ref_imp_* below wrap
//   PD_ARD_ImplicitSolver::{init,assemble,step,compute_adaptive_dt,apply_phase_change}  src/pd_ard_implicit.cpp:9,104,371,438,538
#include <omp.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// Pull in the reference driver TU so that its file-static initialize_fields()
// is reachable; its main() is renamed, not modified.
#define main pd_reference_main
#include "main.cpp"
#undef main

#include "boundary.h"

#ifndef PD_REF_IMPLICIT
// ---- aborting stubs for the implicit solver (explicit-only build) -----------
static void implicit_abort(const char* what) {
    std::fprintf(stderr,
                 "oracle/_ref: PD_ARD_ImplicitSolver::%s called, but the implicit "
                 "(Eigen) branch is not built. Set use_implicit = 0.\n", what);
    std::abort();
}
void PD_ARD_ImplicitSolver::init(const Grid&, const Config&) { implicit_abort("init"); }
void PD_ARD_ImplicitSolver::assemble(const Fields&, const Grid&, const Config&) { implicit_abort("assemble"); }
int PD_ARD_ImplicitSolver::step(Fields&, const Grid&, const Config&, double) { implicit_abort("step"); return 0; }
double PD_ARD_ImplicitSolver::compute_adaptive_dt(const Fields&, const Grid&, const Config&) const {
    implicit_abort("compute_adaptive_dt");
    return 0.0;
}
int PD_ARD_ImplicitSolver::apply_phase_change(Fields&, Grid&, const Config&) {
    implicit_abort("apply_phase_change");
    return 0;
}
#endif

namespace {
struct RefSim {
    Config cfg;
    Grid grid;
    GrainStructure grains;
    Fields fields;
    PD_NS_Solver ns;
    PD_ARD_Solver ard;
    PD_ARD_ImplicitSolver imp;
    bool have_grains = false;
};
inline RefSim* S(void* h) { return static_cast<RefSim*>(h); }
}  // namespace

extern "C" {

int ref_dim() { return DIM; }
void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_max_threads() { return omp_get_max_threads(); }

void* ref_create(const char* cfg_path) {
    RefSim* s = new RefSim();
    s->cfg.load(cfg_path);
    return s;
}
void ref_destroy(void* h) { delete S(h); }

// Derived / mutated configuration values (c0 is raised to 25*U_in by compute_derived).
void ref_get_config(void* h, double* out) {
    const Config& c = S(h)->cfg;
    double v[] = {c.dx, (double)c.m_ratio, c.R_wire, c.L_wire, c.R_tube, c.L_upstream,
                  c.L_downstream, c.rho_f, c.mu_f, c.gamma_eos, c.c0, c.eta_density,
                  c.Q_flow, c.D_liquid, c.D_grain, c.D_gb, c.D_precip, c.C_solid_init,
                  c.C_liquid_init, c.C_thresh, c.C_sat, c.alpha_art_diff,
                  c.corrosion_decay_l, c.cfl_factor, c.cfl_factor_corr, c.delta, c.U_in,
                  (double)c.flow_max_iters, c.flow_conv_tol, c.T_final,
                  (double)c.corrosion_steps_per_check, (double)c.output_every_corr,
                  (double)c.use_implicit, (double)c.channel_flow_corrections,
                  c.precip_fraction, c.grain_size_mean, (double)c.gb_width_cells,
                  (double)c.precip_cluster_cells, (double)c.output_every_flow};
    std::memcpy(out, v, sizeof(v));
}
int ref_config_len() { return 39; }
// the remaining numeric members (implicit branch, AMR, their derived values) + rho_m
void ref_get_config_extra(void* h, double* out) {
    const Config& c = S(h)->cfg;
    double v[] = {c.implicit_dt_fraction, c.implicit_dt_max, (double)c.implicit_output_every, (double)c.diagnostic_every,
                  c.newton_tol, (double)c.newton_max_iter, (double)c.use_amr, (double)c.amr_ratio, c.amr_buffer,
                  c.dx_coarse, c.delta_coarse, c.rho_m};
    std::memcpy(out, v, sizeof(v));
}
const char* ref_config_output_dir(void* h) { return S(h)->cfg.output_dir.c_str(); }
// re-run Config::load on the existing object: keys present in `path` override, compute_derived runs again
void ref_config_apply(void* h, const char* path) { S(h)->cfg.load(path); }

void ref_grid_build(void* h) { S(h)->grid.build(S(h)->cfg); }
// two-level AMR grid (src/grid.cpp:352-842): what main() calls with use_amr = 1 (src/main.cpp:151-154)
void ref_grid_build_amr(void* h) { S(h)->grid.build_amr(S(h)->cfg); }
void ref_build_neighbors_celllist(void* h) { S(h)->grid.build_neighbors_celllist(S(h)->cfg); }
void ref_update_fictitious(void* h) { S(h)->grid.update_fictitious(S(h)->fields); }
long long ref_fict_entries(void* h) { return (long long)S(h)->grid.fict_source.size(); }
void ref_build_neighbors(void* h) { S(h)->grid.build_neighbors(); }

void ref_generate_grains(void* h) {
    // vector<bool> write races in grains.cpp:76-105,152-166 -> single thread.
    int prev = omp_get_max_threads();
    omp_set_num_threads(1);
    S(h)->grains = GrainStructure();
    S(h)->grains.generate(S(h)->grid, S(h)->cfg);
    omp_set_num_threads(prev);
    S(h)->have_grains = true;
}
int ref_n_grains(void* h) { return S(h)->grains.n_grains; }

void ref_fields_init(void* h) {
    RefSim* s = S(h);
    if (!s->have_grains) ref_generate_grains(h);
    s->fields = Fields();
    s->fields.allocate(s->grid.N_total);
    initialize_fields(s->fields, s->grid, s->grains, s->cfg);
}

// out: Nx, Ny, Nz, N_total, nnz (CSR entries, -1 if not built)
void ref_get_dims(void* h, long long* out) {
    const Grid& g = S(h)->grid;
    out[0] = g.Nx; out[1] = g.Ny; out[2] = g.Nz; out[3] = g.N_total;
    out[4] = g.nbr_offset.empty() ? -1 : (long long)g.nbr_offset[g.N_total];
}
void ref_get_origin(void* h, double* out) {
    const Grid& g = S(h)->grid;
    out[0] = g.origin_x; out[1] = g.origin_y; out[2] = g.origin_z;
}

void* ref_ptr(void* h, const char* name) {
    RefSim* s = S(h);
    std::string n(name);
    Grid& g = s->grid;
    Fields& f = s->fields;
    if (n == "pos") return g.pos.data();
    if (n == "node_type") return g.node_type.data();
    if (n == "nbr_offset") return g.nbr_offset.data();
    if (n == "nbr_index") return g.nbr_index.data();
    if (n == "nbr_dist") return g.nbr_dist.data();
    if (n == "nbr_evec") return g.nbr_evec.data();
    if (n == "nbr_vol") return g.nbr_vol.data();
    if (n == "dx_local") return g.dx_local.data();
    if (n == "delta_local") return g.delta_local.data();
    if (n == "grid_level") return g.grid_level.data();
    if (n == "fict_offset") return g.fict_offset.data();
    if (n == "fict_source") return g.fict_source.data();
    if (n == "fict_weight") return g.fict_weight.data();
    if (n == "rho") return f.rho.data();
    if (n == "vel") return f.vel.data();
    if (n == "pressure") return f.pressure.data();
    if (n == "C") return f.C.data();
    if (n == "D_map") return f.D_map.data();
    if (n == "phase") return f.phase.data();
    if (n == "grain_id") return f.grain_id.data();
    if (n == "is_gb") return f.is_gb.data();
    if (n == "is_precip") return f.is_precip.data();
    if (n == "rho_new") return f.rho_new.data();
    if (n == "vel_new") return f.vel_new.data();
    if (n == "C_new") return f.C_new.data();
    return nullptr;
}

// ---- boundary operators -----------------------------------------------------
void ref_apply_inlet_bc(void* h) { apply_inlet_bc(S(h)->fields, S(h)->grid, S(h)->cfg); }
void ref_apply_outlet_bc(void* h) { apply_outlet_bc(S(h)->fields, S(h)->grid, S(h)->cfg); }
void ref_apply_wall_bc(void* h) { apply_wall_bc(S(h)->fields, S(h)->grid, S(h)->cfg); }
void ref_apply_wall_bc_new(void* h) { apply_wall_bc_new(S(h)->fields, S(h)->grid, S(h)->cfg); }
void ref_apply_wall_concentration_bc(void* h) {
    apply_wall_concentration_bc(S(h)->fields, S(h)->grid, S(h)->cfg);
}
void ref_apply_solid_surface_bc(void* h) { apply_solid_surface_bc(S(h)->fields, S(h)->grid); }
void ref_smooth_boundary_concentration(void* h) { smooth_boundary_concentration(S(h)->fields, S(h)->grid, S(h)->cfg); }
void ref_update_node_types(void* h) { update_node_types_after_dissolution(S(h)->grid, S(h)->fields); }

// ---- PD-NS --------------------------------------------------------------------
void ref_ns_init(void* h) { S(h)->ns.init(S(h)->grid, S(h)->cfg); }
double ref_ns_compute_dt(void* h) { return S(h)->ns.compute_dt(S(h)->fields, S(h)->grid, S(h)->cfg); }
void ref_ns_step(void* h, double dt) { S(h)->ns.step(S(h)->fields, S(h)->grid, S(h)->cfg, dt); }
int ref_ns_solve_steady(void* h) { return S(h)->ns.solve_steady(S(h)->fields, S(h)->grid, S(h)->cfg); }
void ref_swap_flow(void* h) { S(h)->fields.swap_buffers(); }

// One full solve_steady iteration body without the convergence block
// (src/pd_ns.cpp:196-205,325): BCs, step, wall_new, swap.
void ref_ns_iterate(void* h, int iters, double dt) {
    RefSim* s = S(h);
    for (int it = 0; it < iters; ++it) {
        apply_inlet_bc(s->fields, s->grid, s->cfg);
        apply_outlet_bc(s->fields, s->grid, s->cfg);
        apply_wall_bc(s->fields, s->grid, s->cfg);
        apply_solid_surface_bc(s->fields, s->grid);
        s->ns.step(s->fields, s->grid, s->cfg, dt);
        apply_wall_bc_new(s->fields, s->grid, s->cfg);
        s->fields.swap_buffers();
    }
}

// The same loop body on the AMR grid: solve_steady refreshes the FICTITIOUS nodes after the swap (:328).
void ref_ns_iterate_amr(void* h, int iters, double dt) {
    RefSim* s = S(h);
    for (int it = 0; it < iters; ++it) {
        apply_inlet_bc(s->fields, s->grid, s->cfg);
        apply_outlet_bc(s->fields, s->grid, s->cfg);
        apply_wall_bc(s->fields, s->grid, s->cfg);
        apply_solid_surface_bc(s->fields, s->grid);
        s->ns.step(s->fields, s->grid, s->cfg, dt);
        apply_wall_bc_new(s->fields, s->grid, s->cfg);
        s->fields.swap_buffers();
        s->grid.update_fictitious(s->fields);
    }
}

// ---- PD-ARD (explicit) ---------------------------------------------------------
void ref_ard_init(void* h) { S(h)->ard.init(S(h)->grid, S(h)->cfg); }
void ref_ard_set_volume_loss(void* h, double vl) { S(h)->ard.set_volume_loss(vl); }
double ref_ard_compute_dt(void* h) { return S(h)->ard.compute_dt(S(h)->fields, S(h)->grid, S(h)->cfg); }
void ref_ard_step(void* h, double dt) { S(h)->ard.step(S(h)->fields, S(h)->grid, S(h)->cfg, dt); }
int ref_ard_phase_change(void* h) { return S(h)->ard.apply_phase_change(S(h)->fields, S(h)->grid, S(h)->cfg); }
void ref_swap_C(void* h) { std::swap(S(h)->fields.C, S(h)->fields.C_new); }

// Explicit coupling-loop body (src/coupling.cpp:232-240): BCs, step, swap C.
void ref_ard_iterate(void* h, int steps, double dt) {
    RefSim* s = S(h);
    for (int it = 0; it < steps; ++it) {
        apply_inlet_bc(s->fields, s->grid, s->cfg);
        apply_outlet_bc(s->fields, s->grid, s->cfg);
        apply_wall_concentration_bc(s->fields, s->grid, s->cfg);
        s->ard.step(s->fields, s->grid, s->cfg, dt);
        std::swap(s->fields.C, s->fields.C_new);
    }
}

// Wall-clock of the two loop bodies, for bench.py's cpu_baseline / --impl reference.
double ref_time_ns_iterate(void* h, int iters, double dt) {
    double t0 = omp_get_wtime();
    ref_ns_iterate(h, iters, dt);
    return omp_get_wtime() - t0;
}
double ref_time_ard_iterate(void* h, int steps, double dt) {
    double t0 = omp_get_wtime();
    ref_ard_iterate(h, steps, dt);
    return omp_get_wtime() - t0;
}

// The reference's own VTKWriter::write (src/vtk_writer.cpp:16-146) on the current state; returns seconds.
double ref_write_vti(void* h, const char* path) {
    RefSim* s = S(h);
    VTKWriter w;
    double t0 = omp_get_wtime();
    w.write(path, s->grid, s->fields, s->cfg);
    return omp_get_wtime() - t0;
}

#ifdef PD_REF_IMPLICIT
// ---- PD-ARD (implicit) ---------------------------------------------------------
// The system of the last PD_ARD_ImplicitSolver::step, observed through the work-alike's solve hook:
// A = I - dt M (row-wise), b = C_old + dt bc_rhs, the solution before the clamp, iterations, relative residual.
namespace {
struct LastSolve {
    std::vector<long long> ptr;
    std::vector<int> col;
    std::vector<double> val, b, x;
    int iters = 0;
    double err = 0.0;
} g_last;
void record_solve(const Eigen::SparseMatrix<double>& A, const Eigen::VectorXd& b, const Eigen::VectorXd& x, int iters,
                  double err) {
    const long long n = (long long)A.rows();
    g_last.ptr.assign(1, 0);
    g_last.col.clear(); g_last.val.clear();
    for (long long i = 0; i < n; ++i) {
        for (const auto& e : A.row(i)) { g_last.col.push_back(e.first); g_last.val.push_back(e.second); }
        g_last.ptr.push_back((long long)g_last.col.size());
    }
    g_last.b.assign(b.data(), b.data() + n);
    g_last.x.assign(x.data(), x.data() + n);
    g_last.iters = iters;
    g_last.err = err;
}
}  // namespace
int ref_has_implicit() { return 1; }
void ref_imp_init(void* h) { Eigen::pd_solve_hook = record_solve; S(h)->imp.init(S(h)->grid, S(h)->cfg); }
void ref_imp_set_volume_loss(void* h, double vl) { S(h)->imp.set_volume_loss(vl); }
void ref_imp_assemble(void* h) { S(h)->imp.assemble(S(h)->fields, S(h)->grid, S(h)->cfg); }
double ref_imp_compute_adaptive_dt(void* h) { return S(h)->imp.compute_adaptive_dt(S(h)->fields, S(h)->grid, S(h)->cfg); }
int ref_imp_step(void* h, double dt) { return S(h)->imp.step(S(h)->fields, S(h)->grid, S(h)->cfg, dt); }
int ref_imp_phase_change(void* h) { return S(h)->imp.apply_phase_change(S(h)->fields, S(h)->grid, S(h)->cfg); }
// out: n, nnz, iterations; err: relative preconditioned residual
void ref_imp_last_info(long long* out, double* err) {
    out[0] = (long long)g_last.b.size(); out[1] = (long long)g_last.col.size(); out[2] = g_last.iters;
    *err = g_last.err;
}
void ref_imp_last_system(long long* ptr, int* col, double* val, double* b, double* x) {
    std::memcpy(ptr, g_last.ptr.data(), sizeof(long long) * g_last.ptr.size());
    std::memcpy(col, g_last.col.data(), sizeof(int) * g_last.col.size());
    std::memcpy(val, g_last.val.data(), sizeof(double) * g_last.val.size());
    std::memcpy(b, g_last.b.data(), sizeof(double) * g_last.b.size());
    std::memcpy(x, g_last.x.data(), sizeof(double) * g_last.x.size());
}
#else
int ref_has_implicit() { return 0; }
#endif

// The reference's own VTKWriter::write_vtu (src/vtk_writer.cpp:199-346; AMR clouds, 2D) on the current state.
void ref_write_vtu(void* h, const char* path) {
    RefSim* s = S(h);
    VTKWriter w;
    w.write_vtu(path, s->grid, s->fields, s->cfg);
}

// The reference's own main(): whole-run diagnostics.csv for the 1e-6 parity check.
int ref_main(const char* cfg_path) {
    char prog[] = "pd_corrosion";
    std::string p(cfg_path);
    char* argv[] = {prog, p.data(), nullptr};
    return pd_reference_main(2, argv);
}

}  // extern "C"
