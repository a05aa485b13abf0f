/* oracle/pd_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, stencil-implicit CPU restatement of the reference's explicit PD hot path
 * (grid classification + neighbour list, PD-NS step, explicit PD-ARD step, boundary
 * operators, CFL/phase-change logic).  "Stencil-implicit" = neighbours are enumerated
 * from the fixed horizon-offset table in the reference's CSR order instead of a
 * materialised CSR, so it also covers sizes where the reference's int32 CSR overflows
 * (3D dx <= 2 um).
 *
 * Parity status: PINNED -- tests/test_oracle_vs_ref.py checks every function here
 * against oracle/_ref (the unmodified reference compiled from its own sources) on the
 * reference's configs, and tests/golden/ holds vectors generated from oracle/_ref by
 * tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this.
 */
#ifndef PD_ORACLE_H
#define PD_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Numeric Config after compute_derived() (reference src/config.h:4-99,
 * src/config.cpp:98-112).  Same member order as include/pdgpu.h::PdConfig so the
 * test harness can share one ctypes.Structure; the two definitions are independent. */
typedef struct PdoConfig {
    double dx, R_wire, L_wire, R_tube, L_upstream, L_downstream;
    double rho_f, mu_f, gamma_eos, c0, eta_density, Q_flow, rho_m;
    double D_liquid, D_grain, D_gb, D_precip;
    double C_solid_init, C_liquid_init, C_thresh, C_sat, alpha_art_diff, corrosion_decay_l;
    double cfl_factor, cfl_factor_corr, flow_conv_tol, T_final;
    double delta, U_in;
    int m_ratio, flow_max_iters, corrosion_steps_per_check;
    int output_every_flow, output_every_corr, channel_flow_corrections, use_implicit;
    int reserved;
} PdoConfig;

enum { PDO_FLUID = 0, PDO_SOLID = 1, PDO_WALL = 2, PDO_INLET = 3, PDO_OUTLET = 4, PDO_OUTSIDE = 5 };

/* Lattice + stencil description (host arrays owned by the caller / pdo_grid_free). */
typedef struct PdoGrid {
    int dim, Nx, Ny, Nz, m, n_off;
    long long N;
    double dx, delta, origin[3];
    uint8_t* node_type;      /* [N] */
    int* off_d;              /* [n_off][3] (di,dj,dk), reference CSR order */
    double* off_dist;        /* [n_off] */
    double* off_evec;        /* [n_off][dim] */
    double* off_vol;         /* [n_off] beta * dx^dim */
    int* wall_mirror;        /* [N]: mirror node of a WALL node, -1 = none / not WALL */
} PdoGrid;

void pdo_set_threads(int n);

/* Grid::build (src/grid.cpp:29-155) + stencil of build_neighbors (:161-187,274-288). */
PdoGrid* pdo_grid_build(const PdoConfig* cfg, int dim);
void pdo_grid_free(PdoGrid* g);
/* accessors for ctypes */
void pdo_grid_info(const PdoGrid* g, long long* out /* dim,Nx,Ny,Nz,m,n_off,N */, double* origin);
uint8_t* pdo_grid_types(PdoGrid* g);
int* pdo_grid_off_d(PdoGrid* g);
double* pdo_grid_off_dist(PdoGrid* g);
double* pdo_grid_off_evec(PdoGrid* g);
double* pdo_grid_off_vol(PdoGrid* g);
int* pdo_grid_wall_mirror(PdoGrid* g);

/* CSR (src/grid.cpp:190-291); offsets are 64-bit here. */
void pdo_csr_offsets(const PdoGrid* g, long long* offset /* [N+1] */);
void pdo_csr_fill(const PdoGrid* g, const long long* offset, int* index, double* dist,
                  double* evec, double* vol);

/* Static wall-mirror table (src/boundary.cpp:143-264); call again after a phase change. */
void pdo_wall_mirror_build(PdoGrid* g, const PdoConfig* cfg);

/* Boundary operators (src/boundary.cpp). vel is [N][dim] like std::vector<Vec>. */
void pdo_inlet_bc(const PdoGrid* g, const PdoConfig* cfg, double* rho, double* vel, double* Cc);
void pdo_outlet_bc(const PdoGrid* g, const PdoConfig* cfg, double* rho, double* vel, double* Cc);
void pdo_wall_bc(const PdoGrid* g, const PdoConfig* cfg, double* rho, double* vel);
void pdo_wall_conc_bc(const PdoGrid* g, double* Cc);
/* smooth_boundary_concentration (src/boundary.cpp:332-376) */
void pdo_smooth_conc(const PdoGrid* g, const PdoConfig* cfg, double* Cc);
void pdo_solid_bc(const PdoGrid* g, double* vel);

/* PD-NS (src/pd_ns.cpp:36-180). */
double pdo_max_fluid_speed(const PdoGrid* g, const double* vel);
double pdo_ns_compute_dt(const PdoGrid* g, const PdoConfig* cfg, const double* vel);
void pdo_ns_step(const PdoGrid* g, const PdoConfig* cfg, double dt, const double* rho,
                 const double* vel, double* pressure, double* rho_new, double* vel_new);
/* convergence block of solve_steady (src/pd_ns.cpp:273-301): out = num, den, vmax,
 * rho_min, rho_max, has_nan */
void pdo_ns_residual(const PdoGrid* g, const double* vel, const double* vel_new,
                     const double* rho_new, double* out);

/* PD-ARD explicit (src/pd_ard.cpp:34-212). */
double pdo_ard_compute_dt(const PdoGrid* g, const PdoConfig* cfg, const double* vel);
void pdo_ard_step(const PdoGrid* g, const PdoConfig* cfg, double dt, double volume_loss,
                  const double* Cc, const double* vel, const uint8_t* is_gb,
                  const uint8_t* is_precip, double* C_new);
int pdo_phase_change(PdoGrid* g, const PdoConfig* cfg, uint8_t* phase, double* rho, double* vel,
                     double* Cc, int* dissolved /* [N] scratch, may be NULL */);

/* Loop bodies (src/pd_ns.cpp:196-205,325 and src/coupling.cpp:232-240); buffers are
 * swapped by pointer inside, results end in the first-named arrays. Returns seconds. */
double pdo_ns_iterate(PdoGrid* g, const PdoConfig* cfg, int iters, double dt, double* rho,
                      double* vel, double* pressure, double* Cc, double* rho_new, double* vel_new);
double pdo_ard_iterate(PdoGrid* g, const PdoConfig* cfg, int steps, double dt, double volume_loss,
                       double* rho, double* vel, double* Cc, double* C_new, const uint8_t* is_gb,
                       const uint8_t* is_precip);

/* VTKWriter::write (src/vtk_writer.cpp:16-146): ASCII ImageData snapshot, `ostream << double`
 * restated as printf("%g"). grain_id / D_map may be NULL (-1 / 0). Returns 0 on success. */
int pdo_write_vti(const PdoGrid* g, const char* path, const double* rho, const double* vel,
                  const double* pressure, const double* Cc, const uint8_t* phase, const int* grain_id,
                  const double* D_map, const uint8_t* is_gb, const uint8_t* is_precip);

#ifdef __cplusplus
}
#endif
#endif
