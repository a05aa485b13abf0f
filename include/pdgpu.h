/* include/pdgpu.h -- C ABI of libpdgpu.so, the B200 (sm_100a) implementation of the
 * per-timestep peridynamic bond-summation hot path of alhermann/pd-mg-pin-corrosion.
 *
 * The reference has no FFI/plugin layer: the seam is the C++ surface
 * that src/main.cpp and src/coupling.cpp call.  Every entry point below names the
 * reference interface (file:line under the reference's src/) it replaces.  Plain
 * pointers and sizes only; the library owns all device memory; the caller is single
 * threaded per context (like the reference's host thread).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; pdgpu_last_error()
 *     returns a message for the calling thread's last failure.  There is NO CPU
 *     fallback: without a CUDA device pdgpu_create* fails.
 *   - host arrays use the reference's layouts: node index n = k*Nx*Ny + j*Nx + i
 *     (src/grid.h:58-64), velocity interleaved [N][dim] (std::vector<Vec>,
 *     src/utils.h:14), node types uint8 (src/grid.h:9-17).
 *   - a context covers a z-slab [a0,a1) of the axial index (k in 3D, j in 2D) of the
 *     global grid plus `reach` ghost planes per side; pdgpu_create makes the slab the
 *     whole grid.  Host arrays passed to upload/download are GLOBAL-sized; a slab
 *     context reads/writes only its own planes (and its ghosts on upload).
 */
#ifndef PDGPU_H
#define PDGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDGPU_VERSION 100

/* POD copy of the numeric members of the reference's Config AFTER compute_derived()
 * (src/config.h:4-99, src/config.cpp:98-112): c0 already raised to 25*U_in, delta and
 * U_in filled in. */
typedef struct PdConfig {
    double dx, R_wire, L_wire, R_tube, L_upstream, L_downstream;
    double rho_f, mu_f, gamma_eos, c0, eta_density, Q_flow, rho_m;
    double D_liquid, D_grain, D_gb, D_precip;
    double C_solid_init, C_liquid_init, C_thresh, C_sat, alpha_art_diff, corrosion_decay_l;
    double cfl_factor, cfl_factor_corr, flow_conv_tol, T_final;
    double delta, U_in;
    int m_ratio, flow_max_iters, corrosion_steps_per_check;
    int output_every_flow, output_every_corr, channel_flow_corrections, use_implicit;
    int reserved;
} PdConfig;

/* NodeType (src/grid.h:9-17). FICTITIOUS (6) is AMR-only and never produced. */
enum { PDGPU_FLUID = 0, PDGPU_SOLID_MG = 1, PDGPU_WALL = 2, PDGPU_INLET = 3, PDGPU_OUTLET = 4,
       PDGPU_OUTSIDE = 5 };

/* Members of Fields (src/fields.h:7-26) + Grid::node_type addressable by upload/download. */
enum { PDGPU_F_RHO = 0, PDGPU_F_VEL = 1, PDGPU_F_PRESSURE = 2, PDGPU_F_C = 3, PDGPU_F_RHO_NEW = 4,
       PDGPU_F_VEL_NEW = 5, PDGPU_F_C_NEW = 6, PDGPU_F_PHASE = 7, PDGPU_F_IS_GB = 8,
       PDGPU_F_IS_PRECIP = 9, PDGPU_F_NODE_TYPE = 10 };

typedef struct pdgpu_ctx pdgpu_ctx;

typedef struct PdGridInfo {
    int dim, Nx, Ny, Nz;          /* Grid::Nx,Ny,Nz (src/grid.h:20) */
    int m, n_off, reach;          /* m_ratio, stencil size (36 / 178 for m=3), ghost width */
    int a0, a1;                   /* owned axial planes [a0,a1) */
    long long N_total;            /* Nx*Ny*Nz */
    long long plane;              /* nodes per axial plane: Nx (2D) or Nx*Ny (3D) */
    long long counts[6];          /* owned nodes per NodeType */
    long long ns_bonds;           /* sum of CSR row lengths over owned FLUID rows   */
    long long ard_bonds;          /* ... over owned FLUID + SOLID_MG rows           */
    long long nnz;                /* CSR entries of owned rows (all non-OUTSIDE)    */
    double origin[3];             /* Grid::origin_x,y,z */
} PdGridInfo;

/* Convergence-block scalars of PD_NS_Solver::solve_steady (src/pd_ns.cpp:273-301). */
typedef struct PdResidual {
    double num, den;              /* sum |dv|^2, sum |v|^2 over FLUID */
    double v_max, rho_min, rho_max;
    int has_nan, pad;
} PdResidual;

/* Outcome of pdgpu_ns_solve_steady. status: 0 converged, 1 hit flow_max_iters,
 * 2 diverged (NaN), 3 diverged (v_max > 100 U_in). iters = the reference's return value. */
typedef struct PdSteadyResult {
    int iters, status;
    double eps, dt, v_max, rho_min, rho_max;
    double poiseuille_l2;         /* 2D only (src/pd_ns.cpp:341-368), -1 if not evaluated */
    int poiseuille_nodes, pad;
} PdSteadyResult;

/* Reductions of CoupledSolver::write_diagnostics (src/coupling.cpp:20-49). */
typedef struct PdDiag {
    long long solid_count;
    double v_max, C_max_fluid;
} PdDiag;

const char* pdgpu_last_error(void);
int pdgpu_version(void);
int pdgpu_device_count(int* count);

/* ---- host-only helpers (no device needed) -------------------------------------- */
/* Grid extents of Grid::build (src/grid.cpp:38-67). */
int pdgpu_grid_extents(const PdConfig* cfg, int dim, int* Nx, int* Ny, int* Nz, double origin[3]);
/* Balanced z-slab [a0,a1) of `rank` among `nranks` over n_axial planes. */
int pdgpu_partition(int n_axial, int nranks, int rank, int* a0, int* a1);
/* what pdgpu_create_slab uses: equal COST per rank -- planes within reach of the wire carry a measured surcharge
 * (the ARD kernel's general body, solid rows), so the wire slabs get slightly fewer planes. Deterministic in
 * (cfg, dim, nranks); env PDGPU_SLAB_PIN_COST = 0 gives equal plane counts. */
int pdgpu_partition_balanced(const PdConfig* cfg, int dim, int nranks, int rank, int* a0, int* a1);
int pdgpu_slab_layout_range(int a0, int a1, long long plane, int reach, long long* out);
/* Local layout of a z-slab (what the halo exchange uses): out[0..9] = a0, a1, local planes,
 * local nodes, own_lo, own_hi, send_lo, recv_lo, send_hi, recv_hi (node offsets into a local
 * array; a halo block is reach*plane nodes). New: the reference has no distributed layer. */
int pdgpu_slab_layout(int n_axial, long long plane, int reach, int nranks, int rank, long long* out);
/* Horizon-offset stencil of Grid::build_neighbors (src/grid.cpp:161-187,274-288) in CSR
 * order: off_d [n][3], dist [n], evec [n][dim], vol [n]; returns count in *n_off
 * (arrays may be NULL to query the count). */
int pdgpu_stencil(const PdConfig* cfg, int dim, int* n_off, int* off_d, double* dist, double* evec,
                  double* vol);

/* ---- lifetime --------------------------------------------------------------------- */
int pdgpu_create(const PdConfig* cfg, int dim, int device, pdgpu_ctx** out);
int pdgpu_create_slab(const PdConfig* cfg, int dim, int device, int rank, int nranks,
                      pdgpu_ctx** out);
int pdgpu_destroy(pdgpu_ctx* ctx);
int pdgpu_sync(pdgpu_ctx* ctx);

/* ---- Grid (src/grid.h:52-53) -------------------------------------------------------- */
/* Grid::build: node classification on device, stencil table, per-type node lists,
 * wall-mirror table, outlet sweep schedule. */
int pdgpu_grid_build(pdgpu_ctx* ctx);
/* Replace the classification by a caller-made one (hand-built geometries as in the
 * reference's tests/test_implicit.cpp:737-772); rebuilds the derived tables. */
int pdgpu_grid_set_types(pdgpu_ctx* ctx, const uint8_t* node_type_global);
int pdgpu_grid_info(pdgpu_ctx* ctx, PdGridInfo* out);
/* Grid::build_neighbors: materialise the reference-layout CSR on device
 * (count -> prefix scan -> fill). Optional: the solvers use the offset table. */
int pdgpu_grid_build_neighbors(pdgpu_ctx* ctx, long long* nnz);
/* nbr_offset is 64-bit here (int32 overflows at 3D dx=2um). Rows are the
 * context's owned nodes in index order; indices are GLOBAL node indices. */
int pdgpu_grid_download_csr(pdgpu_ctx* ctx, long long* nbr_offset, int* nbr_index,
                            double* nbr_dist, double* nbr_evec, double* nbr_vol);
int pdgpu_grid_free_neighbors(pdgpu_ctx* ctx);
/* Mirror node (global index, -1 = none) of every owned node; -1 for non-WALL nodes
 * (apply_wall_mirror_proper, src/boundary.cpp:143-264). [N_total], owned part written. */
int pdgpu_grid_download_wall_mirror(pdgpu_ctx* ctx, int* mirror_global);

/* ---- Fields (src/fields.h:28-58, src/main.cpp:9-127) ------------------------------- */
int pdgpu_fields_upload(pdgpu_ctx* ctx, int field, const void* host_global);
int pdgpu_fields_download(pdgpu_ctx* ctx, int field, void* host_global);
/* initialize_fields on device (is_gb / is_precip from the host grain generator, NULL = 0). */
int pdgpu_fields_init(pdgpu_ctx* ctx, const uint8_t* is_gb_global, const uint8_t* is_precip_global);
int pdgpu_swap_flow(pdgpu_ctx* ctx);                 /* Fields::swap_buffers       */
int pdgpu_swap_C(pdgpu_ctx* ctx);                    /* std::swap(C, C_new), coupling.cpp:238 */
/* out[t] = field[idx[t]] for scalar double fields (ordered host sums, coupling.cpp:32-38).
 * Slab contexts: COLLECTIVE -- every rank passes the same index list and receives every value
 * (each node is read by its owner), so that all ranks form the same ordered sum. */
int pdgpu_gather(pdgpu_ctx* ctx, int field, const int* idx_global, long long n, double* out);
/* Slab contexts: COLLECTIVE download of the WHOLE global array on every rank (host mirrors of the
 * coupling loop: node types for the initial-solid list of src/coupling.cpp:95-103, final fields).
 * One rank: same as pdgpu_fields_download. */
int pdgpu_fields_download_all(pdgpu_ctx* ctx, int field, void* host_global);

/* ---- boundary operators (src/boundary.h:6-13) ---------------------------------------- */
int pdgpu_bc_inlet(pdgpu_ctx* ctx);                  /* apply_inlet_bc               */
int pdgpu_bc_outlet(pdgpu_ctx* ctx);                 /* apply_outlet_bc (in-place GS)*/
int pdgpu_bc_wall(pdgpu_ctx* ctx);                   /* apply_wall_bc                */
int pdgpu_bc_wall_new(pdgpu_ctx* ctx);               /* apply_wall_bc_new            */
int pdgpu_bc_wall_conc(pdgpu_ctx* ctx);              /* apply_wall_concentration_bc  */
int pdgpu_bc_solid(pdgpu_ctx* ctx);                  /* apply_solid_surface_bc       */
/* smooth_boundary_concentration (src/boundary.cpp:332-376; called after every implicit ARD step,
 * src/coupling.cpp:186): in place, in the reference's index order. */
int pdgpu_bc_smooth_conc(pdgpu_ctx* ctx);

/* ---- PD_NS_Solver (src/pd_ns.h:9-17) --------------------------------------------------- */
int pdgpu_ns_compute_dt(pdgpu_ctx* ctx, double* dt);
int pdgpu_ns_step(pdgpu_ctx* ctx, double dt);        /* EOS + bond sums + Euler -> *_new */
/* `iters` x { inlet, outlet, wall, solid, step, wall_new, swap } (src/pd_ns.cpp:196-205,325),
 * device resident, no host round trip. */
int pdgpu_ns_iterate(pdgpu_ctx* ctx, int iters, double dt);
int pdgpu_ns_residual(pdgpu_ctx* ctx, PdResidual* out);
int pdgpu_ns_solve_steady(pdgpu_ctx* ctx, PdSteadyResult* out, int verbose);

/* ---- PD_ARD_Solver, explicit (src/pd_ard.h:9-20) ---------------------------------------- */
int pdgpu_ard_set_volume_loss(pdgpu_ctx* ctx, double vl);
int pdgpu_ard_compute_dt(pdgpu_ctx* ctx, double* dt);
int pdgpu_ard_step(pdgpu_ctx* ctx, double dt);       /* salt pre-pass + bond sums -> C_new */
/* `steps` x { inlet, outlet, wall_conc, step, swap C } (src/coupling.cpp:232-240). */
int pdgpu_ard_iterate(pdgpu_ctx* ctx, int steps, double dt);
/* `steps` x { NS loop body ; ARD loop body }, device resident, one host synchronisation at the end: the unit of the
 * throughput metric. Same state as pdgpu_ns_iterate(1) + pdgpu_ard_iterate(1) repeated. */
int pdgpu_step_iterate(pdgpu_ctx* ctx, int steps, double dt_ns, double dt_ard);

/* ---- One coupling-loop pass on HOST-resident Fields -------------------------------------
 * What a caller that keeps the reference's Fields vectors (src/fields.h:28-58) in host memory
 * runs per pass: rho/vel/C in, { NS loop body (src/pd_ns.cpp:196-205,325) ; ARD loop body
 * (src/coupling.cpp:232-240) } on the device, rho/vel/C out, in place. `vel` is the reference's
 * AoS std::vector<Vec> ([N][dim]); all arrays are global [N_total]. The axial planes are cut
 * into `n_chunks` chunks (0/1 = whole domain at once) whose uploads, kernels and downloads
 * overlap; results are bit-identical to upload + pdgpu_ns_iterate(1) + pdgpu_ard_iterate(1) +
 * download. Pin the arrays (pdgpu_host_register) or the copies serialise. */
int pdgpu_step_host(pdgpu_ctx* ctx, double dt_ns, double dt_ard, double* rho, double* vel, double* C,
                    int n_chunks);
/* Chunks pdgpu_step_host would use for `n_chunks` on this geometry, and why it fell back to 1. */
int pdgpu_step_host_chunks(pdgpu_ctx* ctx, int n_chunks, int* n_used, char* why, int why_len);
/* Timeline of the last chunked pdgpu_step_host call: per chunk (axial order) milliseconds from the
 * start of the call to { upload done, kernels done, download done }; out[3 * n_chunks]. */
int pdgpu_step_host_trace(pdgpu_ctx* ctx, double* out, int cap, int* n_chunks);
/* ---- Output staging (src/vtk_writer.cpp:16-146, src/coupling.cpp:242-249) ------
 * VTKWriter::write: the ASCII ImageData snapshot of the current state, formatted ON THE DEVICE
 * ("%g" of every value, WALL/OUTSIDE velocities zeroed, NaN/Inf/|v|<1e-300 flushed) and written to
 * `path`; byte-identical to the reference's file for the same state. grain_id[N] and D_map[N] are
 * host-side arrays of the driver (the solvers never read them); NULL writes -1 / 0.
 * bytes_out = bytes of DataArray bodies, format_ms = device time of the formatting kernels. */
int pdgpu_vti_write(pdgpu_ctx* ctx, const char* path, const int* grain_id, const double* D_map,
                    long long* bytes_out, float* format_ms);
/* printf("%g") of n doubles on the device into 16-byte zero-padded cells (formatter unit tests). */
int pdgpu_format_g(pdgpu_ctx* ctx, const double* host_vals, long long n, char* host_cells16);
/* Binary checkpoint / resume of the device-resident state (new: the reference cannot resume).
 * A run continued after pdgpu_checkpoint_load on a context built from the same Config is
 * bit-identical to the uninterrupted run. */
int pdgpu_checkpoint_save(pdgpu_ctx* ctx, const char* path, long long* bytes_out);
int pdgpu_checkpoint_load(pdgpu_ctx* ctx, const char* path);
/* cudaHostRegister / cudaHostUnregister of caller-owned memory (e.g. std::vector storage). */
int pdgpu_host_register(void* ptr, size_t bytes);
int pdgpu_host_unregister(void* ptr);
/* apply_phase_change + update_node_types + table rebuild (src/coupling.cpp:256-271).
 * dissolved_global (may be NULL) receives up to `cap` ascending global indices. */
int pdgpu_phase_change(pdgpu_ctx* ctx, int* n_dissolved, int* dissolved_global, int cap);
int pdgpu_diag(pdgpu_ctx* ctx, PdDiag* out);

/* ---- PD_ARD_ImplicitSolver (src/pd_ard_implicit.h:12-40), matrix-free ---------------
 * assemble(): salt-layer flags + interface diffusivities from the current C (once per coupling cycle,
 *   src/coupling.cpp:166); the operator M itself is never stored.
 * compute_adaptive_dt(): src/pd_ard_implicit.cpp:438-487 (implicit_dt_fraction, implicit_dt_max of Config).
 * step(): solves (I - dt M) C_new = C_old + dt bc_rhs with restarted GMRES on the device, clamps to
 *   [0, C_solid_init] and stores into the current C (:371-429). precond: 0 none, 1 Jacobi, 2 forward sweep
 *   over the axial planes. The reference's own solver is Eigen's GMRES + IncompleteLUT (tolerance 1e-10,
 *   restart 50, 200 iterations).
 * matvec / rhs: y = (I - dt M) x and b for GLOBAL host vectors (inspection, tests). Single-GPU contexts. */
typedef struct PdLinSolveInfo {
    int iters, converged;
    double rel_res;               /* ||b - A x|| / ||b|| of the returned iterate */
    int pad;
} PdLinSolveInfo;
/* SOLID_MG nodes with C < C_thresh (all ranks): the implicit cycle ends at the first one (src/coupling.cpp:206-211) */
int pdgpu_solid_below_thresh(pdgpu_ctx* ctx, int* count);
int pdgpu_implicit_assemble(pdgpu_ctx* ctx);
int pdgpu_implicit_compute_dt(pdgpu_ctx* ctx, double dt_fraction, double dt_max, double* dt);
int pdgpu_implicit_step(pdgpu_ctx* ctx, double dt, double tol, int restart, int max_iters, int precond,
                        PdLinSolveInfo* info);
int pdgpu_implicit_matvec(pdgpu_ctx* ctx, double dt, const double* x_global, double* y_global);
int pdgpu_implicit_rhs(pdgpu_ctx* ctx, double dt, double* b_global);

/* ---- GrainStructure::generate on device (src/grains.cpp:55-107,152-166) -------
 * The lattice passes of the grain generator: nearest-seed Voronoi assignment of every SOLID_MG node
 * (brute force over the seeds), grain-boundary detection over the immediate neighbours and
 * gb_width_cells dilation passes; then the ball growth of clustered precipitates.  The random draws
 * (seed picks, shuffle) stay with the caller: host/grains.cpp keeps the reference's libstdc++ RNG call
 * sequence and calls these two.  Outputs are WHOLE global arrays on every rank (collective for slabs). */
int pdgpu_grains_voronoi(pdgpu_ctx* ctx, const double* seeds_xyz, int n_grains, int gb_width_cells,
                         int* grain_id_global, uint8_t* is_gb_global);
int pdgpu_grains_grow_precip(pdgpu_ctx* ctx, const uint8_t* is_gb_global, const uint8_t* seed_flags_global,
                             int cluster_cells, uint8_t* is_precip_global);

/* ---- multi-GPU: one process per GPU, z-slabs, NCCL halo exchange ------------------------ */
int pdgpu_comm_uid_bytes(void);
int pdgpu_comm_get_uid(void* uid_out);               /* rank 0; broadcast by the caller */
int pdgpu_comm_init(pdgpu_ctx* ctx, const void* uid, int rank, int nranks);
int pdgpu_halo_exchange(pdgpu_ctx* ctx, int which);  /* 0: rho,vel,p   1: C   2: all (incl. _new) */
/* Host scalars of the coupling loop combined over the ranks (op 0 sum, 1 max, 2 min), e.g. the
 * dissolved-node count of a check (src/coupling.cpp:256-275). No-op for one rank. n <= 1024. */
int pdgpu_comm_allreduce(pdgpu_ctx* ctx, double* host_vals, int n, int op);

/* ---- two-level AMR grid (2D like the reference) ------------------------------
 * Replaces Grid::build_amr / build_neighbors_celllist / update_fictitious (src/grid.cpp:352-842) and the AMR
 * branches of the explicit solvers and BCs (src/pd_ns.cpp:18-33,328; src/pd_ard.cpp:17-31,86,130;
 * src/boundary.cpp:185-201). The grid build is host code (no device needed until pdamr_device_init);
 * every array it produces is bit-identical to the reference's. amr_ratio / amr_buffer: src/config.h:86-88. */
typedef struct pdamr_ctx pdamr_ctx;
typedef struct PdAmrInfo {
    long long N_total, n_fine, n_coarse, n_fict, nnz, n_fict_entries;   /* nnz = -1 before build_neighbors */
    long long counts[7];                                                 /* per NodeType, FICTITIOUS = 6 */
    double origin[2], dx_coarse, delta_coarse;
} PdAmrInfo;
int pdamr_create(const PdConfig* cfg, int amr_ratio, double amr_buffer, pdamr_ctx** out);
int pdamr_build(pdamr_ctx* ctx);                 /* Grid::build_amr */
int pdamr_build_neighbors(pdamr_ctx* ctx);       /* Grid::build_neighbors_celllist (+ wall-mirror table, outlet levels) */
int pdamr_info(pdamr_ctx* ctx, PdAmrInfo* out);
/* host copies: pos [N][2], node_type (u8), dx_local, delta_local, grid_level (int), fict_offset [N+1], fict_source,
 * fict_weight, nbr_offset [N+1], nbr_index, nbr_dist, nbr_evec [nnz][2], nbr_vol, wall_mirror (int, -2: not a WALL) */
int pdamr_get(pdamr_ctx* ctx, const char* name, void* out);
int pdamr_device_init(pdamr_ctx* ctx, int device);   /* uploads the tables, allocates the fields; needs a B200 */
/* fields by the reference's names: rho, vel [N][2], pressure, C, rho_new, vel_new, C_new, phase, is_gb, is_precip */
int pdamr_field_set(pdamr_ctx* ctx, const char* name, const void* src);
int pdamr_field_get(pdamr_ctx* ctx, const char* name, void* dst);
int pdamr_update_fictitious(pdamr_ctx* ctx);     /* Grid::update_fictitious: IDW of C, rho, pressure, vel */
int pdamr_bc(pdamr_ctx* ctx, int which);         /* 0 inlet, 1 outlet, 2 wall, 3 solid, 4 wall conc., 5 wall (new buffers),
                                                    6 smooth_boundary_concentration (src/boundary.cpp:332-376) */
int pdamr_ns_compute_dt(pdamr_ctx* ctx, double* dt);
int pdamr_ns_step(pdamr_ctx* ctx, double dt);    /* compute_pressure + step -> new buffers (no swap) */
int pdamr_ns_iterate(pdamr_ctx* ctx, int iters, double dt);   /* loop bodies of solve_steady incl. the IDW update */
int pdamr_ns_solve_steady(pdamr_ctx* ctx, PdSteadyResult* out, int verbose);
int pdamr_ard_set_volume_loss(pdamr_ctx* ctx, double vl);
int pdamr_ard_compute_dt(pdamr_ctx* ctx, double* dt);
int pdamr_ard_step(pdamr_ctx* ctx, double dt);
int pdamr_ard_iterate(pdamr_ctx* ctx, int steps, double dt);
int pdamr_phase_change(pdamr_ctx* ctx, int* n_dissolved);
/* implicit ARD branch on the cloud: PD_ARD_ImplicitSolver with use_amr (src/pd_ard_implicit.cpp:22-38 per-node
 * constants, :104-346 assemble, :371-429 step, :438-487 compute_adaptive_dt, :500-531 apply_fictitious_coupling).
 * Vectors of matvec / rhs have one entry per node of the cloud (0 on nodes that are not FLUID / SOLID_MG /
 * FICTITIOUS).  precond: 0 none, 1 axial sweep.  set the volume loss with pdamr_ard_set_volume_loss. */
int pdamr_implicit_assemble(pdamr_ctx* ctx);
int pdamr_implicit_matvec(pdamr_ctx* ctx, double dt, const double* x, double* y);
int pdamr_implicit_rhs(pdamr_ctx* ctx, double dt, double* b);
int pdamr_implicit_compute_dt(pdamr_ctx* ctx, double dt_fraction, double dt_max, double* dt);
int pdamr_implicit_step(pdamr_ctx* ctx, double dt, double tol, int restart, int max_iters, int precond,
                        PdLinSolveInfo* info);
int pdamr_destroy(pdamr_ctx* ctx);

/* ---- instrumentation --------------------------------------------------------------------- */
int pdgpu_timer_start(pdgpu_ctx* ctx);               /* cudaEventRecord on the ctx stream */
int pdgpu_timer_stop(pdgpu_ctx* ctx, float* ms);     /* record + synchronize + elapsed    */
int pdgpu_launch_count(pdgpu_ctx* ctx, long long* launches, int reset);
/* Kernel selection (defaults in parentheses; every variant gives the same results, tests/test_gpu_parity.py):
 * ns_kernel (2) 0 generic / 1 block tiles / 2 z-streaming / 3 materialised CSR; ard_kernel (1) 0 / 1 / 3;
 * outlet_kernel (3); overlap (1); comm_overlap (1); graph (1); lazy_wallc (1); stream_chunk (0 = auto);
 * ns2d (1): 2D, one rank -- the NS loop bodies between two convergence polls as one persistent kernel. */
int pdgpu_set_option(pdgpu_ctx* ctx, const char* name, int value);
int pdgpu_flush_l2(pdgpu_ctx* ctx);                  /* write a > L2-sized scratch buffer */
/* kernel-only timing of the dominant kernel: average ms of `reps` launches of the NS
 * (which=0) or ARD (which=1) bond kernel alone, CUDA events on the ctx stream. */
int pdgpu_time_kernel(pdgpu_ctx* ctx, int which, int reps, float* ms_avg);
/* measured FP64 FMA peak of the device (TFLOP/s, DFMA micro-benchmark) */
int pdgpu_fp64_peak(pdgpu_ctx* ctx, double* tflops);
/* same with three distinct register sources per DFMA (register-file operand bandwidth bound) */
int pdgpu_fp64_peak3(pdgpu_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
