// common.cuh -- context, error handling and small device helpers shared by the
// translation units of libpdgpu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pdgpu.h"

// ------------------------------------------------------------------ errors ----
void pd_set_error(const char* fmt, ...);
#define PD_FAIL(...)            \
    do {                        \
        pd_set_error(__VA_ARGS__); \
        return 1;               \
    } while (0)
#define CUDA_OK(expr)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            pd_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)
#define PD_TRY(expr)            \
    do {                        \
        int _r = (expr);        \
        if (_r) return _r;      \
    } while (0)
#define CHECK_CTX(c)                                   \
    do {                                               \
        if (!(c)) PD_FAIL("null context");             \
        CUDA_OK(cudaSetDevice((c)->device));           \
    } while (0)
#define NEED_GRID(c)                                                        \
    do {                                                                    \
        CHECK_CTX(c);                                                       \
        if (!(c)->grid_built) PD_FAIL("pdgpu_grid_build has not been called"); \
    } while (0)

// ---------------------------------------------------------- stencil entry -----
// One horizon offset.  dist/evec/vol are the reference's CSR values
// (src/grid.cpp:176-183,275-288); w1 = vol/dist and w2 = vol/dist^2 are the hoisted
// bond weights; lin = local linear offset d_axial*plane + in-plane part.
struct OffEntry {
    int di, dj, dk, pad;
    long long lin;
    double dist, ex, ey, ez, vol, w1, w2;
};

// ------------------------------------------------------------------ context ---
struct pdgpu_ctx {
    PdConfig cfg;
    int dim = 0, device = 0, rank = 0, nranks = 1;
    int Nx = 0, Ny = 0, Nz = 0;     // global grid
    int Na = 0;                     // number of axial planes (Ny in 2D, Nz in 3D)
    int R = 0;                      // stencil reach == ghost width (= m_ratio)
    int a0 = 0, a1 = 0;             // owned axial planes
    int nlp = 0;                    // local planes incl. ghosts
    long long P = 0;                // nodes per plane
    long long NL = 0;               // local nodes incl. ghosts
    long long own_lo = 0, own_hi = 0;  // owned local index range
    long long halo_off[4] = {0, 0, 0, 0};  // send_lo, recv_lo, send_hi, recv_hi (pdgpu_slab_layout)
    long long N_total = 0;
    double origin[3] = {0, 0, 0};
    bool grid_built = false, fields_ready = false;

    cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr;   // main, outlet side stream, slab boundary / exchange stream
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_e = nullptr;

    // stencil
    int n_off = 0;
    std::vector<OffEntry> h_off;
    OffEntry* d_off = nullptr;

    // per-node device arrays (local indexing, ghosts included)
    uint8_t *type = nullptr, *phase = nullptr, *is_gb = nullptr, *is_precip = nullptr, *salt = nullptr;
    uint8_t* nbfast = nullptr;      // 1: FLUID node whose whole horizon is fluid-like (3D; ard_tile.cu fast path)
    double *rho[2] = {nullptr, nullptr}, *p[2] = {nullptr, nullptr}, *C[2] = {nullptr, nullptr};
    double* v[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    double* vmag = nullptr;         // |v| per fluid-like node, -1 otherwise (ARD artificial diffusion)
    double* dsol = nullptr;         // interface diffusivity of SOLID_MG nodes (0 elsewhere / salt blocked)
    double* wpack = nullptr;        // packed bond weight of the tiled ARD kernel (ard.cu: k_ard_vmag)
    int cur = 0, curC = 0;          // which buffer is "current"
    int p_input = 0;                // buffer whose p is the reference's `pressure` member

    // per-type node lists (owned nodes, local indices, ascending)
    int *l_wall = nullptr, *l_wall_mirror = nullptr, *l_inlet = nullptr, *l_outlet = nullptr,
        *l_solid = nullptr;
    long long n_wall = 0, n_inlet = 0, n_outlet = 0, n_solid = 0;
    int* l_ssolid = nullptr;        // SOLID_MG nodes with a fluid-like neighbour (subset of l_solid, ascending)
    long long n_ssolid = 0;
    // multi-GPU: WALL nodes in ghost planes + their mirrors, relative mirror offsets per node
    int *l_gwall = nullptr, *l_gwall_mirror = nullptr, *moff = nullptr;
    long long n_gwall = 0;
    double* inlet_vax = nullptr;    // prescribed axial inlet velocity per inlet-list entry
    // outlet Gauss-Seidel wavefront schedule
    int* out_nodes = nullptr;       // outlet nodes ordered by wavefront level
    int* out_level_off = nullptr;   // [n_levels+1]
    int n_levels = 0;
    int max_level_width = 0;
    // fast outlet sweep (outlet.cu)
    double *out_base_v = nullptr, *out_base_c = nullptr;
    int* out_cnt = nullptr;
    unsigned* out_mask = nullptr;
    void* out_early = nullptr;
    void* out_rows = nullptr;
    bool out_mod = false;
    int out_RJ = 1, out_n_rows = 0;
    size_t out_smem_mod = 0;
    int out_rows_doubled = 1;
    int out_rows_G = 0, out_rows_M = 0, out_row_start[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // row-walking sweep
    size_t out_smem_rows = 0;
    bool out_fast = false;
    int out_KP = 0, out_Wj = 0, out_ring = 0, out_mask_words = 0, out_tau_max = 0;
    size_t out_smem = 0;
    long long out_l0 = 0;
    long long out_l0_any = 0;       // local index of the first node of the first outlet plane (any sweep variant)
    // overlap of the outlet sweep with the bulk bond kernel: walls below the first outlet plane,
    // first local plane whose stencil touches an outlet plane (tile aligned), -1 = no overlap
    long long n_wall_lo = 0;
    // slab contexts: walls in the `reach` owned planes next to the lower / upper neighbour are
    // l_wall[0, n_wall_b0) and l_wall[n_wall_b1, n_wall) (0 / n_wall without that neighbour)
    long long n_wall_b0 = 0, n_wall_b1 = 0;
    int z_cut = -1;
    bool solids_below_cut = true;

    // reductions
    double* d_red = nullptr;        // device scratch
    double* h_red = nullptr;        // pinned host mirror
    unsigned long long* d_u64 = nullptr;
    int* d_int = nullptr;
    int* d_dissolved = nullptr;
    double* d_dissolved_rho = nullptr;   // density of the dissolved nodes before the phase change
    // nodes dissolved since the last NS step: the reference's `pressure` member still holds EOS(old density)
    // there (apply_phase_change does not touch it, src/pd_ard.cpp:193-212); snapshots reproduce that
    std::vector<int> stale_p_idx;
    std::vector<double> stale_p_rho;
    long long dissolved_cap = 0;

    // materialised CSR (optional)
    long long* csr_off = nullptr;
    int* csr_idx = nullptr;
    double *csr_dist = nullptr, *csr_evec = nullptr, *csr_vol = nullptr;
    long long nnz = -1;

    long long counts[6] = {0, 0, 0, 0, 0, 0};
    long long ns_bonds = 0, ard_bonds = 0, nnz_rows = 0;
    bool full_rows = false;         // every owned FLUID/SOLID row has the full in-box stencil

    double volume_loss = 0.0;
    // vmag cache: valid when computed from v[vmag_buf] at flow_epoch (velocities of FLUID nodes are
    // frozen during the ARD phase; only the outlet planes are refreshed per step)
    // lazy wall-concentration BC: WALL C is never read by a bond; after a device-resident ARD step the
    // BC of that step is still owed: C[wallC_src] holds the pre-step FLUID values it averages
    bool wallC_pending = false;
    int wallC_src = 0;
    long long flow_epoch = 1, vmag_epoch = 0;
    int vmag_buf = -1;
    long long launches = 0;

    // options
    int opt_ns_kernel = 2;          // 0 = generic table loop, 1 = block tiles, 2 = z-streaming (bulk copies), 3 = materialised CSR
    int opt_ard_kernel = 1;         // 0 = generic, 1 = block tiles, 3 = materialised CSR
    int opt_graph = 1;
    int opt_debug_no_halo = 0;      // skip per-step halo exchanges (timing experiments; results are wrong)
    int opt_lazy_wallc = 1;         // evaluate the wall-concentration BC only when somebody reads WALL C
    int opt_overlap = 1;            // run the outlet sweep on a side stream next to the bulk kernel
    int opt_comm_overlap = 1;       // slab contexts: boundary planes first, halo exchange next to the interior kernels
    int opt_host_step_graded = 1;   // pdgpu_step_host: thin chunks at both ends of the slab, thick ones in the middle
    int opt_outlet_rows_g = 0;      // lanes per lattice row of the row-walking sweep: 0 = default (4 if it fits), else 2, 4 or 8
    int opt_outlet_single_rows = 0; // force the single-row ring of the row-walking sweep (tests; large cross-sections use it anyway)
    int opt_outlet_kernel = 3;      // 0 = level-list kernel, 1 = level-addressed ring, 2 = lattice-addressed ring, 3 = row-walking

    // NCCL
    void* comm = nullptr;
    void* stage = nullptr;
    size_t stage_bytes = 0;
    void* l2_scratch = nullptr;
    size_t l2_scratch_bytes = 0;

    // CUDA graphs of one loop body per buffer parity
    cudaGraphExec_t g_ns[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [flow buffer][C buffer the BCs write]
    cudaGraphExec_t g_ard[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    double* d_dt = nullptr;         // device scalars: [0] ns dt, [1] ard dt, [2] decay factor
    long long g_ns_nodes[2][2] = {{0, 0}, {0, 0}}, g_ard_nodes[2][2] = {{0, 0}, {0, 0}};

    // host-array step (host_step.cu): chunk plan, copy streams, AoS staging
    long long tables_epoch = 0;     // bumped by pd_rebuild_tables and by option changes (pd_invalidate_graphs)
    long long types_epoch = 0;      // bumped by pd_rebuild_tables only (node types / lists changed)
    struct HostStep* hs = nullptr;

    // streaming / tiled bond kernels (stream.cuh): column tables, active tile list, per-device attributes
    void* tile_state = nullptr;
    void* impl_state = nullptr;     // implicit ARD branch (implicit.cu)
    void* ns2d_state = nullptr;     // persistent 2D flow loop (ns2d.cu)
    int opt_ns2d = 1;               // 2D, one rank: run batches of NS loop bodies as one persistent kernel
    int opt_stream_chunk = 0;       // planes per work item of the streaming kernels (0 = automatic)
    // double field arrays carry `field_pad` zeroed elements in front and behind (the bulk row copies
    // of the streaming kernels may start up to 3 rows + 4 elements outside the lattice box)
    size_t field_pad = 0;
    std::vector<void*> raw_fields;  // cudaMalloc bases of the padded arrays
};

// PD constants of PD_NS_Solver::init / PD_ARD_Solver::init (src/pd_ns.cpp:7-16).
struct PdConsts {
    double alpha, V_H, inv_VH, beta_lap, dens_diff_coeff, B_eos;
};
inline PdConsts pd_consts(const PdConfig& c, int dim) {
    const double PI = 3.14159265358979323846;
    PdConsts k;
    k.alpha = (double)dim;
    if (dim == 2) {
        k.V_H = PI * c.delta * c.delta;
        k.beta_lap = 4.0 / (PI * c.delta * c.delta);
    } else {
        k.V_H = (4.0 / 3.0) * PI * c.delta * c.delta * c.delta;
        k.beta_lap = 12.0 / (PI * c.delta * c.delta);
    }
    k.inv_VH = 1.0 / k.V_H;
    k.dens_diff_coeff = k.beta_lap * (c.eta_density * c.c0 * c.delta);
    k.B_eos = c.rho_f * c.c0 * c.c0 / c.gamma_eos;
    return k;
}

// ------------------------------------------------------------ launch helpers --
inline unsigned nblocks(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }
#define LAUNCH(ctx, kern, grid, block, smem, ...)                          \
    do {                                                                   \
        kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);     \
        (ctx)->launches++;                                                 \
    } while (0)

// device-side geometry of a local index
struct Lat {
    int Nx, Ny;          // in-plane extents (Ny == 1 in 2D)
    long long P;         // plane size
    int a0, R;           // owned axial start, ghost width
    int Na;              // global number of axial planes
};
inline Lat make_lat(const pdgpu_ctx* c) {
    Lat L;
    L.Nx = c->Nx;
    L.Ny = (c->dim == 3) ? c->Ny : 1;
    L.P = c->P;
    L.a0 = c->a0;
    L.R = c->R;
    L.Na = c->Na;
    return L;
}

#ifdef __CUDACC__
// Tait EOS of PD_NS_Solver::compute_pressure (src/pd_ns.cpp:36-50)
__device__ __forceinline__ double eos_pressure(double rho, double rho0, double gamma, double B) {
    double ratio = rho / rho0;
    if (ratio < 0.5) ratio = 0.5;
    if (ratio > 2.0) ratio = 2.0;
    return B * (pow(ratio, gamma) - 1.0);
}
// local index -> global lattice indices. In 2D the axial index is j.
__device__ __forceinline__ void local_to_ijk(const Lat& L, long long l, int dim, int* i, int* j, int* k,
                                             int* a_glob) {
    int al = (int)(l / L.P);
    int q = (int)(l - (long long)al * L.P);
    int a = al - L.R + L.a0;
    *a_glob = a;
    if (dim == 2) { *i = q; *j = a; *k = 0; }
    else { *j = q / L.Nx; *i = q - *j * L.Nx; *k = a; }
}

// neighbour through offset o of local node (in-plane ii,jj; local index l): local index or -1
__device__ __forceinline__ long long nbr_local(const Lat& L, const OffEntry& e, int dim, int ii, int jj,
                                               long long l, const uint8_t* __restrict__ type) {
    int ni = ii + e.di;
    if (ni < 0 || ni >= L.Nx) return -1;
    if (dim == 3) {
        int nj = jj + e.dj;
        if (nj < 0 || nj >= L.Ny) return -1;
    }
    long long nn = l + e.lin;
    return type[nn] == PDGPU_OUTSIDE ? -1 : nn;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif

// ------------------------------------------------- cross-TU internal functions
int pd_alloc_fields(pdgpu_ctx* c);
int pd_rebuild_tables(pdgpu_ctx* c);                 // lists, mirror table, bond counts
int pd_build_nbfast(pdgpu_ctx* c);                   // ard_tile.cu
int pd_enqueue_ard_solid_rows(pdgpu_ctx* c, int srcC, const double* d_dt);   // ard_tile.cu
int pd_enqueue_bc_inlet(pdgpu_ctx* c, int buf, int bufC);
int pd_enqueue_bc_outlet(pdgpu_ctx* c, int buf, int bufC);
int pd_enqueue_bc_outlet_fast(pdgpu_ctx* c, int buf, int bufC);   // outlet.cu, -1 = not applicable
int pd_outlet_setup(pdgpu_ctx* c);
int pd_enqueue_bc_outlet_prepass(pdgpu_ctx* c, int buf, int bufC);   // outlet.cu: the two halves of the fast outlet BC
int pd_enqueue_bc_outlet_sweep(pdgpu_ctx* c, int buf, int bufC);
int pd_enqueue_bc_wall_range(pdgpu_ctx* c, int buf, long long first, long long n);   // l_wall[first, first+n)
int pd_enqueue_bc_wall(pdgpu_ctx* c, int buf, int part = 0);   // part 0 all owned, 1 below the outlet planes, 2 in them, 3 ghost planes
int pd_enqueue_bc_wall_conc(pdgpu_ctx* c, int bufC, bool both_buffers = false, int srcC = -1);
int pd_flush_wall_c(pdgpu_ctx* c);   // run an owed wall-concentration BC (no-op otherwise)
int pd_enqueue_bc_solid(pdgpu_ctx* c, int buf);
int pd_enqueue_ns_step(pdgpu_ctx* c, int src, const double* d_dt, int zb = -1, int ze = -1);   // local plane range
int pd_enqueue_ard_step(pdgpu_ctx* c, int buf, int srcC, const double* d_dt);
int pd_enqueue_ard_prepass_solids(pdgpu_ctx* c, int srcC);
int pd_enqueue_ard_vmag_range(pdgpu_ctx* c, int buf, long long lo, long long hi);
int pd_ensure_vmag(pdgpu_ctx* c, int buf);
inline void pd_touch_flow(pdgpu_ctx* c) { c->flow_epoch++; }
// an NS step recomputes the pressure of every node (src/pd_ns.cpp:79): nothing stale is left
inline void pd_pressure_recomputed(pdgpu_ctx* c) { c->stale_p_idx.clear(); c->stale_p_rho.clear(); }
int pd_enqueue_ard_main(pdgpu_ctx* c, int buf, int srcC, const double* d_dt, int zb, int ze, bool do_solid,
                        bool skip_wall_copy = false);
int pd_enqueue_halo(pdgpu_ctx* c, int which, int buf, int bufC);
int pd_ns2d_prepare(pdgpu_ctx* c);                   // ns2d.cu
bool pd_ns2d_ok(const pdgpu_ctx* c);
int pd_enqueue_ns2d(pdgpu_ctx* c, int src, int iters);
int pd_enqueue_ard2d(pdgpu_ctx* c, int buf, int srcC, int steps);
int pd_enqueue_ard_vmag_range(pdgpu_ctx* c, int buf, long long lo, long long hi);
int pd_max_fluid_speed(pdgpu_ctx* c, double* vmax);
void pd_invalidate_graphs(pdgpu_ctx* c);
int pd_refresh_eos(pdgpu_ctx* c, int buf);           // p[buf] = EOS(rho[buf]) on all local nodes
int pd_enqueue_eos_range(pdgpu_ctx* c, int buf, long long lo, long long n);             // fields.cu
int pd_enqueue_eos_to(pdgpu_ctx* c, int buf, long long lo, long long n, double* out);
int pd_enqueue_deinterleave(pdgpu_ctx* c, const double* aos, long long lo, long long n, int buf);
int pd_enqueue_interleave(pdgpu_ctx* c, double* aos, long long lo, long long n, int buf);
int pd_enqueue_channel_corrections(pdgpu_ctx* c, int buf);                              // ns.cu
void pd_host_step_free(pdgpu_ctx* c);                                                   // host_step.cu
int pd_enqueue_ns_step_csr(pdgpu_ctx* c, int src, const double* d_dt);                 // csr_path.cu
int pd_enqueue_ard_step_csr(pdgpu_ctx* c, int buf, int srcC, const double* d_dt);       // csr_path.cu
int pd_set_dt(pdgpu_ctx* c, int slot, double value);
int pd_comm_allreduce(pdgpu_ctx* c, double* d_buf, int n, int op);   // op 0 sum, 1 max
int pd_comm_allreduce_bytes(pdgpu_ctx* c, void* d_buf, size_t bytes, int elem);   // in-place sum, elem = 1, 4 or 8 bytes
#define NEED_FIELDS(c)                                                                       \
    do {                                                                                     \
        NEED_GRID(c);                                                                        \
        if (!(c)->fields_ready) PD_FAIL("fields not initialised (pdgpu_fields_init/upload)"); \
    } while (0)
#define VXYZ(c, b) (c)->v[b][0], (c)->v[b][1], (c)->v[b][2]

// enqueue on another stream for the lifetime of the object (the LAUNCH macro uses c->stream)
struct StreamSwap {
    pdgpu_ctx* c;
    cudaStream_t saved;
    StreamSwap(pdgpu_ctx* ctx, cudaStream_t s) : c(ctx), saved(ctx->stream) { ctx->stream = s; }
    ~StreamSwap() { c->stream = saved; }
};
// slab contexts: compute the boundary planes first and exchange them while the interior runs
inline bool pd_comm_overlap(const pdgpu_ctx* c) {
    if (!(c->nranks > 1 && c->comm && c->opt_comm_overlap && !c->opt_debug_no_halo)) return false;
    if (c->dim != 3 || !c->full_rows || c->cfg.m_ratio != 3 || c->cfg.channel_flow_corrections) return false;
    if (c->opt_ns_kernel < 1 || c->opt_ns_kernel > 2 || c->opt_ard_kernel < 1 || c->opt_ard_kernel > 2) return false;
    if (c->a1 - c->a0 < 8 * c->R) return false;
    if (c->n_outlet > 0 && (c->z_cut < 0 || c->z_cut < 6 * c->R || !c->solids_below_cut)) return false;
    return true;
}
inline bool pd_can_overlap(const pdgpu_ctx* c) {
    return c->opt_overlap && c->z_cut > 0 && c->n_outlet > 0 && c->out_fast && c->opt_outlet_kernel > 0 &&
           c->opt_ns_kernel != 3 && c->opt_ard_kernel != 3;   // the CSR kernels take no plane range
}
