// fields.cu -- device mirror of the reference's Fields (src/fields.h:7-59): SoA layout in
// HBM (rho, vx, vy, vz, p, C each contiguous, two ping-pong copies), upload/download in
// the reference's host layouts, initialize_fields (src/main.cpp:9-127) on device.
#include <algorithm>

#include "common.cuh"
#include "geom.cuh"

// double arrays: zeroed, with field_pad zeroed elements in front and behind (stream.cuh)
static int alloc_padded(pdgpu_ctx* c, double** out) {
    const size_t n = (size_t)c->NL + 2 * c->field_pad;
    double* raw = nullptr;
    CUDA_OK(cudaMalloc(&raw, sizeof(double) * n));
    CUDA_OK(cudaMemset(raw, 0, sizeof(double) * n));
    c->raw_fields.push_back(raw);
    *out = raw + c->field_pad;
    return 0;
}

int pd_alloc_fields(pdgpu_ctx* c) {
    // pad: reach rows + one staged row, rounded to 32 elements (keeps the 256-byte alignment)
    c->field_pad = ((size_t)(c->R + 1) * (size_t)c->Nx + 64 + 31) & ~(size_t)31;
    CUDA_OK(cudaMalloc(&c->type, c->NL));
    CUDA_OK(cudaMalloc(&c->phase, c->NL));
    CUDA_OK(cudaMalloc(&c->is_gb, c->NL));
    CUDA_OK(cudaMalloc(&c->is_precip, c->NL));
    CUDA_OK(cudaMalloc(&c->salt, c->NL));
    CUDA_OK(cudaMemset(c->type, PDGPU_OUTSIDE, c->NL));
    CUDA_OK(cudaMemset(c->phase, 1, c->NL));
    CUDA_OK(cudaMemset(c->is_gb, 0, c->NL));
    CUDA_OK(cudaMemset(c->is_precip, 0, c->NL));
    CUDA_OK(cudaMemset(c->salt, 0, c->NL));
    for (int b = 0; b < 2; ++b) {
        PD_TRY(alloc_padded(c, &c->rho[b]));
        PD_TRY(alloc_padded(c, &c->p[b]));
        PD_TRY(alloc_padded(c, &c->C[b]));
        for (int d = 0; d < c->dim; ++d) PD_TRY(alloc_padded(c, &c->v[b][d]));
    }
    PD_TRY(alloc_padded(c, &c->vmag));
    PD_TRY(alloc_padded(c, &c->dsol));
    PD_TRY(alloc_padded(c, &c->wpack));
    return 0;
}

// ------------------------------------------------------------------ kernels ----
template <int DIM>
__global__ void k_deinterleave(const double* __restrict__ aos, long long n, double* __restrict__ x,
                               double* __restrict__ y, double* __restrict__ z) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    x[t] = aos[t * DIM];
    y[t] = aos[t * DIM + 1];
    if (DIM == 3) z[t] = aos[t * DIM + 2];
}
template <int DIM>
__global__ void k_interleave(double* __restrict__ aos, long long n, const double* __restrict__ x,
                             const double* __restrict__ y, const double* __restrict__ z) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    aos[t * DIM] = x[t];
    aos[t * DIM + 1] = y[t];
    if (DIM == 3) aos[t * DIM + 2] = z[t];
}

__global__ void k_eos(const double* __restrict__ rho, double* __restrict__ p, long long n, double rho0,
                      double gamma, double B) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    p[t] = eos_pressure(rho[t], rho0, gamma, B);
}


// initialize_fields (src/main.cpp:9-127) for every local node, both buffers.
template <int DIM>
__global__ void k_init_fields(GeomParams g, Lat L, long long NL, PdConfig cfg, double B,
                              const uint8_t* __restrict__ type, uint8_t* __restrict__ phase,
                              double* __restrict__ rho0, double* __restrict__ rho1, double* __restrict__ p0,
                              double* __restrict__ p1, double* __restrict__ C0, double* __restrict__ C1,
                              double* __restrict__ vx0, double* __restrict__ vy0, double* __restrict__ vz0,
                              double* __restrict__ vx1, double* __restrict__ vy1, double* __restrict__ vz1) {
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= NL) return;
    int al = (int)(l / L.P);
    int q = (int)(l - (long long)al * L.P);
    int i, j;
    if (DIM == 2) { i = q; j = 0; }
    else { j = q / L.Nx; i = q - j * L.Nx; }
    uint8_t t = type[l];
    double r = cfg.rho_f, cc = 0.0, vax = 0.0;
    uint8_t ph = 1;
    switch (t) {
        case PDGPU_FLUID: cc = cfg.C_liquid_init; vax = geom_inlet_velocity(g, cfg.U_in, i, j); break;
        case PDGPU_SOLID_MG: cc = cfg.C_solid_init; ph = 0; break;
        case PDGPU_WALL: cc = 0.0; break;
        case PDGPU_INLET: cc = cfg.C_liquid_init; vax = geom_inlet_velocity(g, cfg.U_in, i, j); break;
        case PDGPU_OUTLET: cc = cfg.C_liquid_init; break;
        default: r = 0.0; cc = 0.0; break;   // OUTSIDE
    }
    double pp = eos_pressure(r, cfg.rho_f, cfg.gamma_eos, B);
    phase[l] = ph;
    rho0[l] = r; rho1[l] = r; p0[l] = pp; p1[l] = pp; C0[l] = cc; C1[l] = cc;
    vx0[l] = 0.0; vx1[l] = 0.0;
    if (DIM == 2) { vy0[l] = vax; vy1[l] = vax; }
    else { vy0[l] = 0.0; vy1[l] = 0.0; vz0[l] = vax; vz1[l] = vax; }
}

// out[t] = f[idx[t]] for nodes this context OWNS, 0 otherwise (a slab context sums the vectors of all ranks)
__global__ void k_gather(const double* __restrict__ f, const int* __restrict__ idx, long long n,
                         long long halo_shift, long long own_lo, long long own_hi, double* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    long long l = (long long)idx[t] - halo_shift;
    out[t] = (l >= own_lo && l < own_hi) ? f[l] : 0.0;
}

// --------------------------------------------------------------- host side ----
int pd_refresh_eos(pdgpu_ctx* c, int buf) {
    PdConsts k = pd_consts(c->cfg, c->dim);
    LAUNCH(c, k_eos, nblocks(c->NL, 256), 256, 0, c->rho[buf], c->p[buf], c->NL, c->cfg.rho_f, c->cfg.gamma_eos,
           k.B_eos);
    return 0;
}

int pd_enqueue_eos_range(pdgpu_ctx* c, int buf, long long lo, long long n) {
    if (n <= 0) return 0;
    PdConsts k = pd_consts(c->cfg, c->dim);
    LAUNCH(c, k_eos, nblocks(n, 256), 256, 0, c->rho[buf] + lo, c->p[buf] + lo, n, c->cfg.rho_f, c->cfg.gamma_eos,
           k.B_eos);
    return 0;
}

// out[0..n) = EOS(rho[buf][lo..lo+n)) with the reference's formula (pow), without touching the shadow p
int pd_enqueue_eos_to(pdgpu_ctx* c, int buf, long long lo, long long n, double* out) {
    if (n <= 0) return 0;
    PdConsts k = pd_consts(c->cfg, c->dim);
    LAUNCH(c, k_eos, nblocks(n, 256), 256, 0, c->rho[buf] + lo, out, n, c->cfg.rho_f, c->cfg.gamma_eos, k.B_eos);
    return 0;
}

// AoS (reference std::vector<Vec>) <-> SoA for n nodes starting at local index lo of flow buffer buf
int pd_enqueue_deinterleave(pdgpu_ctx* c, const double* aos, long long lo, long long n, int buf) {
    if (n <= 0) return 0;
    if (c->dim == 2)
        LAUNCH(c, k_deinterleave<2>, nblocks(n, 256), 256, 0, aos, n, c->v[buf][0] + lo, c->v[buf][1] + lo, nullptr);
    else
        LAUNCH(c, k_deinterleave<3>, nblocks(n, 256), 256, 0, aos, n, c->v[buf][0] + lo, c->v[buf][1] + lo,
               c->v[buf][2] + lo);
    return 0;
}
int pd_enqueue_interleave(pdgpu_ctx* c, double* aos, long long lo, long long n, int buf) {
    if (n <= 0) return 0;
    if (c->dim == 2)
        LAUNCH(c, k_interleave<2>, nblocks(n, 256), 256, 0, aos, n, c->v[buf][0] + lo, c->v[buf][1] + lo, nullptr);
    else
        LAUNCH(c, k_interleave<3>, nblocks(n, 256), 256, 0, aos, n, c->v[buf][0] + lo, c->v[buf][1] + lo,
               c->v[buf][2] + lo);
    return 0;
}


// persistent AoS<->SoA staging buffer (grown on demand; no cudaMalloc on the per-step path)
static int pd_stage(pdgpu_ctx* c, size_t bytes, double** out) {
    if (c->stage_bytes < bytes) {
        CUDA_OK(cudaStreamSynchronize(c->stream));
        if (c->stage) CUDA_OK(cudaFree(c->stage));
        c->stage = nullptr;
        c->stage_bytes = 0;
        CUDA_OK(cudaMalloc(&c->stage, bytes));
        c->stage_bytes = bytes;
    }
    *out = (double*)c->stage;
    return 0;
}

struct FieldRef {
    void* ptr[3];
    int comps;      // 1 scalar, dim for velocity
    int elem;       // bytes per element
    int flow_buf;   // >=0: flow buffer whose EOS must be refreshed after upload
};

static int field_ref(pdgpu_ctx* c, int field, FieldRef* r) {
    r->comps = 1; r->elem = 8; r->flow_buf = -1;
    r->ptr[1] = r->ptr[2] = nullptr;
    int cur = c->cur, nw = 1 - c->cur, cC = c->curC, nC = 1 - c->curC;
    switch (field) {
        case PDGPU_F_RHO: r->ptr[0] = c->rho[cur]; r->flow_buf = cur; break;
        case PDGPU_F_RHO_NEW: r->ptr[0] = c->rho[nw]; r->flow_buf = nw; break;
        case PDGPU_F_VEL: r->comps = c->dim; for (int d = 0; d < c->dim; ++d) r->ptr[d] = c->v[cur][d]; break;
        case PDGPU_F_VEL_NEW: r->comps = c->dim; for (int d = 0; d < c->dim; ++d) r->ptr[d] = c->v[nw][d]; break;
        case PDGPU_F_PRESSURE: r->ptr[0] = c->p[c->p_input]; break;
        case PDGPU_F_C: r->ptr[0] = c->C[cC]; break;
        case PDGPU_F_C_NEW: r->ptr[0] = c->C[nC]; break;
        case PDGPU_F_PHASE: r->ptr[0] = c->phase; r->elem = 1; break;
        case PDGPU_F_IS_GB: r->ptr[0] = c->is_gb; r->elem = 1; break;
        case PDGPU_F_IS_PRECIP: r->ptr[0] = c->is_precip; r->elem = 1; break;
        case PDGPU_F_NODE_TYPE: r->ptr[0] = c->type; r->elem = 1; break;
        default: PD_FAIL("unknown field id %d", field);
    }
    return 0;
}

extern "C" int pdgpu_fields_upload(pdgpu_ctx* c, int field, const void* host) {
    NEED_GRID(c);
    if (field == PDGPU_F_C || field == PDGPU_F_C_NEW) PD_TRY(pd_flush_wall_c(c));
    if (!host) PD_FAIL("pdgpu_fields_upload: null host array");
    if (field == PDGPU_F_NODE_TYPE) return pdgpu_grid_set_types(c, (const uint8_t*)host);
    if (field == PDGPU_F_PRESSURE) return 0;   // derived: p == EOS(rho) is maintained on device
    FieldRef r;
    PD_TRY(field_ref(c, field, &r));
    // global planes covered by the local array (owned + in-domain ghosts)
    int ga = std::max(c->a0 - c->R, 0), gb = std::min(c->a1 + c->R, c->Na);
    long long loff = (long long)(ga - (c->a0 - c->R)) * c->P;
    long long n = (long long)(gb - ga) * c->P;
    long long goff = (long long)ga * c->P;
    if (r.comps == 1) {
        CUDA_OK(cudaMemcpyAsync((char*)r.ptr[0] + loff * r.elem, (const char*)host + goff * r.elem, (size_t)n * r.elem,
                                cudaMemcpyHostToDevice, c->stream));
    } else {
        double* stage = nullptr;
        PD_TRY(pd_stage(c, sizeof(double) * n * r.comps, &stage));
        CUDA_OK(cudaMemcpyAsync(stage, (const double*)host + goff * r.comps, sizeof(double) * n * r.comps,
                                cudaMemcpyHostToDevice, c->stream));
        if (c->dim == 2)
            LAUNCH(c, k_deinterleave<2>, nblocks(n, 256), 256, 0, stage, n, (double*)r.ptr[0] + loff,
                   (double*)r.ptr[1] + loff, nullptr);
        else
            LAUNCH(c, k_deinterleave<3>, nblocks(n, 256), 256, 0, stage, n, (double*)r.ptr[0] + loff,
                   (double*)r.ptr[1] + loff, (double*)r.ptr[2] + loff);
    }
    if (r.flow_buf >= 0) PD_TRY(pd_refresh_eos(c, r.flow_buf));
    pd_touch_flow(c);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->fields_ready = true;
    return 0;
}

extern "C" int pdgpu_fields_download(pdgpu_ctx* c, int field, void* host) {
    NEED_GRID(c);
    if (field == PDGPU_F_C || field == PDGPU_F_C_NEW) PD_TRY(pd_flush_wall_c(c));
    if (!host) PD_FAIL("pdgpu_fields_download: null host array");
    FieldRef r;
    PD_TRY(field_ref(c, field, &r));
    // owned planes; a slab context (nranks > 1) also returns its in-domain ghost planes, which
    // hold the neighbours' boundary values after the last halo exchange
    int ga = c->a0, gb = c->a1;
    if (c->nranks > 1) { ga = std::max(c->a0 - c->R, 0); gb = std::min(c->a1 + c->R, c->Na); }
    long long n = (long long)(gb - ga) * c->P;
    long long goff = (long long)ga * c->P;
    long long lo = (long long)(ga - (c->a0 - c->R)) * c->P;
    if (r.comps == 1) {
        CUDA_OK(cudaMemcpyAsync((char*)host + goff * r.elem, (const char*)r.ptr[0] + lo * r.elem,
                                (size_t)n * r.elem, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    } else {
        double* stage = nullptr;
        PD_TRY(pd_stage(c, sizeof(double) * n * r.comps, &stage));
        if (c->dim == 2)
            LAUNCH(c, k_interleave<2>, nblocks(n, 256), 256, 0, stage, n, (const double*)r.ptr[0] + lo,
                   (const double*)r.ptr[1] + lo, nullptr);
        else
            LAUNCH(c, k_interleave<3>, nblocks(n, 256), 256, 0, stage, n, (const double*)r.ptr[0] + lo,
                   (const double*)r.ptr[1] + lo, (const double*)r.ptr[2] + lo);
        CUDA_OK(cudaMemcpyAsync((double*)host + goff * r.comps, stage, sizeof(double) * n * r.comps,
                                cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// Collective download for slab contexts: every rank receives the WHOLE global array (each rank
// contributes its owned planes to a zeroed global-size device buffer, NCCL sums them).
extern "C" int pdgpu_fields_download_all(pdgpu_ctx* c, int field, void* host) {
    NEED_GRID(c);
    if (c->nranks == 1 || !c->comm) return pdgpu_fields_download(c, field, host);
    if (field == PDGPU_F_C || field == PDGPU_F_C_NEW) PD_TRY(pd_flush_wall_c(c));
    if (!host) PD_FAIL("pdgpu_fields_download_all: null host array");
    FieldRef r;
    PD_TRY(field_ref(c, field, &r));
    const long long n_own = c->own_hi - c->own_lo, goff = (long long)c->a0 * c->P;
    const size_t per = (size_t)r.elem * r.comps, total = (size_t)c->N_total * per;
    char* g = nullptr;
    CUDA_OK(cudaMalloc(&g, total));
    CUDA_OK(cudaMemsetAsync(g, 0, total, c->stream));
    if (r.comps == 1) {
        CUDA_OK(cudaMemcpyAsync(g + goff * r.elem, (const char*)r.ptr[0] + c->own_lo * r.elem, (size_t)n_own * r.elem,
                                cudaMemcpyDeviceToDevice, c->stream));
    } else {
        double* dst = (double*)g + goff * r.comps;
        const long long lo = c->own_lo;
        if (c->dim == 2)
            LAUNCH(c, k_interleave<2>, nblocks(n_own, 256), 256, 0, dst, n_own, (const double*)r.ptr[0] + lo,
                   (const double*)r.ptr[1] + lo, nullptr);
        else
            LAUNCH(c, k_interleave<3>, nblocks(n_own, 256), 256, 0, dst, n_own, (const double*)r.ptr[0] + lo,
                   (const double*)r.ptr[1] + lo, (const double*)r.ptr[2] + lo);
    }
    int rc = pd_comm_allreduce_bytes(c, g, total, r.elem);
    if (!rc && cudaMemcpyAsync(host, g, total, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = 1;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = 1;
    cudaFree(g);
    if (rc) PD_FAIL("pdgpu_fields_download_all failed: %s", pdgpu_last_error());
    return 0;
}

extern "C" int pdgpu_fields_init(pdgpu_ctx* c, const uint8_t* is_gb, const uint8_t* is_precip) {
    NEED_GRID(c);
    c->cur = 0; c->curC = 0; c->p_input = 0;
    c->wallC_pending = false; c->wallC_src = 0;   // a reused context starts like a fresh one
    c->volume_loss = 0.0;
    pd_pressure_recomputed(c);
    pd_touch_flow(c);
    pd_invalidate_graphs(c);
    c->fields_ready = true;   // allow the flag uploads below
    if (is_gb) PD_TRY(pdgpu_fields_upload(c, PDGPU_F_IS_GB, is_gb));
    else CUDA_OK(cudaMemsetAsync(c->is_gb, 0, c->NL, c->stream));
    if (is_precip) PD_TRY(pdgpu_fields_upload(c, PDGPU_F_IS_PRECIP, is_precip));
    else CUDA_OK(cudaMemsetAsync(c->is_precip, 0, c->NL, c->stream));
    Lat L = make_lat(c);
    double org[3] = {c->origin[0], c->origin[1], c->origin[2]};
    GeomParams g = geom_params(c->cfg, c->dim, org);
    PdConsts k = pd_consts(c->cfg, c->dim);
    if (c->dim == 2)
        LAUNCH(c, k_init_fields<2>, nblocks(c->NL, 256), 256, 0, g, L, c->NL, c->cfg, k.B_eos, c->type, c->phase,
               c->rho[0], c->rho[1], c->p[0], c->p[1], c->C[0], c->C[1], c->v[0][0], c->v[0][1], nullptr,
               c->v[1][0], c->v[1][1], nullptr);
    else
        LAUNCH(c, k_init_fields<3>, nblocks(c->NL, 256), 256, 0, g, L, c->NL, c->cfg, k.B_eos, c->type, c->phase,
               c->rho[0], c->rho[1], c->p[0], c->p[1], c->C[0], c->C[1], c->v[0][0], c->v[0][1], c->v[0][2],
               c->v[1][0], c->v[1][1], c->v[1][2]);
    pd_touch_flow(c);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int pdgpu_swap_flow(pdgpu_ctx* c) {
    CHECK_CTX(c);
    pd_touch_flow(c);
    c->cur = 1 - c->cur;
    return 0;
}
extern "C" int pdgpu_swap_C(pdgpu_ctx* c) {
    CHECK_CTX(c);
    PD_TRY(pd_flush_wall_c(c));
    c->curC = 1 - c->curC;
    return 0;
}

extern "C" int pdgpu_gather(pdgpu_ctx* c, int field, const int* idx, long long n, double* out) {
    NEED_GRID(c);
    if (field == PDGPU_F_C || field == PDGPU_F_C_NEW) PD_TRY(pd_flush_wall_c(c));
    if (n <= 0) return 0;
    FieldRef r;
    PD_TRY(field_ref(c, field, &r));
    if (r.comps != 1 || r.elem != 8) PD_FAIL("pdgpu_gather: scalar double fields only");
    int* d_idx = nullptr;
    double* d_out = nullptr;
    CUDA_OK(cudaMalloc(&d_idx, sizeof(int) * n));
    CUDA_OK(cudaMalloc(&d_out, sizeof(double) * n));
    CUDA_OK(cudaMemcpyAsync(d_idx, idx, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    long long halo_shift = (long long)(c->a0 - c->R) * c->P;
    LAUNCH(c, k_gather, nblocks(n, 256), 256, 0, (const double*)r.ptr[0], d_idx, n, halo_shift, c->own_lo, c->own_hi,
           d_out);
    // slab contexts: collective -- every rank passes the same index list and receives every value
    // (each node has exactly one owner; the merge is an integer sum of the bit patterns)
    if (c->nranks > 1 && c->comm) PD_TRY(pd_comm_allreduce_bytes(c, d_out, sizeof(double) * (size_t)n, 8));
    CUDA_OK(cudaMemcpyAsync(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d_idx));
    CUDA_OK(cudaFree(d_out));
    return 0;
}
