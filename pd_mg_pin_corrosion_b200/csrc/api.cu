// api.cu -- context lifetime, error reporting, instrumentation of libpdgpu.so.
#include <cstdarg>

#include "common.cuh"
#include "geom.cuh"

static thread_local char g_err[1024] = "";

void pd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* pdgpu_last_error(void) { return g_err; }
extern "C" int pdgpu_version(void) { return PDGPU_VERSION; }

extern "C" int pdgpu_device_count(int* count) {
    if (!count) PD_FAIL("pdgpu_device_count: null output");
    *count = 0;
    CUDA_OK(cudaGetDeviceCount(count));
    return 0;
}

extern "C" int pdgpu_create_slab(const PdConfig* cfg, int dim, int device, int rank, int nranks,
                                 pdgpu_ctx** out) {
    if (!cfg || !out) PD_FAIL("pdgpu_create: null argument");
    if (dim != 2 && dim != 3) PD_FAIL("pdgpu_create: dim must be 2 or 3 (PD_DIM, src/utils.h:8-12)");
    if (cfg->m_ratio < 1 || cfg->m_ratio > 5) PD_FAIL("pdgpu_create: m_ratio must be in [1,5]");
    if (!(cfg->dx > 0.0) || !(cfg->delta > 0.0)) PD_FAIL("pdgpu_create: dx/delta must be positive (run compute_derived)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        PD_FAIL("pdgpu_create: no CUDA device available (%s); libpdgpu has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) PD_FAIL("pdgpu_create: device %d out of range [0,%d)", device, ndev);
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) PD_FAIL("pdgpu_create: built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);

    pdgpu_ctx* c = new pdgpu_ctx();
    c->cfg = *cfg;
    c->dim = dim;
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    geom_extents(*cfg, dim, &c->Nx, &c->Ny, &c->Nz, c->origin);
    c->Na = (dim == 2) ? c->Ny : c->Nz;
    c->P = (dim == 2) ? c->Nx : (long long)c->Nx * c->Ny;
    c->N_total = (long long)c->Nx * c->Ny * c->Nz;
    c->R = cfg->m_ratio;
    if (pdgpu_partition_balanced(cfg, dim, nranks, rank, &c->a0, &c->a1)) { delete c; return 1; }
    if (nranks > 1 && (c->a1 - c->a0) < 2 * c->R + 2) {
        int planes = c->a1 - c->a0;
        delete c;
        PD_FAIL("pdgpu_create_slab: slab of %d planes is thinner than 2*reach+2", planes);
    }
    long long lay[10];
    if (pdgpu_slab_layout_range(c->a0, c->a1, c->P, c->R, lay)) { delete c; return 1; }
    c->nlp = (int)lay[2];
    c->NL = lay[3];
    c->own_lo = lay[4];
    c->own_hi = lay[5];
    for (int t = 0; t < 4; ++t) c->halo_off[t] = lay[6 + t];
    if (c->NL >= (1LL << 31)) {
        long long nl = c->NL;
        delete c;
        PD_FAIL("pdgpu_create: slab of %lld nodes exceeds int32 local indexing; use more ranks", nl);
    }

    // stencil (+ local linear offsets)
    int n_off = 0;
    pdgpu_stencil(cfg, dim, &n_off, nullptr, nullptr, nullptr, nullptr);
    std::vector<int> od(3 * n_off);
    std::vector<double> dist(n_off), evec((size_t)dim * n_off), vol(n_off);
    pdgpu_stencil(cfg, dim, &n_off, od.data(), dist.data(), evec.data(), vol.data());
    c->n_off = n_off;
    c->h_off.resize(n_off);
    for (int o = 0; o < n_off; ++o) {
        OffEntry& en = c->h_off[o];
        en.di = od[3 * o]; en.dj = od[3 * o + 1]; en.dk = od[3 * o + 2]; en.pad = 0;
        en.lin = (dim == 2) ? (long long)en.dj * c->P + en.di
                            : (long long)en.dk * c->P + (long long)en.dj * c->Nx + en.di;
        en.dist = dist[o];
        en.ex = evec[(size_t)dim * o]; en.ey = evec[(size_t)dim * o + 1];
        en.ez = (dim == 3) ? evec[(size_t)dim * o + 2] : 0.0;
        en.vol = vol[o];
        double inv_xi = 1.0 / en.dist;
        en.w1 = inv_xi * en.vol;
        en.w2 = inv_xi * inv_xi * en.vol;
    }
    CUDA_OK(cudaMalloc(&c->d_off, sizeof(OffEntry) * n_off));
    CUDA_OK(cudaMemcpy(c->d_off, c->h_off.data(), sizeof(OffEntry) * n_off, cudaMemcpyHostToDevice));

    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    // side stream (outlet sweep, halo exchange): highest priority, so that its few CTAs are placed
    // as soon as an SM frees up even when the bulk bond kernel was launched first
    int prio_least = 0, prio_greatest = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    CUDA_OK(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_greatest));
    CUDA_OK(cudaStreamCreateWithPriority(&c->stream3, cudaStreamNonBlocking, prio_greatest));
    CUDA_OK(cudaEventCreate(&c->ev_t0));
    CUDA_OK(cudaEventCreate(&c->ev_t1));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_c, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_d, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_e, cudaEventDisableTiming));
    CUDA_OK(cudaMalloc(&c->d_red, sizeof(double) * 8192));
    CUDA_OK(cudaMallocHost(&c->h_red, sizeof(double) * 64));
    CUDA_OK(cudaMalloc(&c->d_u64, sizeof(unsigned long long) * 16));
    CUDA_OK(cudaMalloc(&c->d_int, sizeof(int) * 16));
    CUDA_OK(cudaMalloc(&c->d_dt, sizeof(double) * 8));
    *out = c;
    return 0;
}

extern "C" int pdgpu_create(const PdConfig* cfg, int dim, int device, pdgpu_ctx** out) {
    return pdgpu_create_slab(cfg, dim, device, 0, 1, out);
}

void pd_invalidate_graphs(pdgpu_ctx* c) {
    c->tables_epoch++;   // also drops the chunk plan of pdgpu_step_host
    for (int a = 0; a < 2; ++a) {
        for (int b = 0; b < 2; ++b) {
            if (c->g_ns[a][b]) { cudaGraphExecDestroy(c->g_ns[a][b]); c->g_ns[a][b] = nullptr; }
            if (c->g_ard[a][b]) { cudaGraphExecDestroy(c->g_ard[a][b]); c->g_ard[a][b] = nullptr; }
        }
    }
}

int pd_comm_destroy(pdgpu_ctx* c);
void pd_tile_state_free(pdgpu_ctx* c);   // ns_stream.cu
void pd_implicit_free(pdgpu_ctx* c);     // implicit.cu
void pd_ns2d_free(pdgpu_ctx* c);         // ns2d.cu

extern "C" int pdgpu_destroy(pdgpu_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    pd_invalidate_graphs(c);
    pd_host_step_free(c);
    pd_comm_destroy(c);
    pd_tile_state_free(c);
    pd_implicit_free(c);
    pd_ns2d_free(c);
    for (void* p : c->raw_fields)   // padded double arrays (pd_alloc_fields)
        if (p) cudaFree(p);
    void* ptrs[] = {c->nbfast, c->d_off, c->type, c->phase, c->is_gb, c->is_precip, c->salt, c->l_wall, c->l_wall_mirror, c->l_inlet, c->l_outlet,
                    c->l_solid, c->l_ssolid, c->inlet_vax, c->out_nodes, c->out_level_off, c->d_red, c->d_u64, c->d_int,
                    c->d_dissolved, c->d_dissolved_rho, c->csr_off, c->csr_idx, c->csr_dist, c->csr_evec, c->csr_vol,
                    c->l2_scratch, c->d_dt, c->stage, c->out_base_v, c->out_base_c, c->out_cnt,
                    c->out_mask, c->out_early, c->out_rows, c->l_gwall, c->l_gwall_mirror, c->moff};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (c->h_red) cudaFreeHost(c->h_red);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_b) cudaEventDestroy(c->ev_b);
    if (c->ev_c) cudaEventDestroy(c->ev_c);
    if (c->ev_d) cudaEventDestroy(c->ev_d);
    if (c->ev_e) cudaEventDestroy(c->ev_e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream3) cudaStreamDestroy(c->stream3);
    delete c;
    return 0;
}

extern "C" int pdgpu_sync(pdgpu_ctx* c) {
    CHECK_CTX(c);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int pdgpu_timer_start(pdgpu_ctx* c) {
    CHECK_CTX(c);
    CUDA_OK(cudaEventRecord(c->ev_t0, c->stream));
    return 0;
}

extern "C" int pdgpu_timer_stop(pdgpu_ctx* c, float* ms) {
    CHECK_CTX(c);
    CUDA_OK(cudaEventRecord(c->ev_t1, c->stream));
    CUDA_OK(cudaEventSynchronize(c->ev_t1));
    if (ms) CUDA_OK(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return 0;
}

extern "C" int pdgpu_launch_count(pdgpu_ctx* c, long long* launches, int reset) {
    if (!c) PD_FAIL("null context");
    if (launches) *launches = c->launches;
    if (reset) c->launches = 0;
    return 0;
}

extern "C" int pdgpu_set_option(pdgpu_ctx* c, const char* name, int value) {
    if (!c || !name) PD_FAIL("pdgpu_set_option: null argument");
    std::string n(name);
    if (n == "ns_kernel") c->opt_ns_kernel = value;
    else if (n == "ard_kernel") c->opt_ard_kernel = value;
    else if (n == "graph") c->opt_graph = value;
    else if (n == "outlet_kernel") c->opt_outlet_kernel = value;
    else if (n == "overlap") c->opt_overlap = value;
    else if (n == "comm_overlap") c->opt_comm_overlap = value;
    else if (n == "outlet_single_rows") c->opt_outlet_single_rows = value;
    else if (n == "host_step_graded") c->opt_host_step_graded = value;
    else if (n == "outlet_rows_g") {
        c->opt_outlet_rows_g = value;
        if (c->grid_built && pd_outlet_setup(c)) return 1;
    }
    else if (n == "lazy_wallc") { if (pd_flush_wall_c(c)) return 1; c->opt_lazy_wallc = value; }
    else if (n == "debug_no_halo") c->opt_debug_no_halo = value;
    else if (n == "stream_chunk") c->opt_stream_chunk = value;
    else if (n == "ns2d") c->opt_ns2d = value;
    else PD_FAIL("pdgpu_set_option: unknown option '%s'", name);
    pd_invalidate_graphs(c);
    return 0;
}

__global__ void k_flush(double* buf, size_t n, double v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = v;
}

extern "C" int pdgpu_flush_l2(pdgpu_ctx* c) {
    CHECK_CTX(c);
    if (!c->l2_scratch) {
        c->l2_scratch_bytes = (size_t)256 << 20;   // 256 MiB > 126 MB L2
        CUDA_OK(cudaMalloc(&c->l2_scratch, c->l2_scratch_bytes));
    }
    k_flush<<<148 * 8, 256, 0, c->stream>>>((double*)c->l2_scratch, c->l2_scratch_bytes / 8, 1.0);
    return 0;
}
