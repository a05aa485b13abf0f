// amr_implicit.cuh -- implicit ARD branch on the two-level AMR cloud (SURVEY 8(f)-2 x 8(f)-4); included at the end
// of amr.cu (same translation unit: it uses pdamr_ctx and the kernels there).
//
//   pdamr_implicit_assemble    PD_ARD_ImplicitSolver::assemble with use_amr (src/pd_ard_implicit.cpp:104-346): per-node
//                              beta / V_H from delta_local (:22-38), FICTITIOUS neighbours count as fluid (:211), the
//                              bond weights frozen per coupling cycle.  The cloud is small (10^4..10^5 nodes), so the
//                              weights ARE stored, one per CSR entry of the neighbour list + the diagonal -- the
//                              uniform-grid path (implicit.cu) is matrix-free instead.
//   pdamr_implicit_matvec/rhs  A = I - dt M on FLUID / SOLID_MG rows; FICTITIOUS rows are the IDW constraint
//                              C_f - sum_j w_j C_src(j) = sum over known sources (apply_fictitious_coupling, :500-531);
//                              b = C_old + dt bc_rhs (:352-362, :391).
//   pdamr_implicit_compute_dt  compute_adaptive_dt (:438-487).
//   pdamr_implicit_step        restarted GMRES, right-preconditioned with the axial sweep P = D + (couplings to nodes
//                              of strictly smaller axial coordinate), one CTA walking the distinct axial coordinates of
//                              the cloud in one launch; clamp to [0, C_solid_init] into the current C buffer (:409-427).
//   pdamr_bc(ctx, 6)           smooth_boundary_concentration on the cloud (src/boundary.cpp:332-376): in place in
//                              ascending node order like the reference's (single-threaded) loop -- a node reads the
//                              NEW value of an already smoothed lower-index neighbour, the OLD value otherwise.
//
// All vectors have one entry per NODE of the cloud (zeros on WALL / INLET / OUTLET / OUTSIDE nodes): no index maps.
#pragma once

struct AmrImplicit {
    double *w = nullptr, *diag = nullptr;                 // bond weights [nnz] (0 = no bond), row sums [N]
    double *b = nullptr, *x = nullptr, *r = nullptr, *t = nullptr, *z = nullptr, *V = nullptr, *cold = nullptr;
    double *part = nullptr, *h_part = nullptr;            // dot-product partials (device, pinned host)
    double* d_small = nullptr;                            // small coefficient vectors
    unsigned long long* d_min = nullptr;
    int m_cap = 0;
    int *lvl_nodes = nullptr, *d_lvl_off = nullptr;       // unknown nodes grouped by axial coordinate, ascending
    std::vector<int> lvl_off;
    int *sm_nodes = nullptr, *sm_eoff = nullptr, *sm_eidx = nullptr;   // smoother: nodes by level, their source edges
    uint8_t* sm_enew = nullptr;                           // edge reads the already smoothed value
    std::vector<int> sm_lvl_off;
    bool tables = false, assembled = false;
};

namespace {
constexpr int kAmriBlocks = 64, kAmriMaxM = 100;
constexpr bool kAmriSweepOneCta = true;       // false: one launch per axial level (measured 8 ms per GMRES iteration on the shipped cloud)

__device__ __forceinline__ bool amri_unknown(uint8_t t) { return t == T_FLUID || t == T_SOLID || t == T_FICT; }

// bond weights of one row, the reference's expression order (src/pd_ard_implicit.cpp:176-297)
__global__ void __launch_bounds__(128)
k_amri_assemble(AmrDev g, ArdK k, const double* __restrict__ vel, const uint8_t* __restrict__ is_gb,
                const uint8_t* __restrict__ is_precip, const uint8_t* __restrict__ salt, double* __restrict__ w,
                double* __restrict__ diag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.N) return;
    const uint8_t ti = g.type[i];
    const bool i_fl = ti == T_FLUID, i_so = ti == T_SOLID;
    if (!i_fl && !i_so) {
        for (int q = g.off[i]; q < g.off[i + 1]; ++q) w[q] = 0.0;
        diag[i] = 0.0;
        return;
    }
    const double d = g.delta[i];
    const double V_H = kPi * d * d, beta_i = 4.0 / (kPi * d * d), div_coeff = 2.0 / V_H;
    const double vi0 = i_fl ? vel[2 * i] : 0.0, vi1 = i_fl ? vel[2 * i + 1] : 0.0;
    double dg = 0.0;
    for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
        const int j = g.idx[q];
        const double xi = g.dist[q], Vj = g.vol[q];
        double wq = 0.0;
        const uint8_t tj = g.type[j];
        const bool j_fl = tj == T_FLUID || tj == T_INLET || tj == T_OUTLET || tj == T_FICT, j_so = tj == T_SOLID;
        if (!(Vj < 1e-30) && tj != T_WALL && tj != T_OUTSIDE && !(i_so && j_so)) {
            const double inv_xi = 1.0 / xi, inv_xi2 = inv_xi * inv_xi;
            double D_avg = 0.0;
            if (i_fl && j_fl) D_avg = k.D_liquid;
            else if ((i_fl && j_so) || (i_so && j_fl)) {
                const int si = i_so ? i : j;
                if (!salt[si]) {
                    double D_s = is_gb[si] ? k.D_gb : (is_precip[si] ? k.D_precip : k.D_grain);
                    D_s *= k.decay;
                    D_avg = 2.0 * k.D_liquid * D_s / (k.D_liquid + D_s + 1e-30);
                }
            }
            const double w_diff = beta_i * D_avg * inv_xi2 * Vj;
            wq = w_diff;
            if (i_fl && j_fl) {
                const double v_dot_e = vi0 * g.evec[2 * q] + vi1 * g.evec[2 * q + 1];
                const double w_adv = div_coeff * v_dot_e * inv_xi * Vj;
                const double w_stab = fmax(0.0, w_adv - w_diff);
                wq = (w_diff + w_stab) - w_adv;
            }
            dg -= wq;
        }
        w[q] = wq;
    }
    diag[i] = dg;
}

// y = A x
__global__ void __launch_bounds__(128)
k_amri_matvec(AmrDev g, const int* __restrict__ foff, const int* __restrict__ fsrc, const double* __restrict__ fw,
              const double* __restrict__ w, const double* __restrict__ diag, double dt, const double* __restrict__ x,
              double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.N) return;
    const uint8_t ti = g.type[i];
    double out = 0.0;
    if (ti == T_FLUID || ti == T_SOLID) {
        double s = diag[i] * x[i];
        for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
            const int j = g.idx[q];
            if (amri_unknown(g.type[j])) s += w[q] * x[j];
        }
        out = x[i] - dt * s;
    } else if (ti == T_FICT) {
        double s = 0.0;
        for (int p = foff[i]; p < foff[i + 1]; ++p) {
            const int j = fsrc[p];
            if (amri_unknown(g.type[j])) s += fw[p] * x[j];
        }
        out = x[i] - s;
    }
    y[i] = out;
}

// b = C_old + dt bc_rhs on FLUID / SOLID_MG rows, the known part of the IDW sum on FICTITIOUS rows
__global__ void __launch_bounds__(128)
k_amri_rhs(AmrDev g, const int* __restrict__ foff, const int* __restrict__ fsrc, const double* __restrict__ fw,
           const double* __restrict__ w, double dt, const double* __restrict__ C, double* __restrict__ b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.N) return;
    const uint8_t ti = g.type[i];
    double out = 0.0;
    if (ti == T_FLUID || ti == T_SOLID) {
        double s = 0.0;
        for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
            const int j = g.idx[q];
            const uint8_t tj = g.type[j];
            if (tj == T_INLET || tj == T_OUTLET) s += w[q] * C[j];
        }
        out = C[i] + dt * s;
    } else if (ti == T_FICT) {
        for (int p = foff[i]; p < foff[i + 1]; ++p) {
            const int j = fsrc[p];
            if (!amri_unknown(g.type[j])) out += fw[p] * C[j];
        }
    }
    b[i] = out;
}

// z_i of one row of z = P^-1 r, P = D + strictly-lower-axial part of A
__device__ __forceinline__ void amri_sweep_row(const AmrDev& g, int i, const double* __restrict__ pos,
                                               const int* __restrict__ foff, const int* __restrict__ fsrc,
                                               const double* __restrict__ fw, const double* __restrict__ w,
                                               const double* __restrict__ diag, double dt, const double* __restrict__ r,
                                               double* z) {
    const double yi = pos[2 * i + 1];
    double s = r[i], aii = 1.0;
    if (g.type[i] == T_FICT) {
        for (int p = foff[i]; p < foff[i + 1]; ++p) {
            const int j = fsrc[p];
            if (amri_unknown(g.type[j]) && pos[2 * j + 1] < yi) s += fw[p] * z[j];
        }
    } else {
        aii = 1.0 - dt * diag[i];
        for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
            const int j = g.idx[q];
            if (amri_unknown(g.type[j]) && pos[2 * j + 1] < yi) s += dt * w[q] * z[j];
        }
    }
    z[i] = s / aii;
}

// one axial level per launch
__global__ void __launch_bounds__(128)
k_amri_sweep(AmrDev g, const int* __restrict__ nodes, int lo, int hi, const double* __restrict__ pos,
             const int* __restrict__ foff, const int* __restrict__ fsrc, const double* __restrict__ fw,
             const double* __restrict__ w, const double* __restrict__ diag, double dt, const double* __restrict__ r,
             double* z) {
    const int n = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (n < hi) amri_sweep_row(g, nodes[n], pos, foff, fsrc, fw, w, diag, dt, r, z);
}

// the whole sweep in ONE launch: a level of the cloud is one lattice row (~10^2 nodes), so a single CTA walks the
// levels with a block barrier between them instead of one launch per level (hundreds of levels: launch bound)
__global__ void __launch_bounds__(256)
k_amri_sweep_all(AmrDev g, const int* __restrict__ nodes, const int* __restrict__ lvl_off, int nlev,
                 const double* __restrict__ pos, const int* __restrict__ foff, const int* __restrict__ fsrc,
                 const double* __restrict__ fw, const double* __restrict__ w, const double* __restrict__ diag, double dt,
                 const double* __restrict__ r, double* z) {
    for (int l = 0; l < nlev; ++l) {
        const int lo = lvl_off[l], hi = lvl_off[l + 1];
        for (int n = lo + threadIdx.x; n < hi; n += blockDim.x) amri_sweep_row(g, nodes[n], pos, foff, fsrc, fw, w, diag, dt, r, z);
        __syncthreads();                         // z of this level is read by the next ones
    }
}

// t_phase of the dissolving interface solids (src/pd_ard_implicit.cpp:453-479): minimum as the bit pattern
__global__ void __launch_bounds__(128)
k_amri_tphase(AmrDev g, const double* __restrict__ w, const double* __restrict__ diag, const double* __restrict__ C,
              double C_thresh, unsigned long long* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.N || g.type[i] != T_SOLID) return;
    const double Ci = C[i];
    if (Ci <= C_thresh) return;
    double mc = diag[i] * Ci, bc = 0.0;
    for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
        const int j = g.idx[q];
        const uint8_t tj = g.type[j];
        if (amri_unknown(tj)) mc += w[q] * C[j];
        else if (tj == T_INLET || tj == T_OUTLET) bc += w[q] * C[j];
    }
    const double dCdt = mc + bc;
    if (dCdt >= 0.0) return;
    const double rate = -dCdt;
    if (rate < 1e-30) return;
    const double tp = (Ci - C_thresh) / rate;
    if (tp > 0.0) atomicMin(out, (unsigned long long)__double_as_longlong(tp));
}

__global__ void __launch_bounds__(256) k_amri_dot(const double* __restrict__ a, const double* __restrict__ b, int n,
                                                  double* __restrict__ part) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += a[i] * b[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) s += sh[q];
        part[blockIdx.x] = s;
    }
}
__global__ void k_amri_axpy(double a, const double* __restrict__ x, int n, double* __restrict__ y) {   // y += a x
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += a * x[i];
}
__global__ void k_amri_scale_to(const double* __restrict__ x, double a, int n, double* __restrict__ y) {   // y = a x
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = a * x[i];
}
__global__ void k_amri_sub(const double* __restrict__ a, const double* __restrict__ b, int n, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = a[i] - b[i];
}
__global__ void k_amri_combine(const double* __restrict__ V, int n, int nvec, const double* __restrict__ coef,
                               double* __restrict__ y) {   // y = sum_k coef[k] V_k
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < nvec; ++k) s += coef[k] * V[(size_t)k * n + i];
    y[i] = s;
}
__global__ void k_amri_clamp_store(const uint8_t* __restrict__ type, const double* __restrict__ x, int n, double cmax,
                                   double* __restrict__ C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !amri_unknown(type[i])) return;
    double v = x[i];
    if (v < 0.0) v = 0.0;
    if (v > cmax) v = cmax;
    C[i] = v;
}

// one dependency level of the boundary smoother
__global__ void k_amri_smooth(const int* __restrict__ nodes, int lo, int hi, const int* __restrict__ eoff,
                              const int* __restrict__ eidx, const uint8_t* __restrict__ enew,
                              const double* __restrict__ cold, double* __restrict__ C) {
    const int n = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= hi) return;
    double s = 0.0;
    const int e0 = eoff[n], e1 = eoff[n + 1];
    for (int e = e0; e < e1; ++e) s += enew[e] ? C[eidx[e]] : cold[eidx[e]];
    if (e1 > e0) C[nodes[n]] = s / (double)(e1 - e0);
}

void amri_free_tables(AmrImplicit* s) {
    for (void* q : {(void*)s->lvl_nodes, (void*)s->d_lvl_off, (void*)s->sm_nodes, (void*)s->sm_eoff, (void*)s->sm_eidx,
                    (void*)s->sm_enew})
        if (q) cudaFree(q);
    s->lvl_nodes = s->d_lvl_off = s->sm_nodes = s->sm_eoff = s->sm_eidx = nullptr;
    s->sm_enew = nullptr;
    s->tables = false;
}

// host tables from the current node types: sweep levels of the unknowns, smoother nodes / edges / levels
int amri_tables(pdamr_ctx* c) {
    AmrImplicit* s = c->imp;
    amri_free_tables(s);
    const int N = c->N;
    const PdConfig& k = c->cfg;
    std::vector<int> unk;
    for (int i = 0; i < N; ++i)
        if (c->type[i] == T_FLUID || c->type[i] == T_SOLID || c->type[i] == T_FICT) unk.push_back(i);
    std::stable_sort(unk.begin(), unk.end(), [&](int a, int b) { return c->pos[2 * a + 1] < c->pos[2 * b + 1]; });
    s->lvl_off.assign(1, 0);
    for (size_t q = 1; q <= unk.size(); ++q)
        if (q == unk.size() || c->pos[2 * unk[q] + 1] != c->pos[2 * unk[q - 1] + 1]) s->lvl_off.push_back((int)q);
    PD_TRY(up(&s->lvl_nodes, unk));
    PD_TRY(up(&s->d_lvl_off, s->lvl_off));
    // smoother (src/boundary.cpp:332-376)
    const double y_min = -k.L_upstream, y_max = k.L_wire + k.L_downstream;
    std::vector<int> lev(N, -1);                   // level of a node that IS rewritten, -1 otherwise
    struct Node { int i, lev; std::vector<int> e; std::vector<uint8_t> nw; };
    std::vector<Node> nodes;
    int max_lev = -1;
    for (int i = 0; i < N; ++i) {
        if (c->type[i] != T_FLUID) continue;
        const double delta = c->deltal[i], y = c->pos[2 * i + 1];
        const bool near_in = (y - y_min < delta), near_out = (y_max - y < delta);
        if (!near_in && !near_out) continue;
        Node nd;
        nd.i = i; nd.lev = 0;
        for (int q = c->nbr_off[i]; q < c->nbr_off[i + 1]; ++q) {
            const int j = c->nbr_idx[q];
            if (c->type[j] != T_FLUID) continue;
            const double yj = c->pos[2 * j + 1];
            if ((near_out && yj < y) || (near_in && yj > y)) {       // :358-364
                const bool nw = j < i && lev[j] >= 0;        // already rewritten by the ascending loop
                nd.e.push_back(j); nd.nw.push_back(nw ? 1 : 0);
                if (nw) nd.lev = std::max(nd.lev, lev[j] + 1);
            }
        }
        if (nd.e.empty()) continue;                 // count == 0: value kept
        lev[i] = nd.lev;
        max_lev = std::max(max_lev, nd.lev);
        nodes.push_back(std::move(nd));
    }
    std::stable_sort(nodes.begin(), nodes.end(), [](const Node& a, const Node& b) { return a.lev < b.lev; });
    std::vector<int> sn, eo(1, 0), ei;
    std::vector<uint8_t> en;
    s->sm_lvl_off.assign(1, 0);
    for (size_t q = 0; q < nodes.size(); ++q) {
        if (q > 0 && nodes[q].lev != nodes[q - 1].lev) s->sm_lvl_off.push_back((int)q);
        sn.push_back(nodes[q].i);
        ei.insert(ei.end(), nodes[q].e.begin(), nodes[q].e.end());
        en.insert(en.end(), nodes[q].nw.begin(), nodes[q].nw.end());
        eo.push_back((int)ei.size());
    }
    s->sm_lvl_off.push_back((int)nodes.size());
    PD_TRY(up(&s->sm_nodes, sn)); PD_TRY(up(&s->sm_eoff, eo)); PD_TRY(up(&s->sm_eidx, ei)); PD_TRY(up(&s->sm_enew, en));
    s->tables = true;
    return 0;
}

int amri_state(pdamr_ctx* c, int m) {
    if (!c->imp) c->imp = new AmrImplicit();
    AmrImplicit* s = c->imp;
    const size_t N = (size_t)c->N, nnz = c->nbr_idx.size();
    if (!s->w) {
        CUDA_OK(cudaMalloc(&s->w, sizeof(double) * std::max<size_t>(nnz, 1)));
        for (double** q : {&s->diag, &s->b, &s->x, &s->r, &s->t, &s->z, &s->cold}) {
            CUDA_OK(cudaMalloc(q, sizeof(double) * N));
            CUDA_OK(cudaMemset(*q, 0, sizeof(double) * N));
        }
        CUDA_OK(cudaMalloc(&s->part, sizeof(double) * kAmriBlocks));
        CUDA_OK(cudaMallocHost(&s->h_part, sizeof(double) * kAmriBlocks));
        CUDA_OK(cudaMalloc(&s->d_small, sizeof(double) * (kAmriMaxM + 2)));
        CUDA_OK(cudaMalloc(&s->d_min, sizeof(unsigned long long)));
    }
    if (m > s->m_cap) {
        if (s->V) CUDA_OK(cudaFree(s->V));
        CUDA_OK(cudaMalloc(&s->V, sizeof(double) * N * (size_t)(m + 1)));
        s->m_cap = m;
    }
    if (!s->tables) PD_TRY(amri_tables(c));
    return 0;
}

int amri_dot(pdamr_ctx* c, const double* a, const double* b, double* out) {
    AmrImplicit* s = c->imp;
    k_amri_dot<<<kAmriBlocks, 256, 0, c->stream>>>(a, b, c->N, s->part);
    CUDA_OK(cudaMemcpyAsync(s->h_part, s->part, sizeof(double) * kAmriBlocks, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    double t = 0.0;
    for (int q = 0; q < kAmriBlocks; ++q) t += s->h_part[q];      // fixed order: deterministic
    *out = t;
    return 0;
}

int amri_matvec(pdamr_ctx* c, double dt, const double* x, double* y) {
    AmrImplicit* s = c->imp;
    k_amri_matvec<<<nb(c->N, 128), 128, 0, c->stream>>>(dev_view(c), c->d_foff, c->d_fsrc, c->d_fw, s->w, s->diag, dt, x, y);
    return 0;
}

int amri_precond(pdamr_ctx* c, int precond, double dt, const double* r, double* z) {
    AmrImplicit* s = c->imp;
    if (precond == 0) {
        CUDA_OK(cudaMemcpyAsync(z, r, sizeof(double) * c->N, cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    CUDA_OK(cudaMemsetAsync(z, 0, sizeof(double) * c->N, c->stream));
    AmrDev g = dev_view(c);
    if (kAmriSweepOneCta) {
        k_amri_sweep_all<<<1, 256, 0, c->stream>>>(g, s->lvl_nodes, s->d_lvl_off, (int)s->lvl_off.size() - 1, c->d_pos, c->d_foff,
                                                   c->d_fsrc, c->d_fw, s->w, s->diag, dt, r, z);
        return 0;
    }
    for (size_t l = 0; l + 1 < s->lvl_off.size(); ++l) {
        const int lo = s->lvl_off[l], hi = s->lvl_off[l + 1];
        k_amri_sweep<<<nb(hi - lo, 128), 128, 0, c->stream>>>(g, s->lvl_nodes, lo, hi, c->d_pos, c->d_foff, c->d_fsrc, c->d_fw,
                                                              s->w, s->diag, dt, r, z);
    }
    return 0;
}

#define AMRI_READY(c)                                                                                   \
    do {                                                                                                \
        AMR_DEV(c);                                                                                     \
        if (!(c)->imp || !(c)->imp->assembled) PD_FAIL("implicit operator not assembled (pdamr_implicit_assemble)"); \
    } while (0)

}   // namespace

static void amr_implicit_free(pdamr_ctx* c) {
    AmrImplicit* s = c->imp;
    if (!s) return;
    amri_free_tables(s);
    for (void* q : {(void*)s->w, (void*)s->diag, (void*)s->b, (void*)s->x, (void*)s->r, (void*)s->t, (void*)s->z, (void*)s->V,
                    (void*)s->cold, (void*)s->part, (void*)s->d_small, (void*)s->d_min})
        if (q) cudaFree(q);
    if (s->h_part) cudaFreeHost(s->h_part);
    delete s;
    c->imp = nullptr;
}
static void amr_implicit_invalidate(pdamr_ctx* c) {      // node types changed (phase change)
    if (c->imp) { c->imp->tables = false; c->imp->assembled = false; }
}

extern "C" int pdamr_implicit_assemble(pdamr_ctx* c) {
    AMR_DEV(c);
    PD_TRY(amri_state(c, 1));
    AmrImplicit* s = c->imp;
    const PdConfig& k = c->cfg;
    AmrDev g = dev_view(c);
    ArdK a;
    a.D_liquid = k.D_liquid; a.D_grain = k.D_grain; a.D_gb = k.D_gb; a.D_precip = k.D_precip;
    a.decay = k.corrosion_decay_l > 0.0 ? std::pow(10.0, -c->volume_loss / k.corrosion_decay_l) : 1.0;
    a.alpha_art = k.alpha_art_diff; a.dx = k.dx;
    k_amr_salt<<<nb(c->N, 128), 128, 0, c->stream>>>(g, c->C[c->curC], k.C_sat, c->d_salt);     // :70-89
    k_amri_assemble<<<nb(c->N, 128), 128, 0, c->stream>>>(g, a, c->vel[c->cur], c->d_gb, c->d_precip, c->d_salt, s->w, s->diag);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    s->assembled = true;
    return 0;
}

// y = A x for host vectors with one entry per node (tests)
extern "C" int pdamr_implicit_matvec(pdamr_ctx* c, double dt, const double* x_host, double* y_host) {
    AMRI_READY(c);
    AmrImplicit* s = c->imp;
    CUDA_OK(cudaMemcpy(s->x, x_host, sizeof(double) * c->N, cudaMemcpyHostToDevice));
    PD_TRY(amri_matvec(c, dt, s->x, s->t));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaMemcpy(y_host, s->t, sizeof(double) * c->N, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int pdamr_implicit_rhs(pdamr_ctx* c, double dt, double* b_host) {
    AMRI_READY(c);
    AmrImplicit* s = c->imp;
    k_amri_rhs<<<nb(c->N, 128), 128, 0, c->stream>>>(dev_view(c), c->d_foff, c->d_fsrc, c->d_fw, s->w, dt, c->C[c->curC], s->b);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaMemcpy(b_host, s->b, sizeof(double) * c->N, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int pdamr_implicit_compute_dt(pdamr_ctx* c, double dt_fraction, double dt_max, double* dt_out) {
    AMRI_READY(c);
    AmrImplicit* s = c->imp;
    const unsigned long long init = (unsigned long long)0x7FF0000000000000ull;     // +inf
    CUDA_OK(cudaMemcpyAsync(s->d_min, &init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    k_amri_tphase<<<nb(c->N, 128), 128, 0, c->stream>>>(dev_view(c), s->w, s->diag, c->C[c->curC], c->cfg.C_thresh, s->d_min);
    unsigned long long bits = 0;
    CUDA_OK(cudaMemcpyAsync(&bits, s->d_min, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    double t_min;
    std::memcpy(&t_min, &bits, sizeof(double));
    double min_t_phase = dt_max;
    if (t_min < min_t_phase) min_t_phase = t_min;
    double dt = dt_fraction * min_t_phase;
    dt = std::min(dt, dt_max);
    dt = std::max(dt, dt_max * 0.01);
    *dt_out = dt;
    return 0;
}

// PD_ARD_ImplicitSolver::step (src/pd_ard_implicit.cpp:371-429)
extern "C" int pdamr_implicit_step(pdamr_ctx* c, double dt, double tol, int restart, int max_iters, int precond,
                                   PdLinSolveInfo* info) {
    AMRI_READY(c);
    if (restart < 1 || restart > kAmriMaxM) PD_FAIL("pdamr_implicit_step: restart must be in [1, %d]", kAmriMaxM);
    if (precond < 0 || precond > 1) PD_FAIL("pdamr_implicit_step: precond 0 (none) or 1 (axial sweep)");
    PD_TRY(amri_state(c, restart));
    AmrImplicit* s = c->imp;
    const int n = c->N, m = restart;
    const unsigned gb = nb(n, 256);
    double* Cbuf = c->C[c->curC];
    k_amri_rhs<<<nb(n, 128), 128, 0, c->stream>>>(dev_view(c), c->d_foff, c->d_fsrc, c->d_fw, s->w, dt, Cbuf, s->b);
    // x0 = C_old on the unknowns, 0 elsewhere
    CUDA_OK(cudaMemsetAsync(s->x, 0, sizeof(double) * n, c->stream));
    k_amri_clamp_store<<<gb, 256, 0, c->stream>>>(c->d_type, Cbuf, n, 1e300, s->x);
    double bb = 0.0;
    PD_TRY(amri_dot(c, s->b, s->b, &bb));
    const double bnorm = std::sqrt(bb), target = tol * (bnorm > 0.0 ? bnorm : 1.0);
    std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), yv(m);
    int iters = 0, converged = 0;
    double res = 0.0;
    while (iters < max_iters) {
        PD_TRY(amri_matvec(c, dt, s->x, s->t));                        // r = b - A x
        k_amri_sub<<<gb, 256, 0, c->stream>>>(s->b, s->t, n, s->r);
        double rr = 0.0;
        PD_TRY(amri_dot(c, s->r, s->r, &rr));
        const double beta = std::sqrt(rr);
        res = beta;
        if (beta <= target) { converged = 1; break; }
        k_amri_scale_to<<<gb, 256, 0, c->stream>>>(s->r, 1.0 / beta, n, s->V);
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        int j = 0;
        for (; j < m && iters < max_iters; ++j, ++iters) {
            double* wv = s->V + (size_t)(j + 1) * n;                   // w = A P^-1 V_j, built in place in V_{j+1}
            PD_TRY(amri_precond(c, precond, dt, s->V + (size_t)j * n, s->z));
            PD_TRY(amri_matvec(c, dt, s->z, wv));
            for (int i = 0; i <= j; ++i) {                             // modified Gram-Schmidt
                double h = 0.0;
                PD_TRY(amri_dot(c, wv, s->V + (size_t)i * n, &h));
                H[(size_t)i * m + j] = h;
                k_amri_axpy<<<gb, 256, 0, c->stream>>>(-h, s->V + (size_t)i * n, n, wv);
            }
            double hh = 0.0;
            PD_TRY(amri_dot(c, wv, wv, &hh));
            const double hn = std::sqrt(hh);
            H[(size_t)(j + 1) * m + j] = hn;
            if (hn > 0.0) k_amri_scale_to<<<gb, 256, 0, c->stream>>>(wv, 1.0 / hn, n, wv);
            for (int i = 0; i < j; ++i) {                              // Givens rotations on column j
                const double a = H[(size_t)i * m + j], b2 = H[(size_t)(i + 1) * m + j];
                H[(size_t)i * m + j] = cs[i] * a + sn[i] * b2;
                H[(size_t)(i + 1) * m + j] = -sn[i] * a + cs[i] * b2;
            }
            const double a = H[(size_t)j * m + j], b2 = H[(size_t)(j + 1) * m + j], rr2 = std::hypot(a, b2);
            cs[j] = rr2 > 0.0 ? a / rr2 : 1.0;
            sn[j] = rr2 > 0.0 ? b2 / rr2 : 0.0;
            H[(size_t)j * m + j] = rr2;
            H[(size_t)(j + 1) * m + j] = 0.0;
            g[j + 1] = -sn[j] * g[j];
            g[j] = cs[j] * g[j];
            res = std::fabs(g[j + 1]);
            if (res <= target || hn == 0.0) { ++j; ++iters; break; }
        }
        const int kk = j;                                              // y = H^-1 g ; x += P^-1 (V y)
        for (int i = kk - 1; i >= 0; --i) {
            double sum = g[i];
            for (int l = i + 1; l < kk; ++l) sum -= H[(size_t)i * m + l] * yv[l];
            yv[i] = sum / H[(size_t)i * m + i];
        }
        CUDA_OK(cudaMemcpyAsync(s->d_small, yv.data(), sizeof(double) * kk, cudaMemcpyHostToDevice, c->stream));
        k_amri_combine<<<gb, 256, 0, c->stream>>>(s->V, n, kk, s->d_small, s->r);
        CUDA_OK(cudaStreamSynchronize(c->stream));                     // yv is reused
        PD_TRY(amri_precond(c, precond, dt, s->r, s->z));
        k_amri_axpy<<<gb, 256, 0, c->stream>>>(1.0, s->z, n, s->x);
    }
    PD_TRY(amri_matvec(c, dt, s->x, s->t));                            // true residual of the returned iterate
    k_amri_sub<<<gb, 256, 0, c->stream>>>(s->b, s->t, n, s->r);
    double rr = 0.0;
    PD_TRY(amri_dot(c, s->r, s->r, &rr));
    res = std::sqrt(rr);
    if (res <= target) converged = 1;
    k_amri_clamp_store<<<gb, 256, 0, c->stream>>>(c->d_type, s->x, n, c->cfg.C_solid_init, Cbuf);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    if (info) { info->iters = iters; info->converged = converged; info->rel_res = bnorm > 0.0 ? res / bnorm : res; info->pad = 0; }
    return 0;
}

// smooth_boundary_concentration (src/boundary.cpp:332-376) on the current C buffer
static int amr_smooth_conc(pdamr_ctx* c) {
    PD_TRY(amri_state(c, 1));
    AmrImplicit* s = c->imp;
    double* Cbuf = c->C[c->curC];
    CUDA_OK(cudaMemcpyAsync(s->cold, Cbuf, sizeof(double) * c->N, cudaMemcpyDeviceToDevice, c->stream));
    for (size_t l = 0; l + 1 < s->sm_lvl_off.size(); ++l) {
        const int lo = s->sm_lvl_off[l], hi = s->sm_lvl_off[l + 1];
        if (hi > lo)
            k_amri_smooth<<<nb(hi - lo, 128), 128, 0, c->stream>>>(s->sm_nodes, lo, hi, s->sm_eoff, s->sm_eidx, s->sm_enew, s->cold,
                                                                   Cbuf);
    }
    return 0;
}
