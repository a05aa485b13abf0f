// outlet.cu -- fast exact implementation of apply_outlet_bc (reference src/boundary.cpp:88-131).
//
// The reference sweep is an in-place Gauss-Seidel pass in index order: OUTLET node n
// becomes the mean of its FLUID and OUTLET neighbours, seeing NEW values of OUTLET
// neighbours with a smaller index and OLD values of those with a larger index.  Split:
//
//   pre-pass (fully parallel, all SMs): per outlet node, count of FLUID|OUTLET neighbours
//       and base = sum over FLUID neighbours + sum over lexicographically LATER outlet
//       neighbours (old values); also writes rho = rho_f, p = 0, transverse velocity = 0.
//   sweep (sequential part only): new[n] = (base[n] + sum over EARLIER outlet neighbours
//       of new[.]) / count, processed by hyperplane levels tau = i + B j + B^2 k'
//       (B = reach+1): nodes of one level are independent and every earlier neighbour has
//       a smaller tau.  The last RING levels live in a shared-memory
//       ring addressed arithmetically (no adjacency lists, no global loads on the
//       dependent path); the axial-velocity sweep and the concentration sweep are
//       independent systems and run as two CTAs on two SMs.
//
// Levels are enumerated over ALL in-box lattice nodes of the outlet planes; non-outlet
// nodes store 0 in the ring, so the sweep needs no type test per neighbour.
#include <algorithm>

#include "common.cuh"

namespace {

struct OutletGeom {
    int Nx, Ny;        // in-plane extents (Ny = 1 in 2D)
    int KP;            // number of outlet planes
    int Wj;            // max nodes per (level, plane)
    int B, B2;         // tau bases
    int ring;          // ring depth (power of two > max tau distance)
    int n_early;       // lexicographically earlier offsets = first half of the stencil
    int tau_max;
    long long P;       // nodes per plane
    long long l0;      // local index of the first node of the first outlet plane
};

__device__ __forceinline__ int ceil_div_pos(int a, int b) { return a <= 0 ? 0 : (a + b - 1) / b; }

// tot / n for a neighbour count n without the division sequence on the level's dependent path: q = RN(tot * RN(1/n)) is
// within one ulp, the exact remainder comes from one FMA, RN(q + r * RN(1/n)) is the correctly rounded quotient (Markstein;
// checked against the division for n <= 130 on 4 x 10^8 operands); totals below the range of the correction take the division.
__device__ __noinline__ double div_slow(double a, double dn) { return a / dn; }
__device__ __forceinline__ double div_by_count(double tot, int n, double rcp) {
    const double dn = (double)n, q = tot * rcp;
    if (fabs(tot) < 1e-280 && tot != 0.0) return div_slow(tot, dn);
    return fma(fma(-q, dn, tot), rcp, q);
}

// node p of level tau: returns false if there is none
__device__ __forceinline__ bool level_node(const OutletGeom& g, int tau, int p, int* kp, int* j, int* i, int* jj) {
    *kp = p / g.Wj;
    *jj = p - *kp * g.Wj;
    int c = tau - g.B2 * *kp;
    if (c < 0) return false;
    int jlo = ceil_div_pos(c - (g.Nx - 1), g.B);
    *j = jlo + *jj;
    *i = c - g.B * *j;
    return *j < g.Ny && *i >= 0;
}

template <int DIM>
__global__ void __launch_bounds__(128)
k_outlet_prepass(Lat L, OutletGeom g, const int* __restrict__ list, long long n, const uint8_t* __restrict__ type,
                 const OffEntry* __restrict__ off, int n_off, double* __restrict__ rho, double* __restrict__ p,
                 double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
                 const double* __restrict__ C, double rho_f, double* __restrict__ base_v,
                 double* __restrict__ base_c, int* __restrict__ cnt) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    long long l = list[t];
    const double* vax = (DIM == 2) ? vy : vz;
    int q = (int)(l % L.P);
    int jj = (DIM == 3) ? q / L.Nx : 0;
    int ii = q - jj * L.Nx;
    // sums in CSR order (as the reference); the loads of a batch of neighbours are issued together
    double sv = 0.0, sc = 0.0;
    int c = 0;
    constexpr int UB = 8;
    for (int o0 = 0; o0 < n_off; o0 += UB) {
        double rv[UB], rc[UB];
        int kind[UB];   // 0 skip, 1 value counted now, 2 earlier OUTLET neighbour (value added by the sweep)
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int o = o0 + u;
            long long nn = -1;
            if (o < n_off) {
                const OffEntry e = off[o];
                const int ni = ii + e.di;
                bool in = ni >= 0 && ni < L.Nx;
                if (DIM == 3) { const int nj = jj + e.dj; in = in && nj >= 0 && nj < L.Ny; }
                if (in) nn = l + e.lin;
            }
            const long long safe = nn >= 0 ? nn : l;
            const uint8_t tj = type[safe];
            kind[u] = 0;
            if (nn >= 0) {
                if (tj == PDGPU_FLUID || (tj == PDGPU_OUTLET && o >= g.n_early)) kind[u] = 1;
                else if (tj == PDGPU_OUTLET) kind[u] = 2;
            }
            rv[u] = vax[safe];
            rc[u] = C[safe];
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            if (kind[u] == 1) { sv += rv[u]; sc += rc[u]; ++c; }
            else if (kind[u] == 2) ++c;
        }
    }
    long long d = l - g.l0;
    base_v[d] = sv;
    base_c[d] = sc;
    cnt[d] = c;
    rho[l] = rho_f;
    p[l] = 0.0;          // EOS(rho_f) = 0 exactly
    vx[l] = 0.0;
    if (DIM == 3) vy[l] = 0.0;
}

// blockIdx.x = 0: axial velocity, 1: concentration
__global__ void __launch_bounds__(1024, 1)
k_outlet_sweep(OutletGeom g, const int4* __restrict__ early, const unsigned* __restrict__ mask_g, int mask_words,
               const double* __restrict__ base_v, const double* __restrict__ base_c, const int* __restrict__ cnt,
               double* vax, double* C, double U_in) {
    extern __shared__ double smem[];
    double* ring = smem;
    int4* s_early = (int4*)(ring + (size_t)g.ring * g.KP * g.Wj);
    unsigned* s_mask = (unsigned*)(s_early + g.n_early);
    const bool is_vel = (blockIdx.x == 0);
    const double* base = is_vel ? base_v : base_c;
    double* out = is_vel ? vax : C;
    for (int e = threadIdx.x; e < g.n_early; e += blockDim.x) s_early[e] = early[e];
    for (int w = threadIdx.x; w < mask_words; w += blockDim.x) s_mask[w] = mask_g[w];
    __syncthreads();

    constexpr int G = 8;
    const int lane = threadIdx.x & (G - 1);
    const int group = threadIdx.x / G;
    const int NG = blockDim.x / G;
    const int npairs = g.KP * g.Wj;
    const int rmask = g.ring - 1;
    constexpr int MAXP = 4;   // pairs per group per level (npairs <= MAXP * NG is checked on the host)

    // software-pipelined loads of (base, cnt) for the next level
    double nb[MAXP];
    int nc[MAXP];
#pragma unroll
    for (int s = 0; s < MAXP; ++s) {
        nb[s] = 0.0; nc[s] = 0;
        int p = group + s * NG, kp, j, i, jj;
        if (p < npairs && level_node(g, 0, p, &kp, &j, &i, &jj)) {
            long long d = (long long)kp * g.P + (long long)j * g.Nx + i;
            nb[s] = base[d]; nc[s] = cnt[d];
        }
    }
    for (int tau = 0; tau <= g.tau_max; ++tau) {
        double cb[MAXP];
        int cc[MAXP];
#pragma unroll
        for (int s = 0; s < MAXP; ++s) { cb[s] = nb[s]; cc[s] = nc[s]; }
        if (tau < g.tau_max) {
#pragma unroll
            for (int s = 0; s < MAXP; ++s) {
                int p = group + s * NG, kp, j, i, jj;
                nb[s] = 0.0; nc[s] = 0;
                if (p < npairs && level_node(g, tau + 1, p, &kp, &j, &i, &jj)) {
                    long long d = (long long)kp * g.P + (long long)j * g.Nx + i;
                    nb[s] = base[d]; nc[s] = cnt[d];
                }
            }
        }
#pragma unroll
        for (int s = 0; s < MAXP; ++s) {
            int p = group + s * NG, kp = 0, j = 0, i = 0, jj = 0;
            bool valid = (p < npairs) && level_node(g, tau, p, &kp, &j, &i, &jj);
            bool is_out = false;
            long long d = 0;
            if (valid) {
                d = (long long)kp * g.P + (long long)j * g.Nx + i;
                is_out = (s_mask[d >> 5] >> (d & 31)) & 1u;
            }
            double sum = 0.0;
            if (is_out) {
                for (int e = lane; e < g.n_early; e += G) {
                    int4 o = s_early[e];                 // di, dj (in-plane), dplane, dtau
                    int k2 = kp + o.z;
                    if (k2 < 0) continue;                // FLUID plane: in the pre-pass
                    int j2 = j + o.y, i2 = i + o.x;
                    if (j2 < 0 || j2 >= g.Ny || i2 < 0 || i2 >= g.Nx) continue;
                    int tau2 = tau + o.w;
                    int c2 = tau2 - g.B2 * k2;
                    int jlo2 = ceil_div_pos(c2 - (g.Nx - 1), g.B);
                    sum += ring[((size_t)(tau2 & rmask) * g.KP + k2) * g.Wj + (j2 - jlo2)];
                }
            }
            // reduce over the 8 lanes of the group (aligned 8-lane segments of a warp)
            sum += __shfl_xor_sync(0xffffffffu, sum, 4);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            if (valid && lane == 0) {
                double val = 0.0;
                if (is_out) {
                    double tot = cb[s] + sum;
                    int n = cc[s];
                    if (is_vel) val = n > 0 ? tot * (1.0 / n) : U_in;   // src/boundary.cpp:113-124
                    else val = n > 0 ? tot / n : 0.0;                   // :129
                    out[g.l0 + d] = val;
                }
                ring[((size_t)(tau & rmask) * g.KP + kp) * g.Wj + jj] = val;
            }
        }
        __syncthreads();
    }
}

// Variant with a lattice-addressed ring: slot(i,j,k') = (k'*RJ + (j mod RJ))*RI + (i mod RI).
// Inside the active window of RI = 64 levels a plane row holds at most 64 consecutive i and
// the window covers at most RJ = 64 rows (needs (maxd + Nx - 1)/B + 1 <= 64), so the mapping is
// collision free and a neighbour's slot is affine in (di,dj,dk): the earlier half of the
// stencil is walked as (dj,dk) rows with contiguous di ranges, ~6 instructions per bond.
__global__ void __launch_bounds__(1024, 1)
k_outlet_sweep_mod(OutletGeom g, int RJ, const int4* __restrict__ rows, int n_rows,
                   const unsigned* __restrict__ mask_g, int mask_words, const double* __restrict__ base_v,
                   const double* __restrict__ base_c, const int* __restrict__ cnt, double* vax, double* C,
                   double U_in) {
    extern __shared__ double smem[];
    double* ring = smem;
    const int RI = g.ring;
    int4* s_rows = (int4*)(ring + (size_t)g.KP * RJ * RI);
    unsigned* s_mask = (unsigned*)(s_rows + n_rows);
    const bool is_vel = (blockIdx.x == 0);
    const double* base = is_vel ? base_v : base_c;
    double* out = is_vel ? vax : C;
    for (int e = threadIdx.x; e < n_rows; e += blockDim.x) s_rows[e] = rows[e];
    for (int w = threadIdx.x; w < mask_words; w += blockDim.x) s_mask[w] = mask_g[w];
    __syncthreads();

    constexpr int G = 4;
    const int lane = threadIdx.x & (G - 1);
    const int group = threadIdx.x / G;
    const int NG = blockDim.x / G;
    const int npairs = g.KP * g.Wj;
    const int imask = RI - 1, jmask = RJ - 1;
    constexpr int MAXP = 2;

    double nb[MAXP];
    int nc[MAXP];
#pragma unroll
    for (int s = 0; s < MAXP; ++s) {
        nb[s] = 0.0; nc[s] = 0;
        int p = group + s * NG, kp, j, i, jj;
        if (p < npairs && level_node(g, 0, p, &kp, &j, &i, &jj)) {
            long long d = (long long)kp * g.P + (long long)j * g.Nx + i;
            nb[s] = base[d]; nc[s] = cnt[d];
        }
    }
    for (int tau = 0; tau <= g.tau_max; ++tau) {
        double cb[MAXP];
        int cc[MAXP];
#pragma unroll
        for (int s = 0; s < MAXP; ++s) { cb[s] = nb[s]; cc[s] = nc[s]; }
        if (tau < g.tau_max) {
#pragma unroll
            for (int s = 0; s < MAXP; ++s) {
                int p = group + s * NG, kp, j, i, jj;
                nb[s] = 0.0; nc[s] = 0;
                if (p < npairs && level_node(g, tau + 1, p, &kp, &j, &i, &jj)) {
                    long long d = (long long)kp * g.P + (long long)j * g.Nx + i;
                    nb[s] = base[d]; nc[s] = cnt[d];
                }
            }
        }
#pragma unroll
        for (int s = 0; s < MAXP; ++s) {
            int p = group + s * NG, kp = 0, j = 0, i = 0, jj = 0;
            if (s > 0 && s * NG >= npairs) break;          // uniform: no second pass needed
            bool valid = (p < npairs) && level_node(g, tau, p, &kp, &j, &i, &jj);
            bool is_out = false;
            long long d = 0;
            if (valid) {
                d = (long long)kp * g.P + (long long)j * g.Nx + i;
                is_out = (s_mask[d >> 5] >> (d & 31)) & 1u;
            }
            double sum = 0.0;
            if (is_out) {
                for (int r = lane; r < n_rows; r += G) {
                    const int4 row = s_rows[r];              // dj, dplane, di_lo, di_hi
                    const int k2 = kp + row.y;
                    if (k2 < 0) continue;                    // FLUID plane: in the pre-pass
                    const int j2 = j + row.x;
                    if ((unsigned)j2 >= (unsigned)g.Ny) continue;
                    const int rb = (k2 * RJ + (j2 & jmask)) * RI;
                    const int lo = max(i + row.z, 0), hi = min(i + row.w, g.Nx - 1);
                    for (int i2 = lo; i2 <= hi; ++i2) sum += ring[rb + (i2 & imask)];
                }
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            if (valid && lane == 0) {
                double val = 0.0;
                if (is_out) {
                    double tot = cb[s] + sum;
                    int n = cc[s];
                    if (is_vel) val = n > 0 ? tot * (1.0 / n) : U_in;   // src/boundary.cpp:113-124
                    else val = n > 0 ? tot / n : 0.0;                   // :129
                    out[g.l0 + d] = val;
                }
                ring[(kp * RJ + (j & jmask)) * RI + (i & imask)] = val;
            }
        }
        __syncthreads();
    }
}

// Row-walking variant (default): a group of G lanes owns one lattice row (k', j) of the outlet
// planes at a time and follows the node i = tau - B^2 k' - B j that the level front cuts out of it;
// the row's successor for the same group is row j + M, which the front reaches after this row has
// left it (B M >= Nx + PF + R).  Per level nothing is divided or looked up: i advances by one, base
// and neighbour count of the next G nodes of the row are prefetched one block ahead (coalesced,
// handed to the level that needs them by a shuffle; count -1 marks a non-OUTLET lattice node) and
// 1/n comes from a table.  The ring is lattice addressed as in k_outlet_sweep_mod, but every ring
// row is stored twice (slots s and s + RI) and the owner also writes zeros for the R virtual nodes
// before and after its row, so a (dj,dplane) row of the earlier half of the stencil is a run of
// <= 2R+1 consecutive shared-memory words: no wrap, no clipping, 3 instructions per bond.
struct RowSweepParams {
    OutletGeom g;
    int RJ, n_rows, M, R;
    int doubled;           // ring rows stored twice (no wrap inside a run); 0 = single rows, wrapped reads
    int row_start[8];      // first table row with k' + dplane >= 0, per outlet plane k'
    int n_rcp;             // reciprocal table entries (stencil size + 1)
};

template <int G>
__global__ void __launch_bounds__(1024, 1)
k_outlet_sweep_rows(const __grid_constant__ RowSweepParams q, const int4* __restrict__ rows,
                    const double* __restrict__ base_v, const double* __restrict__ base_c,
                    const int* __restrict__ cnt, double* vax, double* C, double U_in) {
    extern __shared__ double smem[];
    const OutletGeom& g = q.g;
    double* ring = smem;
    const int RI = g.ring, RJ = q.RJ, RW = q.doubled ? 2 * RI : RI;   // RW = ring row pitch
    int4* s_rows = (int4*)(ring + (size_t)g.KP * RJ * RW);   // even number of doubles: 16-byte aligned
    double* s_rcp = (double*)(s_rows + q.n_rows);
    const bool is_vel = (blockIdx.x == 0);
    const double* base = is_vel ? base_v : base_c;
    double* out = is_vel ? vax : C;
    for (int e = threadIdx.x; e < q.n_rows; e += blockDim.x) s_rows[e] = rows[e];
    for (int e = threadIdx.x; e < q.n_rcp; e += blockDim.x) s_rcp[e] = e > 0 ? 1.0 / (double)e : 0.0;
    for (int e = threadIdx.x; e < g.KP * RJ * RW; e += blockDim.x) ring[e] = 0.0;
    __syncthreads();

    constexpr int PF = G;                       // prefetch block = one value per lane
    const int lane = threadIdx.x & (G - 1);
    const int slot = threadIdx.x / G;
    const int kp = slot / q.M, js = slot - kp * q.M;
    const bool slot_ok = kp < g.KP;
    const int imask = RI - 1, jmask = RJ - 1;
    const int BM = g.B * q.M;
    const int r0 = slot_ok ? q.row_start[kp] : q.n_rows;
    const int i_end = g.Nx + q.R;               // one past the last (virtual) node of a row

    int j = js;
    int i = -g.B2 * kp - g.B * js;              // position of the level front in row j at tau = 0
    double cur_b = 0.0, nxt_b = 0.0;
    int cur_c = -1, nxt_c = -1;
    auto load_block = [&](int i_first, double* b, int* c) {
        const int ib = i_first + lane;
        *b = 0.0; *c = -1;
        if (ib >= 0 && ib < g.Nx) {
            const long long d = (long long)kp * g.P + (long long)j * g.Nx + ib;
            *b = base[d]; *c = cnt[d];
        }
    };
    if (slot_ok && j < g.Ny && i > -PF) {       // rows the front has already entered at tau = 0
        const int blk = (i >= 0 ? i / PF : -((-i + PF - 1) / PF)) * PF;
        if (blk == i) load_block(blk, &nxt_b, &nxt_c);
        else { load_block(blk, &cur_b, &cur_c); load_block(blk + PF, &nxt_b, &nxt_c); }
    }

    for (int tau = 0; tau <= g.tau_max + q.R; ++tau) {
        const bool row_ok = slot_ok && j < g.Ny;
        if (row_ok && i >= -PF && i < g.Nx && (i & (PF - 1)) == 0) {
            cur_b = nxt_b; cur_c = nxt_c;
            load_block(i + PF, &nxt_b, &nxt_c);
        }
        const double b = __shfl_sync(0xffffffffu, cur_b, i & (PF - 1), G);
        int n = __shfl_sync(0xffffffffu, cur_c, i & (PF - 1), G);
        const bool in_row = row_ok && i >= 0 && i < g.Nx;
        if (!in_row) n = -1;
        double s0 = 0.0, s1 = 0.0;
        if (n >= 0) {
            for (int r = r0 + lane; r < q.n_rows; r += G) {
                const int4 row = s_rows[r];                  // dj, dplane, di_lo, di_hi
                const int j2 = j + row.x;
                if ((unsigned)j2 >= (unsigned)g.Ny) continue;
                const double* rowp = ring + ((kp + row.y) * RJ + (j2 & jmask)) * RW;
                const int s_first = (i + row.z) & imask;
                const int w = row.w - row.z + 1;
                if (q.doubled) {
                    const double* src = rowp + s_first;
#pragma unroll
                    for (int u = 0; u < 7; ++u) {            // a row of the reach-3 sphere has <= 7 nodes
                        const double v = (u < w) ? src[u] : 0.0;
                        if (u & 1) s1 += v; else s0 += v;
                    }
                    for (int u = 7; u < w; ++u) s0 += src[u];    // reach > 3
                } else {                                     // large cross-sections: single rows, wrapped
#pragma unroll
                    for (int u = 0; u < 7; ++u) {
                        const double v = (u < w) ? rowp[(s_first + u) & imask] : 0.0;
                        if (u & 1) s1 += v; else s0 += v;
                    }
                    for (int u = 7; u < w; ++u) s0 += rowp[(s_first + u) & imask];
                }
            }
        }
        double sum = s0 + s1;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (row_ok && i >= -q.R && i < i_end && lane == 0) {   // virtual nodes beside the row store 0
            double val = 0.0;
            if (n >= 0) {
                const double tot = b + sum;
                if (is_vel) val = n > 0 ? tot * s_rcp[n] : U_in;   // src/boundary.cpp:113-124
                else val = n > 0 ? div_by_count(tot, n, s_rcp[n]) : 0.0;   // :129 (sc / cnt, correctly rounded)
                out[g.l0 + (long long)kp * g.P + (long long)j * g.Nx + i] = val;
            }
            double* dst = ring + (kp * RJ + (j & jmask)) * RW + (i & imask);
            dst[0] = val;
            if (q.doubled) dst[RI] = val;
        }
        __syncthreads();
        ++i;
        if (i >= i_end && row_ok) { j += q.M; i -= BM; }
    }
}

__global__ void k_outlet_mask(const uint8_t* __restrict__ type, long long l0, long long n, unsigned* __restrict__ mask) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w * 32 >= n) return;
    unsigned m = 0;
    for (int b = 0; b < 32; ++b) {
        long long d = w * 32 + b;
        if (d < n && type[l0 + d] == PDGPU_OUTLET) m |= (1u << b);
    }
    mask[w] = m;
}

}  // namespace

// Build the static data of the fast sweep (called from pd_rebuild_tables).
int pd_outlet_setup(pdgpu_ctx* c) {
    cudaFree(c->out_base_v); cudaFree(c->out_base_c); cudaFree(c->out_cnt); cudaFree(c->out_mask);
    cudaFree(c->out_early); cudaFree(c->out_rows);
    c->out_rows = nullptr;
    c->out_base_v = c->out_base_c = nullptr; c->out_cnt = nullptr; c->out_mask = nullptr; c->out_early = nullptr;
    c->out_fast = false;
    if (c->n_outlet == 0) return 0;
    std::vector<int> nodes(c->n_outlet);
    CUDA_OK(cudaMemcpy(nodes.data(), c->l_outlet, sizeof(int) * c->n_outlet, cudaMemcpyDeviceToHost));
    long long al_min = nodes.front() / c->P, al_max = nodes.back() / c->P;
    int KP = (int)(al_max - al_min + 1);
    int Nx = c->Nx, Ny = (c->dim == 3) ? c->Ny : 1;
    int B = c->R + 1, B2 = B * B;
    int maxd = c->R * (1 + B + B2);
    int ring = 1;
    while (ring <= maxd) ring <<= 1;
    int Wj = std::min(Ny, (Nx - 1) / B + 1);
    int n_early = c->n_off / 2;
    long long nslab = (long long)KP * c->P;
    int mask_words = (int)((nslab + 31) / 32);
    size_t smem = sizeof(double) * (size_t)ring * KP * Wj + sizeof(int4) * n_early + sizeof(unsigned) * mask_words;
    if (KP * Wj > 4 * (1024 / 8) || smem > 220 * 1024) return 0;   // fall back to the level-list kernel
    // the stencil is symmetric and lexicographically ordered: first half = earlier neighbours
    std::vector<int4> early(n_early);
    for (int o = 0; o < n_early; ++o) {
        const OffEntry& e = c->h_off[o];
        int di = e.di, dj = (c->dim == 3) ? e.dj : 0, dp = (c->dim == 3) ? e.dk : e.dj;
        if (!(dp < 0 || (dp == 0 && dj < 0) || (dp == 0 && dj == 0 && di < 0))) return 0;   // unexpected order
        early[o] = make_int4(di, dj, dp, di + B * dj + B2 * dp);
    }
    // (dj,dplane) rows of the earlier half with their contiguous di range (k_outlet_sweep_mod)
    std::vector<int4> rows;
    for (int o = 0; o < n_early; ++o) {
        const int4& e = early[o];
        if (!rows.empty() && rows.back().x == e.y && rows.back().y == e.z && rows.back().w + 1 == e.x) rows.back().w = e.x;
        else rows.push_back(make_int4(e.y, e.z, e.x, e.x));
    }
    // lattice-addressed rings: the level front keeps (maxd + Nx - 1)/B + 1 rows of a plane alive
    int RJ = 1;
    if (Ny > 1) {
        const int win = (maxd + Nx - 1) / B + 1;
        RJ = 64;
        while (RJ < win) RJ <<= 1;
    }
    size_t smem_mod = sizeof(double) * (size_t)KP * RJ * ring + sizeof(int4) * rows.size() + sizeof(unsigned) * mask_words;
    c->out_mod = smem_mod <= 220 * 1024 && KP * Wj <= 2 * (1024 / 4);
    c->out_RJ = RJ; c->out_n_rows = (int)rows.size(); c->out_smem_mod = smem_mod;
    c->out_KP = KP; c->out_Wj = Wj; c->out_ring = ring; c->out_smem = smem; c->out_mask_words = mask_words;
    c->out_l0 = al_min * c->P;
    c->out_tau_max = (Nx - 1) + B * (Ny - 1) + B2 * (KP - 1);
    CUDA_OK(cudaMalloc(&c->out_base_v, sizeof(double) * nslab));
    CUDA_OK(cudaMalloc(&c->out_base_c, sizeof(double) * nslab));
    CUDA_OK(cudaMalloc(&c->out_cnt, sizeof(int) * nslab));
    CUDA_OK(cudaMalloc(&c->out_mask, sizeof(unsigned) * mask_words));
    CUDA_OK(cudaMalloc(&c->out_early, sizeof(int4) * n_early));
    CUDA_OK(cudaMemset(c->out_base_v, 0, sizeof(double) * nslab));
    CUDA_OK(cudaMemset(c->out_base_c, 0, sizeof(double) * nslab));
    CUDA_OK(cudaMemset(c->out_cnt, 0xFF, sizeof(int) * nslab));   // -1 = not an OUTLET node (pre-pass writes OUTLET nodes)
    CUDA_OK(cudaMemcpy(c->out_early, early.data(), sizeof(int4) * n_early, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&c->out_rows, sizeof(int4) * rows.size()));
    CUDA_OK(cudaMemcpy(c->out_rows, rows.data(), sizeof(int4) * rows.size(), cudaMemcpyHostToDevice));
    if (c->out_mod)
        CUDA_OK(cudaFuncSetAttribute(k_outlet_sweep_mod, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mod));
    // row-walking sweep: G lanes per lattice row slot, M row slots per outlet plane
    c->out_rows_G = 0;
    if (KP <= 8) {
        for (int doubled = 1; doubled >= 0 && !c->out_rows_G; --doubled) {
            size_t smem_rows = sizeof(double) * ((size_t)KP * RJ * (doubled ? 2 : 1) * ring + c->n_off + 1) +
                               sizeof(int4) * rows.size();
            for (int G : {4, 8, 2}) {   // 4 lanes per row measured fastest (the sweep is issue bound: fewer warps)
                if (c->opt_outlet_rows_g && G != c->opt_outlet_rows_g) continue;
                int M = (Ny == 1) ? 1 : (Nx + G + c->R + B - 1) / B;
                if (KP * M * G <= 1024 && smem_rows <= 220 * 1024) {
                    c->out_rows_G = G; c->out_rows_M = M; c->out_smem_rows = smem_rows; c->out_rows_doubled = doubled;
                    break;
                }
            }
        }
        const size_t smem_rows = c->out_smem_rows;
        for (int kp = 0; kp < 8; ++kp) {
            int st = 0;
            while (st < (int)rows.size() && kp + rows[st].y < 0) ++st;
            c->out_row_start[kp] = st;
        }
        if (c->out_rows_G == 8)
            CUDA_OK(cudaFuncSetAttribute(k_outlet_sweep_rows<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
        if (c->out_rows_G == 4)
            CUDA_OK(cudaFuncSetAttribute(k_outlet_sweep_rows<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
        if (c->out_rows_G == 2)
            CUDA_OK(cudaFuncSetAttribute(k_outlet_sweep_rows<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    }
    k_outlet_mask<<<nblocks(mask_words, 256), 256, 0, c->stream>>>(c->type, c->out_l0, nslab, c->out_mask);
    CUDA_OK(cudaFuncSetAttribute(k_outlet_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->out_fast = true;
    return 0;
}

static OutletGeom outlet_geom(const pdgpu_ctx* c) {
    OutletGeom g;
    g.Nx = c->Nx; g.Ny = (c->dim == 3) ? c->Ny : 1; g.KP = c->out_KP; g.Wj = c->out_Wj;
    g.B = c->R + 1; g.B2 = g.B * g.B; g.ring = c->out_ring; g.n_early = c->n_off / 2;
    g.tau_max = c->out_tau_max; g.P = c->P; g.l0 = c->out_l0;
    return g;
}

// The two halves of the fast outlet BC. Loop bodies that overlap the sweep with the bulk bond kernel
// run the pre-pass BEFORE they fork, so that the sweep kernel is the first thing the side stream has
// ready: once the bulk kernel owns every SM the two sweep CTAs (197 KB of shared memory each) only
// get in when it drains (measured: the whole sweep then runs after the bulk kernel).
int pd_enqueue_bc_outlet_prepass(pdgpu_ctx* c, int buf, int bufC) {
    if (!c->out_fast || c->opt_outlet_kernel == 0) return -1;
    Lat L = make_lat(c);
    OutletGeom g = outlet_geom(c);
    if (c->dim == 2)
        LAUNCH(c, k_outlet_prepass<2>, nblocks(c->n_outlet, 128), 128, 0, L, g, c->l_outlet, c->n_outlet, c->type,
               c->d_off, c->n_off, c->rho[buf], c->p[buf], VXYZ(c, buf), c->C[bufC], c->cfg.rho_f, c->out_base_v,
               c->out_base_c, c->out_cnt);
    else
        LAUNCH(c, k_outlet_prepass<3>, nblocks(c->n_outlet, 128), 128, 0, L, g, c->l_outlet, c->n_outlet, c->type,
               c->d_off, c->n_off, c->rho[buf], c->p[buf], VXYZ(c, buf), c->C[bufC], c->cfg.rho_f, c->out_base_v,
               c->out_base_c, c->out_cnt);
    return 0;
}

int pd_enqueue_bc_outlet_sweep(pdgpu_ctx* c, int buf, int bufC) {
    if (!c->out_fast || c->opt_outlet_kernel == 0) return -1;
    OutletGeom g = outlet_geom(c);
    double* vax = c->v[buf][c->dim - 1];
    if (c->out_rows_G && c->opt_outlet_kernel >= 3) {
        RowSweepParams q;
        q.g = g; q.RJ = c->out_RJ; q.n_rows = c->out_n_rows; q.M = c->out_rows_M; q.R = c->R; q.n_rcp = c->n_off + 1; q.doubled = (c->out_rows_doubled && !c->opt_outlet_single_rows) ? 1 : 0;
        for (int kp = 0; kp < 8; ++kp) q.row_start[kp] = c->out_row_start[kp];
        const int threads = (g.KP * q.M * c->out_rows_G + 31) / 32 * 32;
        if (c->out_rows_G == 8)
            LAUNCH(c, k_outlet_sweep_rows<8>, 2, threads, c->out_smem_rows, q, (const int4*)c->out_rows, c->out_base_v,
                   c->out_base_c, c->out_cnt, vax, c->C[bufC], c->cfg.U_in);
        else if (c->out_rows_G == 4)
            LAUNCH(c, k_outlet_sweep_rows<4>, 2, threads, c->out_smem_rows, q, (const int4*)c->out_rows, c->out_base_v,
                   c->out_base_c, c->out_cnt, vax, c->C[bufC], c->cfg.U_in);
        else
            LAUNCH(c, k_outlet_sweep_rows<2>, 2, threads, c->out_smem_rows, q, (const int4*)c->out_rows, c->out_base_v,
                   c->out_base_c, c->out_cnt, vax, c->C[bufC], c->cfg.U_in);
    } else if (c->out_mod && c->opt_outlet_kernel >= 2)
        LAUNCH(c, k_outlet_sweep_mod, 2, 1024, c->out_smem_mod, g, c->out_RJ, (const int4*)c->out_rows, c->out_n_rows,
               c->out_mask, c->out_mask_words, c->out_base_v, c->out_base_c, c->out_cnt, vax, c->C[bufC], c->cfg.U_in);
    else
        LAUNCH(c, k_outlet_sweep, 2, 1024, c->out_smem, g, (const int4*)c->out_early, c->out_mask, c->out_mask_words,
               c->out_base_v, c->out_base_c, c->out_cnt, vax, c->C[bufC], c->cfg.U_in);
    return 0;
}

// returns -1 when the fast sweep is not applicable
int pd_enqueue_bc_outlet_fast(pdgpu_ctx* c, int buf, int bufC) {
    if (pd_enqueue_bc_outlet_prepass(c, buf, bufC) < 0) return -1;
    return pd_enqueue_bc_outlet_sweep(c, buf, bufC);
}
