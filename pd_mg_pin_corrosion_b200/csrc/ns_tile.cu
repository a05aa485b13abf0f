// ns_tile.cu -- tiled fast path of the PD-NS bond kernel (3D, m_ratio = 3, full FLUID rows).
// Tile geometry, column walk and weight algebra: tile.cuh.
//
// Arithmetic per bond (10 FP64 ops + ~2.5 amortised): the reference's difference form
//     sum_j (f_j - f_i) e w        (src/pd_ns.cpp:115-157)
// is evaluated as  sum_j f_j e w  for the odd (gradient/divergence) sums -- the f_i part
// multiplies sum_j e_j w_j, which is exactly zero for the full symmetric stencil -- and as
// sum_j f_j w2 - f_i * sum_j w2 for the two Laplacians.  One explicit step changes a state
// value by <= ~1e-3 relative, so this reassociation moves results by O(1e-16) relative
// (DESIGN.md "numerics"); parity against the reference is asserted at 1e-12.
//   g      = rho_j (v_j . d) kappa             mass flux through the bond
//   mc    += g                                  mass convection
//   md    += rho_j kappa                        density Laplacian (x 1/dx at the end)
//   h      = (mu beta/dx) kappa - (alpha/V_H) g momentum convection and viscous Laplacian share
//   a_d   += v_jd h                             one weight per bond
//   P_d   += p_j d_d kappa                      pressure gradient (di/dj parts factored out per column)
// with d = (di,dj,dk).
#include <algorithm>

#include "stream.cuh"

namespace {
using namespace tile;

struct NsTileParams {
    TileGeom g;
    double rho_f, gamma, B;        // EOS
    double c_div, dens_diff, visc, rho_lo, rho_hi;
    double W2;                     // sum of w2 over the stencil (= sum kappa / dx)
    double inv_dx;
    int gamma_is_7;
};

__device__ __forceinline__ double eos_tile(double rho, const NsTileParams& q) {
    double ratio = rho / q.rho_f;
    ratio = fmin(fmax(ratio, 0.5), 2.0);
    if (q.gamma_is_7) {
        // (1+e)^7 - 1 by Horner: no cancellation (the reference's pow(ratio,7)-1 rounds at
        // ulp(1); this agrees with it to ~1 ulp(1) * B)
        double e = ratio - 1.0;
        double s = e + 7.0;
        s = fma(s, e, 21.0);
        s = fma(s, e, 35.0);
        s = fma(s, e, 35.0);
        s = fma(s, e, 21.0);
        s = fma(s, e, 7.0);
        return q.B * (s * e);
    }
    return q.B * (pow(ratio, q.gamma) - 1.0);
}

struct NsAcc {
    double mc[RZ], md[RZ], ax[RZ], ay[RZ], az[RZ], px[RZ], py[RZ], pz[RZ];
};

template <int H>
__device__ __forceinline__ void ns_column(const double* __restrict__ s_rho, const double* __restrict__ s_vx,
                                          const double* __restrict__ s_vy, const double* __restrict__ s_vz,
                                          const double* __restrict__ s_p, int cb, double dI, double dJ,
                                          const double (&kap)[4], const double (&kz)[4], const double (&nk)[4],
                                          double c_div, NsAcc& a) {
    double colp[RZ];
#pragma unroll
    for (int t = 0; t < RZ; ++t) colp[t] = 0.0;
#pragma unroll
    for (int zz = -H; zz < RZ + H; ++zz) {
        const int si = cb + (zz + TR) * SPLANE;
        const double rj = s_rho[si], pj = s_p[si], ux = s_vx[si], uy = s_vy[si], uz = s_vz[si];
        const double mz = rj * uz;
        const double axy = dI * (rj * ux) + dJ * (rj * uy);
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int dk = zz - t;
            if (dk >= -H && dk <= H) {
                const int ak = dk < 0 ? -dk : dk;
                const double k = kap[ak];
                double g = axy * k;
                if (dk > 0) g = fma(mz, kz[ak], g);
                if (dk < 0) g = fma(-mz, kz[ak], g);
                a.mc[t] += g;
                a.md[t] = fma(rj, k, a.md[t]);
                const double h = fma(-c_div, g, nk[ak]);
                a.ax[t] = fma(ux, h, a.ax[t]);
                a.ay[t] = fma(uy, h, a.ay[t]);
                a.az[t] = fma(uz, h, a.az[t]);
                colp[t] = fma(pj, k, colp[t]);
                if (dk > 0) a.pz[t] = fma(pj, kz[ak], a.pz[t]);
                if (dk < 0) a.pz[t] = fma(-pj, kz[ak], a.pz[t]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        a.px[t] = fma(dI, colp[t], a.px[t]);
        a.py[t] = fma(dJ, colp[t], a.py[t]);
    }
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_ns_tile(const __grid_constant__ NsTileParams q, const __grid_constant__ ColTable T,
          const double* __restrict__ d_dt, const uint8_t* __restrict__ type, const double* __restrict__ rho,
          const double* __restrict__ pr, const double* __restrict__ vx, const double* __restrict__ vy,
          const double* __restrict__ vz, double* __restrict__ rho_n, double* __restrict__ pr_n,
          double* __restrict__ vx_n, double* __restrict__ vy_n, double* __restrict__ vz_n) {
    extern __shared__ double sm[];
    double* s_rho = sm;
    double* s_vx = sm + SN;
    double* s_vy = sm + 2 * SN;
    double* s_vz = sm + 3 * SN;
    double* s_p = sm + 4 * SN;

    const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
    const int tid = (tz * TY + ty) * TX + tx;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, z0 = q.g.z_lo + blockIdx.z * TZ;
    const int gx = x0 + tx, gy = y0 + ty, zt = z0 + tz * RZ;   // first z-node of this thread
    const bool in_xy = gx < q.g.Nx && gy < q.g.Ny;

    // node types of this thread's nodes (loads issued first, consumed after the staging copies)
    uint8_t nty[RZ];
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        const int lz = zt + t;
        nty[t] = 255;
        if (in_xy && lz < q.g.z_hi) nty[t] = type[(long long)lz * q.g.P + (long long)gy * q.g.Nx + gx];
    }
    // stage the haloed block with asynchronous copies; outside the box the values are never
    // used by a FLUID row (full rows) -> zero fill
    for (int idx = tid; idx < SN; idx += NTHREADS) {
        const long long l = staged_index(q.g, idx, x0, y0, z0);
        const bool ok = l >= 0;
        const long long ls = ok ? l : 0;
        cp_async8(s_rho + idx, rho + ls, ok);
        cp_async8(s_p + idx, pr + ls, ok);
        cp_async8(s_vx + idx, vx + ls, ok);
        cp_async8(s_vy + idx, vy + ls, ok);
        cp_async8(s_vz + idx, vz + ls, ok);
    }
    bool fl[RZ];
    bool any = false;
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        fl[t] = (nty[t] == PDGPU_FLUID);
        any = any || fl[t];
        if (nty[t] != 255 && !fl[t]) {   // copy-through (src/pd_ns.cpp:93-97)
            const long long l = (long long)(zt + t) * q.g.P + (long long)gy * q.g.Nx + gx;
            rho_n[l] = rho[l]; pr_n[l] = pr[l]; vx_n[l] = vx[l]; vy_n[l] = vy[l]; vz_n[l] = vz[l];
        }
    }
    cp_async_wait_all();
    if (!__syncthreads_or(any)) return;          // tile without FLUID nodes
    if (!__any_sync(0xffffffffu, any)) return;   // warp without FLUID nodes

    NsAcc a;
#pragma unroll
    for (int t = 0; t < RZ; ++t)
        a.mc[t] = a.md[t] = a.ax[t] = a.ay[t] = a.az[t] = a.px[t] = a.py[t] = a.pz[t] = 0.0;
    const int base = (tz * RZ * SY + ty + TR) * SX + (tx + TR);   // node t at base + (t+TR)*SPLANE
#pragma unroll 1
    for (int c = 0; c < NCOL; ++c) {
        const int cb = base + T.off[c];
        const double dI = T.di[c], dJ = T.dj[c];
        const double kap[4] = {T.kap[c][0], T.kap[c][1], T.kap[c][2], T.kap[c][3]};
        const double kz[4] = {T.kz[c][0], T.kz[c][1], T.kz[c][2], T.kz[c][3]};
        const double nk[4] = {T.aux[c][0], T.aux[c][1], T.aux[c][2], T.aux[c][3]};
        const int H = T.h[c];
        if (H == 3) ns_column<3>(s_rho, s_vx, s_vy, s_vz, s_p, cb, dI, dJ, kap, kz, nk, q.c_div, a);
        else if (H == 2) ns_column<2>(s_rho, s_vx, s_vy, s_vz, s_p, cb, dI, dJ, kap, kz, nk, q.c_div, a);
        else ns_column<1>(s_rho, s_vx, s_vy, s_vz, s_p, cb, dI, dJ, kap, kz, nk, q.c_div, a);
    }

    const double dt = *d_dt;
    const double vW = q.visc * q.W2;
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        if (!fl[t]) continue;
        const int si = base + (t + TR) * SPLANE;
        const long long l = (long long)(zt + t) * q.g.P + (long long)gy * q.g.Nx + gx;
        const double rho_i = s_rho[si], vi0 = s_vx[si], vi1 = s_vy[si], vi2 = s_vz[si];
        const double mass_diff = a.md[t] * q.inv_dx - rho_i * q.W2;
        double rn = rho_i + dt * (-q.c_div * a.mc[t] + q.dens_diff * mass_diff);   // src/pd_ns.cpp:160-168
        rn = fmin(fmax(rn, q.rho_lo), q.rho_hi);
        rho_n[l] = rn;
        pr_n[l] = eos_tile(rn, q);
        const double s = dt / rho_i;                                                // :171-178
        vx_n[l] = vi0 + s * (a.ax[t] - q.c_div * a.px[t] - vW * vi0);
        vy_n[l] = vi1 + s * (a.ay[t] - q.c_div * a.py[t] - vW * vi1);
        vz_n[l] = vi2 + s * (a.az[t] - q.c_div * a.pz[t] - vW * vi2);
    }
}

}  // namespace

// returns -1 when the tiled kernel does not apply (caller uses the generic kernel)
int pd_enqueue_ns_step_fast(pdgpu_ctx* c, int src, const double* d_dt, int zb, int ze) {
    if (!c->full_rows) return -1;
    if (pd_stream_prepare(c)) return -1;   // column table per context (stream.cuh)
    stream::TileState* ts = pd_tile_state(c);
    ColTable& T = ts->tcols;
    const double sum_kappa = ts->sum_kappa;
    PdConsts k = pd_consts(c->cfg, c->dim);
    NsTileParams q;
    q.g = make_geom(c);
    if (zb >= 0) { q.g.z_lo = zb; q.g.z_hi = ze; }
    if (q.g.z_hi <= q.g.z_lo) return 0;
    q.rho_f = c->cfg.rho_f; q.gamma = c->cfg.gamma_eos; q.B = k.B_eos;
    q.c_div = k.alpha * k.inv_VH; q.dens_diff = k.dens_diff_coeff; q.visc = c->cfg.mu_f * k.beta_lap;
    q.rho_lo = 0.5 * c->cfg.rho_f; q.rho_hi = 2.0 * c->cfg.rho_f;
    q.inv_dx = 1.0 / c->cfg.dx;
    q.W2 = sum_kappa * q.inv_dx;
    q.gamma_is_7 = (c->cfg.gamma_eos == 7.0);
    for (int col = 0; col < tile::NCOL; ++col)
        for (int kk = 0; kk < 4; ++kk) T.aux[col][kk] = q.visc * q.inv_dx * T.kap[col][kk];
    const size_t smem = sizeof(double) * 5 * SN;
    if (!ts->attr_tile_ns) {   // per context: the attribute is per device
        CUDA_OK(cudaFuncSetAttribute(k_ns_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ts->attr_tile_ns = true;
    }
    int dst = 1 - src;
    dim3 grid((c->Nx + TX - 1) / TX, (c->Ny + TY - 1) / TY, ((q.g.z_hi - q.g.z_lo) + TZ - 1) / TZ);
    dim3 block(TX, TY, NZT);
    k_ns_tile<<<grid, block, smem, c->stream>>>(q, T, d_dt, c->type, c->rho[src], c->p[src], c->v[src][0],
                                                 c->v[src][1], c->v[src][2], c->rho[dst], c->p[dst], c->v[dst][0],
                                                 c->v[dst][1], c->v[dst][2]);
    c->launches++;
    return 0;
}
