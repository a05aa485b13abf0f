// ns_march.cu -- z-marching variant of the tiled PD-NS bond kernel (3D, m_ratio = 3, full rows).
//
// Same arithmetic as ns_tile.cu (see there for the bond algebra and numerics); different
// schedule.  ns_tile.cu stages a fresh haloed block per CTA, so every tile pays the staging
// latency, 5.2x read amplification from L2 and a CTA launch gap (ncu: 25 % of the warp samples
// outside the bond loop).  Here a CTA owns a 16 x 12 column of nodes and MARCHES along z in
// steps of 4 planes through a chunk of 32 planes: a ring of 14 planes per field lives in shared
// memory (222 KB); while the bonds of the current step are summed, the 4 planes the next step
// needs arrive by cp.async (LDGSTS, zero fill outside the box).  Each staged plane is loaded
// once per chunk (2.1x read amplification) and the loads overlap the FP64 work.
//
// Thread (tx, ty, tz) of the 16 x 12 x 2 CTA owns nodes (x0+tx, y0+ty, z0+2tz..z0+2tz+1) of the
// current step; a half-warp covers one x-row: shared-memory rows are contiguous 128 B segments.
#include <algorithm>

#include "tile.cuh"

namespace {
using tile::ColTable;
using tile::TileGeom;
using tile::cp_async8;

constexpr int TR = 3;
constexpr int TX = 16, TY = 12;            // node column of a CTA
constexpr int RZ = 2, NZT = 2, TZ = RZ * NZT;
constexpr int SX = TX + 2 * TR, SY = TY + 2 * TR, SPL = SX * SY;   // staged plane
constexpr int NR = 2 * TZ + 2 * TR;        // ring depth: current step (TZ+6 planes) + next step's TZ planes
constexpr int NTH = TX * TY * NZT;
constexpr int CHUNK = 32;                  // planes per CTA (multiple of TZ)
constexpr int NCOL = tile::NCOL;
constexpr int NF = 5;                      // rho, p, vx, vy, vz

struct NsMarchParams {
    TileGeom g;                            // z_lo/z_hi = plane range of this launch
    double rho_f, gamma, B;
    double c_div, dens_diff, visc, rho_lo, rho_hi, W2, inv_dx;
    int gamma_is_7;
};

__device__ __forceinline__ double eos_march(double rho, const NsMarchParams& q) {
    double ratio = rho / q.rho_f;
    ratio = fmin(fmax(ratio, 0.5), 2.0);
    if (q.gamma_is_7) {   // (1+e)^7 - 1 by Horner, see ns_tile.cu
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

// po[q] = shared-memory offset of ring plane (zt - 3 + q), q = 0 .. RZ+5
template <int H>
__device__ __forceinline__ void ns_column(const double* __restrict__ s_rho, const double* __restrict__ s_p,
                                          const double* __restrict__ s_vx, const double* __restrict__ s_vy,
                                          const double* __restrict__ s_vz, const int (&po)[RZ + 2 * TR], int cb,
                                          double dI, double dJ, const double (&kap)[4], const double (&kz)[4],
                                          const double (&nk)[4], double c_div, NsAcc& a) {
    double colp[RZ];
#pragma unroll
    for (int t = 0; t < RZ; ++t) colp[t] = 0.0;
#pragma unroll
    for (int zz = -H; zz < RZ + H; ++zz) {
        const int si = po[zz + TR] + cb;
        const double rj = s_rho[si], pj = s_p[si], ux = s_vx[si], uy = s_vy[si], uz = s_vz[si];
        const double mz = rj * uz;
        const double axy = rj * fma(dI, ux, dJ * uy);
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

__global__ void __launch_bounds__(NTH, 1)
k_ns_march(const __grid_constant__ NsMarchParams q, const __grid_constant__ ColTable T,
           const double* __restrict__ d_dt, const uint8_t* __restrict__ type, const double* __restrict__ rho,
           const double* __restrict__ pr, const double* __restrict__ vx, const double* __restrict__ vy,
           const double* __restrict__ vz, double* __restrict__ rho_n, double* __restrict__ pr_n,
           double* __restrict__ vx_n, double* __restrict__ vy_n, double* __restrict__ vz_n) {
    extern __shared__ double sm[];
    double* s_rho = sm;
    double* s_p = sm + 1 * NR * SPL;
    double* s_vx = sm + 2 * NR * SPL;
    double* s_vy = sm + 3 * NR * SPL;
    double* s_vz = sm + 4 * NR * SPL;

    const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
    const int tid = (tz * TY + ty) * TX + tx;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int zb = q.g.z_lo + blockIdx.z * CHUNK;
    const int ze = min(zb + CHUNK, q.g.z_hi);
    const int gx = x0 + tx, gy = y0 + ty;
    const bool in_xy = gx < q.g.Nx && gy < q.g.Ny;
    const long long lxy = (long long)gy * q.g.Nx + gx;

    // ---- does this column hold any FLUID node in this chunk? (else: copy-through only) ----
    bool col_any = false;
    if (in_xy)
        for (int z = zb + tz; z < ze; z += NZT) col_any = col_any || (type[(long long)z * q.g.P + lxy] == PDGPU_FLUID);
    if (!__syncthreads_or(col_any)) {
        if (in_xy)
            for (int z = zb + tz; z < ze; z += NZT) {   // src/pd_ns.cpp:93-97
                const long long l = (long long)z * q.g.P + lxy;
                rho_n[l] = rho[l]; pr_n[l] = pr[l]; vx_n[l] = vx[l]; vy_n[l] = vy[l]; vz_n[l] = vz[l];
            }
        return;
    }

    // stage `count` planes starting at local plane zfirst into their ring slots (zero fill
    // outside the box / beyond the local array)
    auto stage = [&](int zfirst, int count) {
        for (int idx = tid; idx < count * SPL; idx += NTH) {
            const int pl = idx / SPL;
            const int rem = idx - pl * SPL;
            const int sy = rem / SX, sx = rem - sy * SX;
            const int az = zfirst + pl, ax = x0 - TR + sx, ay = y0 - TR + sy;
            const bool ok = ax >= 0 && ax < q.g.Nx && ay >= 0 && ay < q.g.Ny && az < q.g.nlp;
            const long long l = ok ? (long long)az * q.g.P + (long long)ay * q.g.Nx + ax : 0;
            const int so = (az % NR) * SPL + rem;
            cp_async8(s_rho + so, rho + l, ok);
            cp_async8(s_p + so, pr + l, ok);
            cp_async8(s_vx + so, vx + l, ok);
            cp_async8(s_vy + so, vy + l, ok);
            cp_async8(s_vz + so, vz + l, ok);
        }
    };
    auto load_types = [&](int z0, uint8_t (&nty)[RZ]) {
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int lz = z0 + tz * RZ + t;
            nty[t] = 255;
            if (in_xy && lz < ze) nty[t] = type[(long long)lz * q.g.P + lxy];
        }
    };

    stage(zb - TR, TZ + 2 * TR);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    uint8_t nty[RZ], nty_next[RZ];
    load_types(zb, nty);

    const int cb0 = (ty + TR) * SX + (tx + TR);
    const double vW = q.visc * q.W2;
    const double dt = *d_dt;

    for (int z0 = zb; z0 < ze; z0 += TZ) {
        const bool more = z0 + TZ < ze;
        if (more) stage(z0 + TZ + TR, TZ);           // planes only the next step needs
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        if (more) load_types(z0 + TZ, nty_next);
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");   // everything but the newest group
        __syncthreads();

        const int zt = z0 + tz * RZ;
        bool fl[RZ];
        bool any = false;
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            fl[t] = (nty[t] == PDGPU_FLUID);
            any = any || fl[t];
            if (nty[t] != 255 && !fl[t]) {   // copy-through
                const long long l = (long long)(zt + t) * q.g.P + lxy;
                rho_n[l] = rho[l]; pr_n[l] = pr[l]; vx_n[l] = vx[l]; vy_n[l] = vy[l]; vz_n[l] = vz[l];
            }
        }
        if (__any_sync(0xffffffffu, any)) {
            int po[RZ + 2 * TR];
#pragma unroll
            for (int u = 0; u < RZ + 2 * TR; ++u) po[u] = ((zt - TR + u) % NR) * SPL;
            NsAcc a;
#pragma unroll
            for (int t = 0; t < RZ; ++t)
                a.mc[t] = a.md[t] = a.ax[t] = a.ay[t] = a.az[t] = a.px[t] = a.py[t] = a.pz[t] = 0.0;
#pragma unroll 1
            for (int c = 0; c < NCOL; ++c) {
                const int cb = cb0 + T.dj_i[c] * SX + T.di_i[c];
                const double dI = T.di[c], dJ = T.dj[c];
                const double kap[4] = {T.kap[c][0], T.kap[c][1], T.kap[c][2], T.kap[c][3]};
                const double kz[4] = {T.kz[c][0], T.kz[c][1], T.kz[c][2], T.kz[c][3]};
                const double nk[4] = {T.aux[c][0], T.aux[c][1], T.aux[c][2], T.aux[c][3]};
                const int H = T.h[c];
                if (H == 3) ns_column<3>(s_rho, s_p, s_vx, s_vy, s_vz, po, cb, dI, dJ, kap, kz, nk, q.c_div, a);
                else if (H == 2) ns_column<2>(s_rho, s_p, s_vx, s_vy, s_vz, po, cb, dI, dJ, kap, kz, nk, q.c_div, a);
                else ns_column<1>(s_rho, s_p, s_vx, s_vy, s_vz, po, cb, dI, dJ, kap, kz, nk, q.c_div, a);
            }
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                if (!fl[t]) continue;
                const int si = po[t + TR] + cb0;
                const long long l = (long long)(zt + t) * q.g.P + lxy;
                const double rho_i = s_rho[si], vi0 = s_vx[si], vi1 = s_vy[si], vi2 = s_vz[si];
                const double mass_diff = a.md[t] * q.inv_dx - rho_i * q.W2;
                double rn = rho_i + dt * (-q.c_div * a.mc[t] + q.dens_diff * mass_diff);   // src/pd_ns.cpp:160-168
                rn = fmin(fmax(rn, q.rho_lo), q.rho_hi);
                rho_n[l] = rn;
                pr_n[l] = eos_march(rn, q);
                const double s = dt / rho_i;                                                // :171-178
                vx_n[l] = vi0 + s * (a.ax[t] - q.c_div * a.px[t] - vW * vi0);
                vy_n[l] = vi1 + s * (a.ay[t] - q.c_div * a.py[t] - vW * vi1);
                vz_n[l] = vi2 + s * (a.az[t] - q.c_div * a.pz[t] - vW * vi2);
            }
        }
        __syncthreads();   // ring slots of this step's oldest planes may be overwritten now
#pragma unroll
        for (int t = 0; t < RZ; ++t) nty[t] = nty_next[t];
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

}  // namespace

// returns -1 when the kernel does not apply
int pd_enqueue_ns_march(pdgpu_ctx* c, int src, const double* d_dt, int zb, int ze) {
    if (!c->full_rows) return -1;
    static ColTable T;
    double sum_kappa = 0.0;
    if (!tile::build_columns(c, &T, &sum_kappa)) return -1;
    PdConsts k = pd_consts(c->cfg, c->dim);
    NsMarchParams q;
    q.g = tile::make_geom(c);
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
    const size_t smem = sizeof(double) * NF * NR * SPL;
    static bool attr_done = false;
    if (!attr_done) {
        CUDA_OK(cudaFuncSetAttribute(k_ns_march, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    int dst = 1 - src;
    dim3 grid((c->Nx + TX - 1) / TX, (c->Ny + TY - 1) / TY, ((q.g.z_hi - q.g.z_lo) + CHUNK - 1) / CHUNK);
    dim3 block(TX, TY, NZT);
    k_ns_march<<<grid, block, smem, c->stream>>>(q, T, d_dt, c->type, c->rho[src], c->p[src], c->v[src][0],
                                                  c->v[src][1], c->v[src][2], c->rho[dst], c->p[dst], c->v[dst][0],
                                                  c->v[dst][1], c->v[dst][2]);
    c->launches++;
    return 0;
}
