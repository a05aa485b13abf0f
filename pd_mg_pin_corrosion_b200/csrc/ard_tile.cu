// ard_tile.cu -- tiled fast path of the explicit PD-ARD bond kernel (3D, m_ratio = 3, full rows).
//
// Same staging scheme as ns_tile.cu (22 x 14 x 14 haloed block, 2 z-nodes per thread, sliding
// z-window, runtime column loop: tile.cuh).  Staged per node:
//     C      concentration
//     w      packed weight (ard.cu, k_ard_vmag): own fluid-fluid diffusivity D_l + alpha dx |v| >= +0
//            for fluid-like nodes (FLUID / INLET / OUTLET); -dsol <= -0 for a SOLID_MG node, dsol =
//            interface diffusivity 2 D_l D_s / (D_l + D_s) (0 when salt blocked; grain-boundary /
//            precipitate / grain class and the volume-loss decay folded in); -0 for WALL/OUTSIDE
//                                                                    (src/pd_ard.cpp:136-170)
// Two staged fields = 77 KB per CTA: two CTAs per SM, so one CTA's staging overlaps the other's
// bond loop.  For a FLUID row the bond classification of src/pd_ard.cpp:120-181 becomes
//     D_ij = | umax(w_i, w_j) |  ( = f_j ? max(w_i, w_j) : dsol_j ),   f_j = [sign bit of w_j clear]
//     diff += D_ij (C_j - C_i) w2 ;   G += f_j (C_j - C_i) e w1 ;   adv = (alpha/V_H) v_i . G
// (WALL neighbours have f = 0, dsol = 0 and drop out).  max and the select are done on the
// integer pipe (non-negative doubles order like their bit patterns): an FP64 max costs a DSETP
// plus six moves/selects on sm_100a.  SOLID_MG rows (2 % of the nodes) are done by
// k_ard_solid_rows from the solid node list.
#include <algorithm>

#include "stream.cuh"

namespace {
using namespace tile;

struct ArdTileParams {
    TileGeom g;
    int skip_wall_copy;   // WALL values of the new buffer are written by the wall-concentration BC kernel
    double beta, div_coeff, inv_dx;
};

struct ArdAcc {
    double diff[RZ], gx[RZ], gy[RZ], gz[RZ];
};

// per bond: 7 FP64 ops (dC, D*dC, diff, f*dC, colg, gz + column tail) and 4 integer ops (64-bit unsigned max)
template <int H>
__device__ __forceinline__ void ard_column(const double* __restrict__ s_C, const double* __restrict__ s_w, int cb,
                                           double dI, double dJ, const double (&kap)[4], const double (&kz)[4],
                                           const double (&Ci)[RZ], const unsigned long long (&wi)[RZ], ArdAcc& a) {
    double colg[RZ];
#pragma unroll
    for (int t = 0; t < RZ; ++t) colg[t] = 0.0;
#pragma unroll
    for (int zz = -H; zz < RZ + H; ++zz) {
        const int si = cb + (zz + TR) * SPLANE;
        const double Cj = s_C[si];
        const unsigned long long wj = (unsigned long long)__double_as_longlong(s_w[si]);
        const double f = (int)(wj >> 32) >= 0 ? 1.0 : 0.0;        // sign bit clear: fluid-like
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int dk = zz - t;
            if (dk >= -H && dk <= H) {
                const int ak = dk < 0 ? -dk : dk;
                const double k = kap[ak];
                const double dC = Cj - Ci[t];
                // unsigned max: a set sign bit (solid / wall neighbour) beats every w_i >= +0, and
                // |.| then yields dsol_j; for a fluid-like neighbour it is max(w_i, w_j)
                const unsigned long long wm = wi[t] > wj ? wi[t] : wj;
                const double D = fabs(__longlong_as_double((long long)wm));
                a.diff[t] = fma(D * dC, k, a.diff[t]);
                const double fd = f * dC;
                colg[t] = fma(fd, k, colg[t]);
                if (dk > 0) a.gz[t] = fma(fd, kz[ak], a.gz[t]);
                if (dk < 0) a.gz[t] = fma(-fd, kz[ak], a.gz[t]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        a.gx[t] = fma(dI, colg[t], a.gx[t]);
        a.gy[t] = fma(dJ, colg[t], a.gy[t]);
    }
}

// All-fluid fast path: every neighbour of every FLUID node of the warp is fluid-like (FLUID / INLET /
// OUTLET; flag built by k_nbr_fluid_only), so f = 1, D_ij = max(w_i, w_j) with both >= +0, and the
// bond classification, the sign handling and the f multiply of ard_column drop out:
// 5 FP64 + 4 integer operations per bond.
template <int H>
__device__ __forceinline__ void ard_column_fluid(const double* __restrict__ s_C, const double* __restrict__ s_w, int cb,
                                                 double dI, double dJ, const double (&kap)[4], const double (&kz)[4],
                                                 const double (&Ci)[RZ], const unsigned long long (&wi)[RZ],
                                                 ArdAcc& a) {
    double colg[RZ];
#pragma unroll
    for (int t = 0; t < RZ; ++t) colg[t] = 0.0;
#pragma unroll
    for (int zz = -H; zz < RZ + H; ++zz) {
        const int si = cb + (zz + TR) * SPLANE;
        const double Cj = s_C[si];
        const unsigned long long wj = (unsigned long long)__double_as_longlong(s_w[si]);
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int dk = zz - t;
            if (dk >= -H && dk <= H) {
                const int ak = dk < 0 ? -dk : dk;
                const double k = kap[ak];
                const double dC = Cj - Ci[t];
                // FP64 compare + two selects: one instruction less than the 64-bit integer max, and this
                // kernel is issue bound with the FP64 pipe half idle
                const double wid = __longlong_as_double((long long)wi[t]), wjd = __longlong_as_double((long long)wj);
                const double D = wid > wjd ? wid : wjd;
                a.diff[t] = fma(D * dC, k, a.diff[t]);
                colg[t] = fma(dC, k, colg[t]);
                if (dk > 0) a.gz[t] = fma(dC, kz[ak], a.gz[t]);
                if (dk < 0) a.gz[t] = fma(-dC, kz[ak], a.gz[t]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        a.gx[t] = fma(dI, colg[t], a.gx[t]);
        a.gy[t] = fma(dJ, colg[t], a.gy[t]);
    }
}

// nbf[l] = 1 for a FLUID node whose whole horizon is fluid-like (in-box, FLUID / INLET / OUTLET)
__global__ void __launch_bounds__(256)
k_nbr_fluid_only(Lat L, long long own_lo, long long own_n, const uint8_t* __restrict__ type,
                 const OffEntry* __restrict__ off, int n_off, uint8_t* __restrict__ nbf) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= own_n) return;
    const long long l = own_lo + t;
    uint8_t ok = 0;
    if (type[l] == PDGPU_FLUID) {
        const int q = (int)(l % L.P);
        const int jj = q / L.Nx, ii = q - jj * L.Nx;
        ok = 1;
        for (int o = 0; o < n_off; ++o) {
            const long long nn = nbr_local(L, off[o], 3, ii, jj, l, type);
            if (nn < 0) { ok = 0; break; }
            const uint8_t tj = type[nn];
            if (!(tj == PDGPU_FLUID || tj == PDGPU_INLET || tj == PDGPU_OUTLET)) { ok = 0; break; }
        }
    }
    nbf[l] = ok;
}

__global__ void __launch_bounds__(NTHREADS, 2)
k_ard_tile(const __grid_constant__ ArdTileParams q, const __grid_constant__ ColTable T,
           const double* __restrict__ d_dt, const uint8_t* __restrict__ type, const uint8_t* __restrict__ nbf,
           const double* __restrict__ C, const double* __restrict__ w_g, const double* __restrict__ vx,
           const double* __restrict__ vy, const double* __restrict__ vz, double* __restrict__ C_n) {
    extern __shared__ double sm[];
    double* s_C = sm;
    double* s_w = sm + SN;

    // warp w owns row pair w >> 2 and thread layer w & 3: the four warps of a row pair sit on the four SM
    // sub-partitions (warp id mod 4), so a row pair without FLUID nodes (tube rim) idles none of them
    const int tid = (threadIdx.z * TY + threadIdx.y) * TX + threadIdx.x;
    const int tx = tid & (TX - 1), ty = 2 * (tid >> 7) + ((tid >> 4) & 1), tz = (tid >> 5) & 3;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, z0 = q.g.z_lo + blockIdx.z * TZ;
    const int gx = x0 + tx, gy = y0 + ty, zt = z0 + tz * RZ;   // first z-node of this thread
    const bool in_xy = gx < q.g.Nx && gy < q.g.Ny;

    uint8_t nty[RZ];
    bool slow = false;   // a FLUID node of this thread has a solid / wall neighbour
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        const int lz = zt + t;
        nty[t] = 255;
        if (in_xy && lz < q.g.z_hi) {
            const long long l = (long long)lz * q.g.P + (long long)gy * q.g.Nx + gx;
            nty[t] = type[l];
            slow = slow || (nty[t] == PDGPU_FLUID && !nbf[l]);
        }
    }
    // asynchronous staging (see ns_tile.cu); outside the box: C = 0, w = +0 (never used by a full row)
    for (int idx = tid; idx < SN; idx += NTHREADS) {
        const long long l = staged_index(q.g, idx, x0, y0, z0);
        const bool ok = l >= 0;
        const long long ls = ok ? l : 0;
        cp_async8(s_C + idx, C + ls, ok);
        cp_async8(s_w + idx, w_g + ls, ok);
    }
    bool fl[RZ];
    bool any = false;
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        fl[t] = (nty[t] == PDGPU_FLUID);
        any = any || fl[t];
        if (nty[t] != 255 && !fl[t] && nty[t] != PDGPU_SOLID_MG && !(q.skip_wall_copy && nty[t] == PDGPU_WALL)) {
            const long long l = (long long)(zt + t) * q.g.P + (long long)gy * q.g.Nx + gx;
            C_n[l] = C[l];   // src/pd_ard.cpp:86-89 (SOLID_MG rows: k_ard_solid_rows)
        }
    }
    cp_async_wait_all();
    if (!__syncthreads_or(any)) return;
    if (!__any_sync(0xffffffffu, any)) return;

    const int base = (tz * RZ * SY + ty + TR) * SX + (tx + TR);   // node t at base + (t+TR)*SPLANE
    double Ci[RZ];
    unsigned long long wi[RZ];     // own w as an integer: sign bit clear for the FLUID nodes that are updated
    ArdAcc a;
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        const int si = base + (t + TR) * SPLANE;
        Ci[t] = s_C[si];
        wi[t] = (unsigned long long)__double_as_longlong(s_w[si]);
        a.diff[t] = a.gx[t] = a.gy[t] = a.gz[t] = 0.0;
    }
    if (__any_sync(0xffffffffu, slow)) {
#pragma unroll 1
        for (int c = 0; c < NCOL; ++c) {
            const int cb = base + T.off[c];
            const double dI = T.di[c], dJ = T.dj[c];
            const double kap[4] = {T.kap[c][0], T.kap[c][1], T.kap[c][2], T.kap[c][3]};
            const double kz[4] = {T.kz[c][0], T.kz[c][1], T.kz[c][2], T.kz[c][3]};
            const int H = T.h[c];
            if (H == 3) ard_column<3>(s_C, s_w, cb, dI, dJ, kap, kz, Ci, wi, a);
            else if (H == 2) ard_column<2>(s_C, s_w, cb, dI, dJ, kap, kz, Ci, wi, a);
            else ard_column<1>(s_C, s_w, cb, dI, dJ, kap, kz, Ci, wi, a);
        }
    } else {
#pragma unroll 1
        for (int c = 0; c < NCOL; ++c) {
            const int cb = base + T.off[c];
            const double dI = T.di[c], dJ = T.dj[c];
            const double kap[4] = {T.kap[c][0], T.kap[c][1], T.kap[c][2], T.kap[c][3]};
            const double kz[4] = {T.kz[c][0], T.kz[c][1], T.kz[c][2], T.kz[c][3]};
            const int H = T.h[c];
            if (H == 3) ard_column_fluid<3>(s_C, s_w, cb, dI, dJ, kap, kz, Ci, wi, a);
            else if (H == 2) ard_column_fluid<2>(s_C, s_w, cb, dI, dJ, kap, kz, Ci, wi, a);
            else ard_column_fluid<1>(s_C, s_w, cb, dI, dJ, kap, kz, Ci, wi, a);
        }
    }

    const double dt = *d_dt;
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        if (!fl[t]) continue;
        const long long l = (long long)(zt + t) * q.g.P + (long long)gy * q.g.Nx + gx;
        // diff and G were accumulated with kappa = dx*w2: e w1 = d kappa, w2 = kappa/dx
        const double adv = q.div_coeff * (vx[l] * a.gx[t] + vy[l] * a.gy[t] + vz[l] * a.gz[t]);
        const double cn = Ci[t] + dt * (q.beta * (a.diff[t] * q.inv_dx) - adv);   // src/pd_ard.cpp:184-189
        C_n[l] = cn < 0.0 ? 0.0 : cn;
    }
}

// SOLID_MG rows (src/pd_ard.cpp:81-190 with i_is_solid): only bonds to fluid-like neighbours,
// D = dsol_i, no advection, no artificial diffusion.
__global__ void __launch_bounds__(128)
k_ard_solid_rows(Lat L, const int* __restrict__ l_solid, long long n_solid, const uint8_t* __restrict__ type,
                 const OffEntry* __restrict__ off, int n_off, const double* __restrict__ d_dt, double beta,
                 const double* __restrict__ C, const double* __restrict__ dsol, double* __restrict__ C_n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_solid) return;
    const long long l = l_solid[t];
    const int q = (int)(l % L.P);
    const int jj = q / L.Nx, ii = q - jj * L.Nx;
    const double C_i = C[l], D = dsol[l];
    double diff = 0.0;
    for (int o = 0; o < n_off; ++o) {
        const long long nn = nbr_local(L, off[o], 3, ii, jj, l, type);
        if (nn < 0) continue;
        const uint8_t tj = type[nn];
        if (tj == PDGPU_FLUID || tj == PDGPU_INLET || tj == PDGPU_OUTLET) diff += beta * D * (C[nn] - C_i) * off[o].w2;
    }
    const double cn = C_i + (*d_dt) * diff;
    C_n[l] = cn < 0.0 ? 0.0 : cn;
}

}  // namespace

// nbfast flags of the owned nodes (pd_rebuild_tables: node types changed)
int pd_build_nbfast(pdgpu_ctx* c) {
    if (c->dim != 3) return 0;
    if (!c->nbfast) {
        CUDA_OK(cudaMalloc(&c->nbfast, c->NL));
        CUDA_OK(cudaMemsetAsync(c->nbfast, 0, c->NL, c->stream));
    }
    const long long own_n = c->own_hi - c->own_lo;
    Lat L = make_lat(c);
    LAUNCH(c, k_nbr_fluid_only, nblocks(own_n, 256), 256, 0, L, c->own_lo, own_n, c->type, c->d_off, c->n_off, c->nbfast);
    return 0;
}

// SOLID_MG rows of the tiled / streaming kernels (all solids of the list, any plane)
int pd_enqueue_ard_solid_rows(pdgpu_ctx* c, int srcC, const double* d_dt) {
    if (!c->n_solid) return 0;
    PdConsts k = pd_consts(c->cfg, c->dim);
    Lat L = make_lat(c);
    LAUNCH(c, k_ard_solid_rows, nblocks(c->n_solid, 128), 128, 0, L, c->l_solid, c->n_solid, c->type, c->d_off,
           c->n_off, d_dt, k.beta_lap, c->C[srcC], c->dsol, c->C[1 - srcC]);
    return 0;
}

// returns -1 when the tiled kernel does not apply. Expects vmag (= vmf) and dsol to be current.
int pd_enqueue_ard_tile(pdgpu_ctx* c, int buf, int srcC, const double* d_dt, int zb, int ze, bool do_solid,
                        bool skip_wall_copy) {
    if (!c->full_rows) return -1;
    if (pd_stream_prepare(c)) return -1;   // column table per context (stream.cuh)
    stream::TileState* ts = pd_tile_state(c);
    ColTable& T = ts->tcols;
    PdConsts k = pd_consts(c->cfg, c->dim);
    ArdTileParams q;
    q.g = make_geom(c);
    q.skip_wall_copy = skip_wall_copy ? 1 : 0;
    if (zb >= 0) { q.g.z_lo = zb; q.g.z_hi = ze; }
    q.beta = k.beta_lap; q.div_coeff = k.alpha / k.V_H; q.inv_dx = 1.0 / c->cfg.dx;
    const size_t smem = sizeof(double) * 2 * SN;
    if (!ts->attr_tile_ard) {   // per context: the attribute is per device
        CUDA_OK(cudaFuncSetAttribute(k_ard_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ts->attr_tile_ard = true;
    }
    int dstC = 1 - srcC;
    if (q.g.z_hi > q.g.z_lo) {
        dim3 grid((c->Nx + TX - 1) / TX, (c->Ny + TY - 1) / TY, ((q.g.z_hi - q.g.z_lo) + TZ - 1) / TZ);
        dim3 block(TX, TY, NZT);
        k_ard_tile<<<grid, block, smem, c->stream>>>(q, T, d_dt, c->type, c->nbfast, c->C[srcC], c->wpack, c->v[buf][0],
                                                      c->v[buf][1], c->v[buf][2], c->C[dstC]);
        c->launches++;
    }
    if (c->n_solid && do_solid) PD_TRY(pd_enqueue_ard_solid_rows(c, srcC, d_dt));
    return 0;
}
