// ard_tile.cu -- tiled fast path of the explicit PD-ARD bond kernel (3D, m_ratio = 3, full rows).
//
// Same staging scheme as ns_tile.cu (38 x 14 x 10 haloed block, 4 z-nodes per thread, sliding
// z-window, fully unrolled bond loop with constant-bank weights).  Staged per node:
//     C      concentration
//     vmf    |v| for fluid-like nodes (FLUID / INLET / OUTLET), -1 otherwise
//     dsol   interface diffusivity 2 D_l D_s / (D_l + D_s) of a SOLID_MG node (0 when salt
//            blocked; grain-boundary / precipitate / grain class and the volume-loss decay
//            folded in), 0 for every other type                      (src/pd_ard.cpp:136-162)
// so the bond classification of src/pd_ard.cpp:120-181 becomes branch free for a FLUID row:
//     D_ij = f_j (D_l + alpha dx max(|v_i|, |v_j|)) + dsol_j ,   f_j = [vmf_j >= 0]
//     diff += D_ij (C_j - C_i) w2 ;   G += f_j (C_j - C_i) e w1 ;   adv = (alpha/V_H) v_i . G
// (WALL neighbours have f = 0, dsol = 0 and drop out).  SOLID_MG rows (2 % of the nodes)
// are done by k_ard_solid_rows from the solid node list.
#include <algorithm>
#include <utility>

#include "common.cuh"

namespace {

constexpr int TR = 3;
constexpr int TX = 32, TY = 8, RZ = 4;
constexpr int SX = TX + 2 * TR, SY = TY + 2 * TR, SZ = RZ + 2 * TR;
constexpr int SPLANE = SX * SY, SN = SPLANE * SZ;
constexpr int NTHREADS = TX * TY;

struct ArdWeights {
    double w[7][7][7][4];   // { e_x w1, e_y w1, e_z w1, w2 }
};

struct ArdTileParams {
    int Nx, Ny, nlp, z_lo, z_hi;
    long long P;
    double D_liquid, alpha_dx, beta, div_coeff;
};

struct ArdAcc {
    double diff[RZ], gx[RZ], gy[RZ], gz[RZ];
};

template <int DI, int DJ>
__device__ __forceinline__ void ard_column(const double* __restrict__ s_C, const double* __restrict__ s_vmf,
                                           const double* __restrict__ s_ds, int base, const ArdWeights& W,
                                           const ArdTileParams& q, const double (&Ci)[RZ], const double (&vmi)[RZ],
                                           ArdAcc& a) {
    constexpr int r2 = 12 - DI * DI - DJ * DJ;
    if constexpr (r2 >= 0) {
        constexpr int H = (r2 >= 9) ? 3 : (r2 >= 4) ? 2 : (r2 >= 1) ? 1 : 0;
        const int cb = base + DJ * SX + DI;
#pragma unroll
        for (int zz = -H; zz < RZ + H; ++zz) {
            const int si = cb + (zz + TR) * SPLANE;
            const double Cj = s_C[si], vmf = s_vmf[si], dsj = s_ds[si];
            const double f = vmf >= 0.0 ? 1.0 : 0.0;
#pragma unroll
            for (int t = 0; t < RZ; ++t) {
                const int dk = zz - t;
                if (dk >= -H && dk <= H && !(DI == 0 && DJ == 0 && dk == 0)) {
                    const double* w = W.w[dk + 3][DJ + 3][DI + 3];
                    const double dC = Cj - Ci[t];
                    const double Dff = fma(q.alpha_dx, fmax(vmi[t], vmf), q.D_liquid);
                    const double D = fma(f, Dff, dsj);
                    a.diff[t] = fma(D * dC, w[3], a.diff[t]);
                    const double fd = f * dC;
                    a.gx[t] = fma(fd, w[0], a.gx[t]);
                    a.gy[t] = fma(fd, w[1], a.gy[t]);
                    a.gz[t] = fma(fd, w[2], a.gz[t]);
                }
            }
        }
    }
}

template <int... Is>
__device__ __forceinline__ void ard_all_columns(std::integer_sequence<int, Is...>, const double* s_C,
                                                const double* s_vmf, const double* s_ds, int base,
                                                const ArdWeights& W, const ArdTileParams& q,
                                                const double (&Ci)[RZ], const double (&vmi)[RZ], ArdAcc& a) {
    (ard_column<(Is % 7) - 3, (Is / 7) - 3>(s_C, s_vmf, s_ds, base, W, q, Ci, vmi, a), ...);
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_ard_tile(const ArdTileParams q, const __grid_constant__ ArdWeights W, const double* __restrict__ d_dt,
           const uint8_t* __restrict__ type, const double* __restrict__ C, const double* __restrict__ vmf_g,
           const double* __restrict__ dsol_g, const double* __restrict__ vx, const double* __restrict__ vy,
           const double* __restrict__ vz, double* __restrict__ C_n) {
    extern __shared__ double sm[];
    double* s_C = sm;
    double* s_vmf = sm + SN;
    double* s_ds = sm + 2 * SN;

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * TX + tx;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, z0 = q.z_lo + blockIdx.z * RZ;
    const int gx = x0 + tx, gy = y0 + ty;
    const bool in_xy = gx < q.Nx && gy < q.Ny;

    bool fl[RZ];
    bool any = false;
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        const int lz = z0 + t;
        fl[t] = false;
        if (in_xy && lz < q.z_hi) {
            const long long l = (long long)lz * q.P + (long long)gy * q.Nx + gx;
            const uint8_t ty_ = type[l];
            if (ty_ == PDGPU_FLUID) {
                fl[t] = true;
                any = true;
            } else if (ty_ != PDGPU_SOLID_MG) {
                C_n[l] = C[l];   // src/pd_ard.cpp:86-89 (SOLID_MG rows: k_ard_solid_rows)
            }
        }
    }
    if (!__syncthreads_or(any)) return;

    for (int idx = tid; idx < SN; idx += NTHREADS) {
        const int sz = idx / SPLANE;
        const int rem = idx - sz * SPLANE;
        const int sy = rem / SX;
        const int sx = rem - sy * SX;
        const int ax = x0 - TR + sx, ay = y0 - TR + sy, az = z0 - TR + sz;
        double cc = 0.0, vm = -1.0, ds = 0.0;
        if (ax >= 0 && ax < q.Nx && ay >= 0 && ay < q.Ny && az < q.nlp) {
            const long long l = (long long)az * q.P + (long long)ay * q.Nx + ax;
            cc = __ldg(C + l); vm = __ldg(vmf_g + l); ds = __ldg(dsol_g + l);
        }
        s_C[idx] = cc; s_vmf[idx] = vm; s_ds[idx] = ds;
    }
    __syncthreads();
    if (!__any_sync(0xffffffffu, any)) return;

    const int base = (ty + TR) * SX + (tx + TR);
    double Ci[RZ], vmi[RZ];
    ArdAcc a;
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        const int si = base + (t + TR) * SPLANE;
        Ci[t] = s_C[si];
        vmi[t] = fmax(s_vmf[si], 0.0);
        a.diff[t] = a.gx[t] = a.gy[t] = a.gz[t] = 0.0;
    }
    ard_all_columns(std::make_integer_sequence<int, 49>{}, s_C, s_vmf, s_ds, base, W, q, Ci, vmi, a);

    const double dt = *d_dt;
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        if (!fl[t]) continue;
        const long long l = (long long)(z0 + t) * q.P + (long long)gy * q.Nx + gx;
        const double adv = q.div_coeff * (vx[l] * a.gx[t] + vy[l] * a.gy[t] + vz[l] * a.gz[t]);
        const double cn = Ci[t] + dt * (q.beta * a.diff[t] - adv);   // src/pd_ard.cpp:184-189
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

// returns -1 when the tiled kernel does not apply. Expects vmag (= vmf) and dsol to be current.
int pd_enqueue_ard_tile(pdgpu_ctx* c, int buf, int srcC, const double* d_dt) {
    if (c->dim != 3 || c->cfg.m_ratio != 3 || c->n_off != 178 || !c->full_rows) return -1;
    static ArdWeights W;
    memset(&W, 0, sizeof(W));
    for (const OffEntry& e : c->h_off) {
        if (e.di * e.di + e.dj * e.dj + e.dk * e.dk > 12) return -1;
        double* w = W.w[e.dk + 3][e.dj + 3][e.di + 3];
        w[0] = e.ex * e.w1; w[1] = e.ey * e.w1; w[2] = e.ez * e.w1; w[3] = e.w2;
    }
    PdConsts k = pd_consts(c->cfg, c->dim);
    ArdTileParams q;
    q.Nx = c->Nx; q.Ny = c->Ny; q.nlp = c->nlp; q.z_lo = c->R; q.z_hi = c->R + (c->a1 - c->a0);
    q.P = c->P;
    q.D_liquid = c->cfg.D_liquid; q.alpha_dx = c->cfg.alpha_art_diff * c->cfg.dx;
    q.beta = k.beta_lap; q.div_coeff = k.alpha / k.V_H;
    const size_t smem = sizeof(double) * 3 * SN;
    static bool attr_done = false;
    if (!attr_done) {
        CUDA_OK(cudaFuncSetAttribute(k_ard_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    int dstC = 1 - srcC;
    dim3 grid((c->Nx + TX - 1) / TX, (c->Ny + TY - 1) / TY, ((c->a1 - c->a0) + RZ - 1) / RZ);
    dim3 block(TX, TY, 1);
    k_ard_tile<<<grid, block, smem, c->stream>>>(q, W, d_dt, c->type, c->C[srcC], c->vmag, c->dsol, c->v[buf][0],
                                                  c->v[buf][1], c->v[buf][2], c->C[dstC]);
    c->launches++;
    if (c->n_solid) {
        Lat L = make_lat(c);
        LAUNCH(c, k_ard_solid_rows, nblocks(c->n_solid, 128), 128, 0, L, c->l_solid, c->n_solid, c->type, c->d_off,
               c->n_off, d_dt, k.beta_lap, c->C[srcC], c->dsol, c->C[dstC]);
    }
    return 0;
}
