// ns_split_p.cu -- pressure-gradient half of the split PD-NS bond kernel (option ns_kernel = 4).
//
// The pressure terms of src/pd_ns.cpp:115-157 are linear in p: P_d = sum_j p_j d_d kappa (2.6 of the 12.3
// FP64 ops per bond of ns_tile.cu).  Taking them -- and the p field -- out of the main kernel leaves four
// staged fields there, which at 16 x 8 x 4 tiles fit TWO CTAs per SM (ns_split_a.cu): one CTA's staging then
// overlaps the other's bond loop, which is what the one-CTA kernel cannot do.  This kernel stages p alone
// (34.5 KB, four CTAs per SM) and leaves (alpha/V_H) P_d of every FLUID node in the NEW velocity buffers,
// where ns_split_a.cu picks it up in its epilogue.
#include "tile.cuh"   // 16 x 8 x 8 tiles (PD_TILE_NZT = 4)

namespace {
using namespace tile;

struct PAcc { double px[RZ], py[RZ], pz[RZ]; };

template <int H>
__device__ __forceinline__ void p_column(const double* __restrict__ s_p, int cb, double dI, double dJ,
                                         const double (&kap)[4], const double (&kz)[4], PAcc& a) {
    double colp[RZ];
#pragma unroll
    for (int t = 0; t < RZ; ++t) colp[t] = 0.0;
#pragma unroll
    for (int zz = -H; zz < RZ + H; ++zz) {
        const double pj = s_p[cb + (zz + TR) * SPLANE];
#pragma unroll
        for (int t = 0; t < RZ; ++t) {
            const int dk = zz - t;
            if (dk >= -H && dk <= H) {
                const int ak = dk < 0 ? -dk : dk;
                colp[t] = fma(pj, kap[ak], colp[t]);
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

__global__ void __launch_bounds__(NTHREADS, 4)
k_ns_pgrad(const __grid_constant__ TileGeom g, const __grid_constant__ ColTable T, double c_div,
           const uint8_t* __restrict__ type, const double* __restrict__ pr, double* __restrict__ vx_n,
           double* __restrict__ vy_n, double* __restrict__ vz_n) {
    extern __shared__ double s_p[];
    const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
    const int tid = (tz * TY + ty) * TX + tx;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, z0 = g.z_lo + blockIdx.z * TZ;
    const int gx = x0 + tx, gy = y0 + ty, zt = z0 + tz * RZ;
    const bool in_xy = gx < g.Nx && gy < g.Ny;
    bool fl[RZ];
    bool any = false;
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        const int lz = zt + t;
        fl[t] = in_xy && lz < g.z_hi && type[(long long)lz * g.P + (long long)gy * g.Nx + gx] == PDGPU_FLUID;
        any = any || fl[t];
    }
    for (int idx = tid; idx < SN; idx += NTHREADS) {
        const long long l = staged_index(g, idx, x0, y0, z0);
        cp_async8(s_p + idx, pr + (l >= 0 ? l : 0), l >= 0);
    }
    cp_async_wait_all();
    if (!__syncthreads_or(any)) return;
    if (!__any_sync(0xffffffffu, any)) return;
    PAcc a;
#pragma unroll
    for (int t = 0; t < RZ; ++t) a.px[t] = a.py[t] = a.pz[t] = 0.0;
    const int base = (tz * RZ * SY + ty + TR) * SX + (tx + TR);
#pragma unroll 1
    for (int c = 0; c < NCOL; ++c) {
        const int cb = base + T.off[c];
        const double dI = T.di[c], dJ = T.dj[c];
        const double kap[4] = {T.kap[c][0], T.kap[c][1], T.kap[c][2], T.kap[c][3]};
        const double kz[4] = {T.kz[c][0], T.kz[c][1], T.kz[c][2], T.kz[c][3]};
        const int H = T.h[c];
        if (H == 3) p_column<3>(s_p, cb, dI, dJ, kap, kz, a);
        else if (H == 2) p_column<2>(s_p, cb, dI, dJ, kap, kz, a);
        else p_column<1>(s_p, cb, dI, dJ, kap, kz, a);
    }
#pragma unroll
    for (int t = 0; t < RZ; ++t) {
        if (!fl[t]) continue;
        const long long l = (long long)(zt + t) * g.P + (long long)gy * g.Nx + gx;
        vx_n[l] = c_div * a.px[t];
        vy_n[l] = c_div * a.py[t];
        vz_n[l] = c_div * a.pz[t];
    }
}

}  // namespace

// pressure half over the local plane range [zb, ze); -1 when not applicable
int pd_enqueue_ns_split_p(pdgpu_ctx* c, int src, int zb, int ze) {
    if (!c->full_rows) return -1;
    static ColTable T;
    double sum_kappa = 0.0;
    if (!build_columns(c, &T, &sum_kappa)) return -1;
    PdConsts k = pd_consts(c->cfg, c->dim);
    TileGeom g = make_geom(c);
    if (zb >= 0) { g.z_lo = zb; g.z_hi = ze; }
    if (g.z_hi <= g.z_lo) return 0;
    const size_t smem = sizeof(double) * SN;
    static bool attr_done = false;
    if (!attr_done) {
        CUDA_OK(cudaFuncSetAttribute(k_ns_pgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    const int dst = 1 - src;
    dim3 grid((c->Nx + TX - 1) / TX, (c->Ny + TY - 1) / TY, ((g.z_hi - g.z_lo) + TZ - 1) / TZ);
    dim3 block(TX, TY, NZT);
    k_ns_pgrad<<<grid, block, smem, c->stream>>>(g, T, k.alpha * k.inv_VH, c->type, c->p[src], c->v[dst][0],
                                                  c->v[dst][1], c->v[dst][2]);
    c->launches++;
    return 0;
}
