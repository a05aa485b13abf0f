// ns_stream.cu -- z-streaming PD-NS bond kernel (3D, m_ratio = 3, full FLUID rows): the default
// NS kernel.  Schedule, staging and thread layout: stream.cuh.  Bond algebra and numerics are those
// of ns_tile.cu (reference: PD_NS_Solver::step, src/pd_ns.cpp:86-179):
//   g      = rho_j (v_j . d) kappa             mass flux through the bond
//   mc    += g                                  mass convection
//   md    += rho_j kappa                        density Laplacian (x 1/dx at the end)
//   h      = (mu beta/dx) kappa - (alpha/V_H) g momentum convection and viscous Laplacian share one weight
//   a_d   += v_jd h
//   P_d   += p_j d_d kappa                      pressure gradient (di/dj parts factored out per column)
// with d = (di,dj,dk), kappa = dx w2; the f_i parts of the reference's difference form vanish for
// the full symmetric stencil (odd sums) or are added analytically (Laplacians).
#include <algorithm>

#include "stream.cuh"

namespace {
using namespace stream;

constexpr int NF = 5;                       // rho, p, vx, vy, vz
constexpr int MSLOT = NF * MFS;             // doubles per plane slot
constexpr int NACC = 8;                     // accumulators per node
constexpr size_t NS_STREAM_SMEM = sizeof(double) * ((size_t)MRING * MSLOT + (size_t)2 * NACC * MGROUP) + 48;

struct NsStreamParams {
    double rho_f, gamma, B;
    double c_div, dens_diff, visc, rho_lo, rho_hi, W2, inv_dx, md_scale;
    int gamma_is_7;
    StreamGeom g;          // lattice, plane range, work items (stream.cuh)
    const double* f[NF];
    double* o[NF];
};

__device__ __forceinline__ double eos_stream(double rho, const NsStreamParams& q) {
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

// one (di,dj) column: window planes q = zz + TR of this thread, po[q] = ring offset of that plane,
// cbA / cbB = in-plane offset for even / odd q (row alignment parity, stream.cuh)
template <int H>
__device__ __forceinline__ void ns_column(const double* __restrict__ ring, const int (&po)[RZ + 2 * TR], int cbA,
                                          int cbB, double dI, double dJ, const double (&kap)[4],
                                          const double (&kz)[4], const double (&nk)[4], double c_div, NsAcc& a) {
    double colp[RZ];
#pragma unroll
    for (int t = 0; t < RZ; ++t) colp[t] = 0.0;
#pragma unroll
    for (int zz = -H; zz < RZ + H; ++zz) {
        const int qq = zz + TR;
        const double* s = ring + po[qq] + ((qq & 1) ? cbB : cbA);
        const double rj = s[0], pj = s[MFS], ux = s[2 * MFS], uy = s[3 * MFS], uz = s[4 * MFS];
        const double mz = rj * uz;
        const double axy = rj * fma(dJ, uy, dI * ux);
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

// columns [cb0, cb1) of the table (cb0 / cb1 warp-uniform)
__device__ __forceinline__ void ns_columns(const double* __restrict__ ring, const int (&po)[RZ + 2 * TR],
                                           const StreamCols& T, const int cb0, const int cb1, const int ctr,
                                           const int c0, const int par_x, const int par_p, const double c_div,
                                           NsAcc& a) {
#pragma unroll 1
    for (int c = cb0; c < cb1; ++c) {
        const int e = c0 ^ (T.djodd[c] & par_x);
        const int cb = ctr + T.off[c];
        const int cbA = cb + e, cbB = cb + (e ^ par_p);
        const double dI = T.di[c], dJ = T.dj[c];
        const double kap[4] = {T.kap[c][0], T.kap[c][1], T.kap[c][2], T.kap[c][3]};
        const double kz[4] = {T.kz[c][0], T.kz[c][1], T.kz[c][2], T.kz[c][3]};
        const double nk[4] = {T.aux[c][0], T.aux[c][1], T.aux[c][2], T.aux[c][3]};
        const int H = T.h[c];
        if (H == 3) ns_column<3>(ring, po, cbA, cbB, dI, dJ, kap, kz, nk, c_div, a);
        else if (H == 2) ns_column<2>(ring, po, cbA, cbB, dI, dJ, kap, kz, nk, c_div, a);
        else ns_column<1>(ring, po, cbA, cbB, dI, dJ, kap, kz, nk, c_div, a);
    }
}

// Synchronisation is point to point only, so the 16 warps drift apart and the integer / memory
// phases of some overlap the FP64 loops of the others:
//   full[2]   copies of a step's planes have landed: every thread arrives through
//             cp.async.mbarrier.arrive.noinc after issuing its share (512 arrivals per phase)
//   empty[2]  every compute warp has finished reading a step's planes (16 arrivals per phase)
//   named barriers 1..8: a warp and its partner in the other column group (exchange of partial sums)
// Loads and steps are numbered through the whole life of the CTA (k = steps of earlier items + step):
// step k uses full[k & 1] / empty[k & 1] in phase (k >> 1) & 1.  The planes of step s+1 are
// requested in the middle of step s, after every warp has left step s-1 (their slots).
__global__ void __launch_bounds__(MTHREADS, 1)
k_ns_stream(const __grid_constant__ NsStreamParams q, const __grid_constant__ StreamCols T,
            const double* __restrict__ d_dt, const uint8_t* __restrict__ type) {
    extern __shared__ __align__(128) double sm[];
    double* ring = sm;
    double* comb = sm + MRING * MSLOT;
    unsigned long long* full = (unsigned long long*)(comb + 2 * NACC * MGROUP);
    unsigned long long* empty = full + 2;
    volatile int* next_item = (volatile int*)(empty + 2);   // [2]: work item of sequence number n in slot n & 1

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&full[0], MTHREADS);
        mbar_init(&full[1], MTHREADS);
        mbar_init(&empty[0], MTHREADS / 32);
        mbar_init(&empty[1], MTHREADS / 32);
        mbar_fence_init();
        next_item[0] = atomicAdd(q.g.work, 1);
    }
    __syncthreads();
    const int n_items = q.g.ntiles * q.g.nchunks;
    // Warp w owns the row pair w >> 2 of the tile, column group (w >> 1) & 1 and thread layer w & 1.  The four
    // warps of a row pair therefore sit on the four different SM sub-partitions (warp id mod 4): a row pair
    // without FLUID nodes (tube rim) thins every sub-partition out by one warp instead of idling one of them.
    const int warp = tid >> 5, lane = tid & 31;
    const int grp = (warp >> 1) & 1, tz = warp & 1;
    const int tx = lane & (TX - 1), ty = 2 * (warp >> 2) + (lane >> 4);
    const int t = (tz * TY + ty) * TX + tx;           // node-owning thread index inside a column group
    const int pair_bar = 1 + 2 * (warp >> 2) + tz;    // named barrier of this warp and its partner in the other group
    const double dt = *d_dt;
    const double vW = q.visc * q.W2;
    const int ctr = (ty + TR) * MPITCH + (tx + TR);   // in-plane offset of the thread's own (x, y)
    unsigned k0 = 0;

    // Work items are handed out dynamically (atomic counter): CTAs that start late -- the outlet sweep
    // of the side stream holds two SMs while this kernel starts -- simply take fewer items.  Thread 0
    // fetches item n+1 at the start of item n; the other warps read it when they finish item n, which
    // they cannot do before thread 0's arrivals on the full barriers of item n (after the fetch).
    for (unsigned seq = 0;; ++seq) {
        const int item = next_item[seq & 1];
        if (item >= n_items) break;
        if (tid == 0) next_item[(seq + 1) & 1] = atomicAdd(q.g.work, 1);
        const StreamItem it = stream_item(q.g, item);
        const int gx = it.x0 + tx, gy = it.y0 + ty;
        const bool in_xy = gx < q.g.Nx && gy < q.g.Ny;
        const long long lxy = (long long)gy * q.g.Nx + gx;
        // alignment parity of the thread's own row in window plane 0 (4 s and 2 tz are even)
        const int c0 = (int)(it.ebase & 1) ^ (((ty + TR) & 1) & q.g.par_x);

        // this thread's pieces of a staged plane: 5 fields x 14 rows x 12 pieces = 840 per plane
        CopyDesc cd[2];
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int j = tid + d * MTHREADS;
            const int row = j / (MPITCH / 2), cc = j - row * (MPITCH / 2);
            const int f = row / MROWS, r = row - f * MROWS;
            const double* base = f == 0 ? q.f[0] : f == 1 ? q.f[1] : f == 2 ? q.f[2] : f == 3 ? q.f[3] : q.f[4];
            const long long e = it.ebase + (long long)r * q.g.Nx;
            cd[d].src = base + e + 2 * cc;
            cd[d].par = (int)(e & 1);
            cd[d].dst = j < NF * MROWS * (MPITCH / 2) ? f * MFS + r * MPITCH + 2 * cc : -1;
        }
        // planes [pl_lo, pl_hi) of this item -> ring (16-byte asynchronous copies, aligned down to an
        // even element: stream.cuh); this thread's arrival on `bar` fires when its copies have landed
        auto issue = [&](int pl_lo, int pl_hi, unsigned long long* bar) {
            pl_hi = min(pl_hi, it.np);
            for (int pl = pl_lo; pl < pl_hi; ++pl) {
                const long long poff = (long long)pl * q.g.P;
                const int slot = (pl % MRING) * MSLOT;
                const int pp = pl & q.g.par_p;
#pragma unroll
                for (int d = 0; d < 2; ++d)
                    if (cd[d].dst >= 0) cp_async16(ring + slot + cd[d].dst, cd[d].src + poff - (cd[d].par ^ pp));
            }
            cp_async_mbar_arrive(bar);
        };
        // the ring is free once every warp has left the previous item's last step
        if (k0 > 0) mbar_wait(&empty[(k0 - 1) & 1], ((k0 - 1) >> 1) & 1);
        issue(0, MWIN, &full[k0 & 1]);

        // node types of the thread's two nodes in the coming step (255 = not a node of this launch)
        uint8_t nty[RZ];
#pragma unroll
        for (int tn = 0; tn < RZ; ++tn) {
            const int rel = 2 * tz + tn;
            nty[tn] = (in_xy && rel < it.len) ? type[(long long)(it.z0 + rel) * q.g.P + lxy] : (uint8_t)255;
        }

        for (int s = 0; s < it.nsteps; ++s) {
            const unsigned k = k0 + s;
            bool any = false;
#pragma unroll
            for (int tn = 0; tn < RZ; ++tn) any = any || nty[tn] == PDGPU_FLUID;
            const uint8_t my_type = grp ? nty[1] : nty[0];
            const bool warp_any = __any_sync(0xffffffffu, any);
            // types of the next step: the loads fly during the bond loop
            const int rel_own = MS * s + 2 * tz + grp;
#pragma unroll
            for (int tn = 0; tn < RZ; ++tn) {
                const int rel = MS * (s + 1) + 2 * tz + tn;
                nty[tn] = (in_xy && rel < it.len) ? type[(long long)(it.z0 + rel) * q.g.P + lxy] : (uint8_t)255;
            }
            // ring offsets of the thread's window planes (4 s + 2 tz + q)
            int po[RZ + 2 * TR];
            {
                const int w = (MS * s + 2 * tz) % MRING;
#pragma unroll
                for (int qq = 0; qq < RZ + 2 * TR; ++qq) {
                    int sl = w + qq;
                    if (sl >= MRING) sl -= MRING;
                    po[qq] = sl * MSLOT;
                }
            }
            mbar_wait(&full[k & 1], (k >> 1) & 1);
            NsAcc a;
#pragma unroll
            for (int tn = 0; tn < RZ; ++tn)
                a.mc[tn] = a.md[tn] = a.ax[tn] = a.ay[tn] = a.az[tn] = a.px[tn] = a.py[tn] = a.pz[tn] = 0.0;
            // the column range is selected by a branch so that the loop counter -- and with it every
            // weight operand (uniform registers) -- stays warp-uniform for the compiler
            if (warp_any) {
                if (grp == 0) ns_columns(ring, po, T, T.beg[0], T.mid[0], ctr, c0, q.g.par_x, q.g.par_p, q.c_div, a);
                else ns_columns(ring, po, T, T.beg[1], T.mid[1], ctr, c0, q.g.par_x, q.g.par_p, q.c_div, a);
            }
            // middle of the step: request the planes of step s+1.  Their slots held the first planes of
            // step s-1, which every warp has normally left by now (the wait only holds a warp that runs
            // more than half a step ahead).  Requesting at the start of the step instead measured slower
            // (2.90 vs 2.85 ms): the warps re-align.
            if (s + 1 < it.nsteps) {
                if (s > 0) mbar_wait(&empty[(k - 1) & 1], ((k - 1) >> 1) & 1);
                issue(MWIN + MS * s, MWIN + MS * (s + 1), &full[(k + 1) & 1]);
            }
            if (warp_any) {
                if (grp == 0) ns_columns(ring, po, T, T.mid[0], T.end[0], ctr, c0, q.g.par_x, q.g.par_p, q.c_div, a);
                else ns_columns(ring, po, T, T.mid[1], T.end[1], ctr, c0, q.g.par_x, q.g.par_p, q.c_div, a);
            }
            // own values of the node this thread finalises (z-node `grp` of the pair): window plane TR + grp
            double own[NF];
            {
                // window plane TR + grp: TR is odd, so plane TR + 0 is an odd and TR + 1 an even window plane
                const double* sp = ring + (grp ? po[TR + 1] + ctr + c0 : po[TR] + ctr + (c0 ^ q.g.par_p));
#pragma unroll
                for (int f = 0; f < NF; ++f) own[f] = sp[f * MFS];
            }
            // this warp is done with the planes of step k
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[k & 1]);

            if (warp_any) {
                // exchange of partial sums with the partner warp (same nodes, other column half)
                named_bar_sync(pair_bar, 64);   // the partner has consumed the sums of the previous step
                double* cw = comb + (grp * NACC) * MGROUP + t;
#define PD_OTHER(x) (grp ? x[0] : x[1])
                cw[0 * MGROUP] = PD_OTHER(a.mc); cw[1 * MGROUP] = PD_OTHER(a.md); cw[2 * MGROUP] = PD_OTHER(a.ax);
                cw[3 * MGROUP] = PD_OTHER(a.ay); cw[4 * MGROUP] = PD_OTHER(a.az); cw[5 * MGROUP] = PD_OTHER(a.px);
                cw[6 * MGROUP] = PD_OTHER(a.py); cw[7 * MGROUP] = PD_OTHER(a.pz);
#undef PD_OTHER
                named_bar_sync(pair_bar, 64);   // sums of this step are visible
            }
            if (my_type != 255) {
                const long long l = (long long)(it.z0 + rel_own) * q.g.P + lxy;
                if (my_type == PDGPU_FLUID) {
                    const double* cr = comb + ((1 - grp) * NACC) * MGROUP + t;
#define PD_MINE(x) (grp ? x[1] : x[0])
                    const double mc = PD_MINE(a.mc) + cr[0 * MGROUP], md = PD_MINE(a.md) + cr[1 * MGROUP];
                    const double ax = PD_MINE(a.ax) + cr[2 * MGROUP], ay = PD_MINE(a.ay) + cr[3 * MGROUP];
                    const double az = PD_MINE(a.az) + cr[4 * MGROUP], px = PD_MINE(a.px) + cr[5 * MGROUP];
                    const double py = PD_MINE(a.py) + cr[6 * MGROUP], pz = PD_MINE(a.pz) + cr[7 * MGROUP];
#undef PD_MINE
                    const double rho_i = own[0], vi0 = own[2], vi1 = own[3], vi2 = own[4];
                    const double mass_diff = md * q.inv_dx - rho_i * q.W2;
                    double rn = rho_i + dt * (-q.c_div * mc + q.dens_diff * mass_diff);   // src/pd_ns.cpp:160-168
                    rn = fmin(fmax(rn, q.rho_lo), q.rho_hi);
                    q.o[0][l] = rn;
                    q.o[1][l] = eos_stream(rn, q);
                    const double sc = dt / rho_i;                                          // :171-178
                    q.o[2][l] = vi0 + sc * (ax - q.c_div * px - vW * vi0);
                    q.o[3][l] = vi1 + sc * (ay - q.c_div * py - vW * vi1);
                    q.o[4][l] = vi2 + sc * (az - q.c_div * pz - vW * vi2);
                } else {   // copy-through (src/pd_ns.cpp:93-97)
#pragma unroll
                    for (int f = 0; f < NF; ++f) q.o[f][l] = own[f];
                }
            }
        }
        k0 += it.nsteps;
    }
}

// flag[tile] = 1 when the 16 x 8 column of nodes holds a non-OUTSIDE node in any local plane
__global__ void __launch_bounds__(256)
k_tile_active(int Nx, int Ny, int nlp, long long P, const uint8_t* __restrict__ type, int* __restrict__ flag) {
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int tx = threadIdx.x & (TX - 1), ty = (threadIdx.x >> 4) & (TY - 1), tz = threadIdx.x >> 7;
    const int gx = x0 + tx, gy = y0 + ty;
    int any = 0;
    if (gx < Nx && gy < Ny)
        for (int z = tz; z < nlp; z += 2)
            if (type[(long long)z * P + (long long)gy * Nx + gx] != PDGPU_OUTSIDE) { any = 1; break; }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) flag[blockIdx.y * gridDim.x + blockIdx.x] = any;
}

}  // namespace

stream::TileState* pd_tile_state(pdgpu_ctx* c) {
    if (!c->tile_state) {
        stream::TileState* s = new stream::TileState();
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, c->device) == cudaSuccess) s->sm_count = prop.multiProcessorCount;
        if (s->sm_count <= 0) s->sm_count = 148;
        c->tile_state = s;
    }
    return (stream::TileState*)c->tile_state;
}

void pd_tile_state_free(pdgpu_ctx* c) {
    stream::TileState* s = (stream::TileState*)c->tile_state;
    if (!s) return;
    if (s->d_tiles) cudaFree(s->d_tiles);
    if (s->d_work) cudaFree(s->d_work);
    delete s;
    c->tile_state = nullptr;
}

int pd_stream_prepare(pdgpu_ctx* c) {
    stream::TileState* s = pd_tile_state(c);
    if (s->epoch == c->types_epoch) return s->cols_ok ? 0 : -1;
    s->epoch = c->types_epoch;
    s->cols_ok = false;
    if (c->dim != 3) return -1;
    if (!s->d_work) CUDA_OK(cudaMalloc(&s->d_work, sizeof(int) * 16));   // here: never inside a stream capture
    if (!tile::build_columns(c, &s->tcols, &s->sum_kappa)) return -1;
    if (!build_stream_cols(s->tcols, &s->cols)) return -1;
    const int ntx = (c->Nx + TX - 1) / TX, nty = (c->Ny + TY - 1) / TY;
    int* d_flag = nullptr;
    CUDA_OK(cudaMalloc(&d_flag, sizeof(int) * ntx * nty));
    k_tile_active<<<dim3(ntx, nty), 256, 0, c->stream>>>(c->Nx, c->Ny, c->nlp, c->P, c->type, d_flag);
    std::vector<int> flag(ntx * nty), tiles;
    CUDA_OK(cudaMemcpyAsync(flag.data(), d_flag, sizeof(int) * ntx * nty, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d_flag));
    for (int by = 0; by < nty; ++by)
        for (int bx = 0; bx < ntx; ++bx)
            if (flag[by * ntx + bx]) tiles.push_back(bx | (by << 16));
    if (s->d_tiles) { CUDA_OK(cudaFree(s->d_tiles)); s->d_tiles = nullptr; }
    s->ntiles = (int)tiles.size();
    CUDA_OK(cudaMalloc(&s->d_tiles, sizeof(int) * std::max<size_t>(tiles.size(), 1)));
    if (!tiles.empty())
        CUDA_OK(cudaMemcpy(s->d_tiles, tiles.data(), sizeof(int) * tiles.size(), cudaMemcpyHostToDevice));
    s->cols_ok = true;
    return 0;
}

// returns -1 when the streaming kernel does not apply (caller falls back)
int pd_enqueue_ns_stream(pdgpu_ctx* c, int src, const double* d_dt, int zb, int ze) {
    if (!c->full_rows || c->dim != 3 || c->cfg.m_ratio != 3) return -1;
    if (pd_stream_prepare(c)) return -1;
    stream::TileState* s = pd_tile_state(c);
    PdConsts k = pd_consts(c->cfg, c->dim);
    NsStreamParams q;
    q.g.zb = zb >= 0 ? zb : c->R;
    q.g.ze = zb >= 0 ? ze : c->R + (c->a1 - c->a0);
    if (q.g.ze <= q.g.zb || s->ntiles == 0) return 0;
    q.rho_f = c->cfg.rho_f; q.gamma = c->cfg.gamma_eos; q.B = k.B_eos;
    q.c_div = k.alpha * k.inv_VH; q.dens_diff = k.dens_diff_coeff; q.visc = c->cfg.mu_f * k.beta_lap;
    q.rho_lo = 0.5 * c->cfg.rho_f; q.rho_hi = 2.0 * c->cfg.rho_f;
    q.inv_dx = 1.0 / c->cfg.dx;
    q.W2 = s->sum_kappa * q.inv_dx;
    q.gamma_is_7 = (c->cfg.gamma_eos == 7.0);
    q.g.Nx = c->Nx; q.g.Ny = c->Ny; q.g.P = c->P;
    q.g.par_p = (int)(c->P & 1); q.g.par_x = c->Nx & 1;
    q.g.tiles = s->d_tiles; q.g.ntiles = s->ntiles;
    // planes per work item: long chunks amortise the 10-plane fill, short ones balance the CTAs;
    // aim at >= 6 items per CTA
    const int planes = q.g.ze - q.g.zb;
    int zc = c->opt_stream_chunk > 0 ? c->opt_stream_chunk : 32;
    while (zc > 8 && (long long)s->ntiles * ((planes + zc - 1) / zc) < 6LL * s->sm_count) zc -= 4;
    q.g.zc = zc;
    q.g.nchunks = (planes + zc - 1) / zc;
    const int dst = 1 - src;
    q.f[0] = c->rho[src]; q.f[1] = c->p[src]; q.f[2] = c->v[src][0]; q.f[3] = c->v[src][1]; q.f[4] = c->v[src][2];
    q.o[0] = c->rho[dst]; q.o[1] = c->p[dst]; q.o[2] = c->v[dst][0]; q.o[3] = c->v[dst][1]; q.o[4] = c->v[dst][2];
    stream::StreamCols& T = s->cols;
    // kernel weights: aux = viscous weight
    q.md_scale = q.inv_dx / q.c_div;
    stream::StreamCols K = T;
    for (int col = 0; col < NCOL; ++col)
        for (int kk = 0; kk < 4; ++kk) {
            K.aux[col][kk] = q.visc * q.inv_dx * T.kap[col][kk];
        }
    if (!s->attr_ns) {
        CUDA_OK(cudaFuncSetAttribute(k_ns_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NS_STREAM_SMEM));
        s->attr_ns = true;
    }
    const long long items = (long long)q.g.ntiles * q.g.nchunks;
    const unsigned grid = (unsigned)std::min<long long>(items, s->sm_count);
    // one counter per launch in flight (the two plane ranges of a loop body run on two streams)
    q.g.work = s->d_work + (s->work_seq++ & 15);
    CUDA_OK(cudaMemsetAsync(q.g.work, 0, sizeof(int), c->stream));
    k_ns_stream<<<grid, MTHREADS, NS_STREAM_SMEM, c->stream>>>(q, K, d_dt, c->type);
    c->launches++;
    return 0;
}
