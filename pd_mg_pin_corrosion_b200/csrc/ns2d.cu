// ns2d.cu -- the 2D flow loop (BASELINE configs 1-2: solve_steady runs 19 000 - 25 000 iterations on a
// lattice of 2 x 10^4 nodes) as ONE persistent kernel per batch of iterations.
//
// With one launch per operator an iteration of PD_NS_Solver::solve_steady (src/pd_ns.cpp:196-205: inlet,
// outlet, wall, solid-surface BCs, step, wall mirror of the new buffers) costs 112 us on a B200 although the
// lattice holds 1 us of FP64 work: the in-place outlet sweep (src/boundary.cpp:88-131) alone is a 1024-thread
// CTA running ~100 block barriers.  Here the whole lattice lives in the shared memory of a cooperative grid:
//
//   * CTA b < n_step owns a band of whole lattice rows (axial planes) and keeps the node types and the staging
//     sources of the band + R halo rows in shared memory for the whole launch.
//   * phase 1: inlet BC (one warp per INLET node, loads in parallel, sum in CSR order) and solid-surface BC on
//     the step CTAs; the OUTLET sweep on a dedicated CTA: the outlet rows and the R fluid rows below them are
//     staged once, the pre-pass sums are taken from shared memory, and the Gauss-Seidel recurrence is walked by
//     one lane per (outlet row, field) with a lag of R+1 nodes between successive rows -- Nx + (R+1)(KP-1)
//     warp-synchronous steps, the sum over the rows below software-pipelined one node ahead so that the
//     recurrence carries one DADD and one multiply / division per step.
//   * grid barrier
//   * phase 2: the band + halo of the four flow fields is staged through L2 (ld.global.cg: other SMs wrote
//     them); the wall mirror (src/boundary.cpp:266-283) is folded into the staging as a gather from the mirror
//     node (and written back for the owned WALL nodes, so the buffers hold what the reference's hold); the
//     bond sums of a node are split over `split` adjacent lanes; new FLUID values and the copy-through of the
//     other nodes go to the second buffer; the channel-flow corrections (src/pd_ns.cpp:209-270) reduce each
//     owned row inside the CTA.
//   * grid barrier.  The wall mirror of the NEW buffers (:205) is only observable after the last iteration of
//     a launch (the next iteration's wall BC overwrites it before anything reads a WALL node), so it runs once.
//
// Two grid barriers per iteration (a monotonic counter in L2, one atomic + one polling thread per CTA).
// Applicability: 2D, one rank, no wall mirror on a SOLID node (the reference mirrors before it zeroes the
// solid surface), the band fits shared memory.  Otherwise the per-operator path (ns.cu) runs.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

constexpr int NT2D = 512;
constexpr int MAXR2D = 5;     // m_ratio <= 5 (api.cu)
constexpr int MAXKP2D = 16;   // outlet rows (one lane per row and field)

struct Off2 {
    int di, dj, ds, pad;      // ds = dj * Nx + di
    double ex, ey, w1, w2;
};

struct Ns2dParams {
    int Nx, R, Na;            // local rows: Na + 2R, owned rows [R, R + Na)
    int n_step;               // CTAs that own rows; CTA n_step = outlet CTA (KP > 0)
    int n_off, n_early, split;
    int out_row0, KP;         // first outlet row (local), number of outlet rows
    int out_rows_staged;      // rows [out_row0 - R, out_row0 - R + out_rows_staged)
    int sweep_parts;          // lanes per T value of the sweep (4, or 2 for more than 8 outlet rows)
    int channel;              // channel_flow_corrections
    int iters, src0;
    double rho_f, gamma, B, c_div, dens_diff, visc, rho_lo, rho_hi;
    double C_in, U_in;
    const double* dt;
    double* rho[2];
    double* p[2];
    double* vx[2];
    double* vy[2];
    double* C;
    const uint8_t* type;
    const int* wsrc;          // per node: WALL -> mirror index or -1 (none); others -2
    const OffEntry* off;
    const int* l_inlet;
    const double* inlet_vax;
    int n_inlet;
    const int* l_solid;
    int n_solid;
    const int* l_wall;
    const int* mirror;
    int n_wall;
    unsigned* bar;
    // ARD loop (k_ard2d_loop)
    double* Cb[2];            // concentration buffers; step `it` reads Cb[cs0 ^ (it & 1)]
    int cs0, fb;              // first source C buffer, flow buffer (frozen)
    const int* l_ssolid;      // SOLID_MG nodes with a fluid-like neighbour
    int n_ssolid;
    const uint8_t *is_gb, *is_precip;
    uint8_t* salt;
    double *dsol, *wpack;
    const double* dt_ard;
    double D_liquid, D_grain, D_gb, D_precip, decay, alpha_dx, beta, div_coeff, C_sat;
    unsigned long long* prof; // optional: globaltimer stamps of the last iteration, 8 per CTA
};

struct Ns2dState {
    long long epoch = -1;
    bool ok = false;
    int n_step = 0, split = 1, sm_count = 0;
    int out_row0 = 0, KP = 0, out_rows_staged = 0;
    int* wsrc = nullptr;
    unsigned* bar = nullptr;
    unsigned long long* prof = nullptr;
    size_t smem = 0;
    bool attr_set = false, attr_set_ard = false;
    size_t attr_smem = 0, attr_smem_ard = 0;
};

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        int spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (++spins > (1 << 22)) __trap();   // a lost CTA must not hang the device
        } while (v < target);
    }
    __syncthreads();
}

// Division by a small positive integer n on the dependent path of the sweep: q = RN(a * RN(1/n)) is within one
// ulp, the exact remainder r = a - q n comes from one FMA and RN(q + r * RN(1/n)) is the correctly rounded
// quotient (Markstein; checked exhaustively for n <= 130 on 4 x 10^8 operands); tiny |a| take the division.
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void stamp(const Ns2dParams& q, bool on, int slot) {
    if (on && q.prof && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        q.prof[blockIdx.x * 8 + slot] = t;
    }
}

__device__ __forceinline__ int row_lo(const Ns2dParams& q, int b) {
    return q.R + (int)(((long long)b * q.Na) / q.n_step);
}

// -------------------------------------------------------------------- outlet CTA ----------
// apply_outlet_bc (src/boundary.cpp:88-131) on one CTA.  Static per launch: for every OUTLET node the bit mask
// of the offsets that are summed from the OLD values (FLUID neighbours and lexicographically LATER outlet
// neighbours), 1/n and n of the FLUID|OUTLET neighbour count.  Per iteration:
//   staging   the outlet rows and the R rows below / above them through L2;
//   pre-pass  base = sum over the masked offsets in CSR order (one thread per node, loads independent of the
//             accumulation chain);
//   sweep     the Gauss-Seidel recurrence in batches of SWB steps.  Row kp trails row kp-1 by LAG = R + SWB
//             nodes, so everything a batch needs from the rows below was final before the batch began:
//             all threads first form T = base + (already swept values of the rows below, CSR order) for the
//             KP x SWB nodes of the batch; then one warp per field walks the SWB steps with one lane per outlet
//             row: x_i = (T_i + x_{i-R} + .. + x_{i-1}) / n, the newest value last, i.e. one DADD and one
//             multiply (velocity, :113-124) or exact small division (concentration, :129) on the dependent path.
// The swept values live in rows of pitch Nx + 2R with R zero rows below row 0 and R zero columns on both
// sides, so a neighbour below is one shared-memory word at a fixed distance from the node.
constexpr int SWB = 8;

struct OutSmem {
    double *v, *c;            // staged axial velocity / concentration (out_rows_staged x Nx)
    double *base_v, *base_c;  // pre-pass sums (KP x Nx)
    double *new_v, *new_c;    // swept values, 0 on non-OUTLET nodes ((KP + R) x (Nx + 2R), zero padded)
    double *rc, *dn;          // 1/n and n per node of the outlet rows (0 / 1: not an OUTLET node or n = 0)
    double *T;                // [field][step of the batch][row]
    double *brc, *bdn;        // [step of the batch][row]
    int* cnt;                 // FLUID|OUTLET neighbours, -1: not an OUTLET node
    int* rel;                 // distance dj * (Nx + 2R) + di of the offsets below the row
    unsigned* mask;           // [node][word]: offsets summed by the pre-pass
    uint8_t* t;               // staged types
};

__device__ __forceinline__ int mask_words(const Ns2dParams& q) { return (q.n_off + 31) >> 5; }

__device__ __forceinline__ OutSmem out_smem(const Ns2dParams& q, unsigned char* raw) {
    OutSmem s;
    const int ns = q.out_rows_staged * q.Nx, nk = q.KP * q.Nx;
    const int npad = (q.KP + q.R) * (q.Nx + 2 * q.R);
    double* d = (double*)raw;
    s.v = d; d += ns;
    s.c = d; d += ns;
    s.base_v = d; d += nk;
    s.base_c = d; d += nk;
    s.new_v = d; d += npad;
    s.new_c = d; d += npad;
    s.rc = d; d += nk;
    s.dn = d; d += nk;
    s.T = d; d += 2 * SWB * MAXKP2D;
    s.brc = d; d += SWB * MAXKP2D;
    s.bdn = d; d += SWB * MAXKP2D;
    s.cnt = (int*)d;
    s.rel = s.cnt + nk;
    s.mask = (unsigned*)(s.rel + q.n_off);
    s.t = (uint8_t*)(s.mask + (size_t)nk * mask_words(q));
    return s;
}

// SWB steps of the recurrence of one (field, outlet row): x_i = (T_i + x_{i-R} + .. + x_{i-1}) * (1/n) resp. / n.
// Everything but the newest value is added one step ahead (`pre`), so the dependent path per step is one DADD
// and one DMUL (velocity) or DMUL + 2 DFMA (concentration: Markstein's correction, exact quotient).
// SWB steps of the recurrence of one (field, outlet row): x_i = (T_i + x_{i-R} + .. + x_{i-1}) * (1/n) resp. / n.
// Everything but the newest value is added one step ahead (`pre`), so the dependent path per step is one DADD
// and one DMUL (velocity) or DMUL + 2 DFMA (concentration: q = RN(tot * RN(1/n)) is within one ulp, the exact
// remainder tot - q n comes from one FMA and RN(q + r * RN(1/n)) is the correctly rounded quotient -- Markstein;
// checked against the division for n <= 130 on 4 x 10^8 operands).  No branch inside the steps: totals below
// the range of the correction (|tot| < 1e-280, not zero) only raise a flag and the batch is redone with the
// division (EXACT = true; out of line, practically never).
template <int RR, bool EXACT>
__device__ __forceinline__ bool chain_steps(const double (&T)[SWB], const double (&rc)[SWB], const double (&dn)[SWB],
                                            double* row, int i0, int Nx, int fld, double (&prev)[MAXR2D]) {
    bool tiny = false;
    double pre = T[0];
#pragma unroll
    for (int x = RR - 1; x >= 1; --x) pre += prev[x];            // di = -R .. -2
#pragma unroll
    for (int u = 0; u < SWB; ++u) {
        double pre_n = 0.0;
        if (u + 1 < SWB) {
            pre_n = T[u + 1];
#pragma unroll
            for (int x = RR - 1; x >= 1; --x) pre_n += prev[x - 1];
        }
        const double tot = pre + prev[0];                         // di = -1: the newest value last
        const double qv = tot * rc[u];                            // src/boundary.cpp:113-124 (rc = 0: not an OUTLET node)
        double val = qv;
        if (fld) {                                                // :129, sc / cnt
            if (EXACT) val = rc[u] != 0.0 ? tot / dn[u] : 0.0;
            else {
                val = fma(fma(-qv, dn[u], tot), rc[u], qv);
                tiny |= fabs(tot) < 1e-280 && tot != 0.0 && rc[u] != 0.0;
            }
        }
        const int i = i0 + u;
        if (i >= 0 && i < Nx) row[i] = val;
#pragma unroll
        for (int x = RR - 1; x > 0; --x) prev[x] = prev[x - 1];
        prev[0] = val;
        pre = pre_n;
    }
    return tiny;
}

template <int RR>
__device__ __noinline__ void chain_redo(const double (&T)[SWB], const double (&rc)[SWB], const double (&dn)[SWB],
                                        double* row, int i0, int Nx, int fld, double (&prev)[MAXR2D]) {
    chain_steps<RR, true>(T, rc, dn, row, i0, Nx, fld, prev);
}

template <int RR>
__device__ __forceinline__ void chain_batch(const OutSmem& s, double* row, int i0, int Nx, int fld, int kp,
                                            double (&prev)[MAXR2D]) {
    double T[SWB], rc[SWB], dn[SWB];
#pragma unroll
    for (int u = 0; u < SWB; ++u) {
        T[u] = s.T[(fld * SWB + u) * MAXKP2D + kp];
        rc[u] = s.brc[u * MAXKP2D + kp];
        dn[u] = s.bdn[u * MAXKP2D + kp];
    }
    double saved[MAXR2D];
#pragma unroll
    for (int x = 0; x < MAXR2D; ++x) saved[x] = prev[x];
    if (chain_steps<RR, false>(T, rc, dn, row, i0, Nx, fld, prev)) {
#pragma unroll
        for (int x = 0; x < MAXR2D; ++x) prev[x] = saved[x];
        chain_redo<RR>(T, rc, dn, row, i0, Nx, fld, prev);
    }
}

// once per launch (node types are fixed while the flow loop runs)
__device__ void outlet_setup(const Ns2dParams& q, const OutSmem& s, const Off2* s_off) {
    const int tid = threadIdx.x, Nx = q.Nx, R = q.R, PW = Nx + 2 * R;
    const int ns = q.out_rows_staged * Nx, nk = q.KP * Nx, MW = mask_words(q);
    const long long g0 = (long long)(q.out_row0 - R) * Nx;
    for (int k = tid; k < ns; k += NT2D) s.t[k] = q.type[g0 + k];
    for (int k = tid; k < (q.KP + R) * PW; k += NT2D) { s.new_v[k] = 0.0; s.new_c[k] = 0.0; }
    for (int o = tid; o < q.n_off; o += NT2D) s.rel[o] = s_off[o].dj * PW + s_off[o].di;
    __syncthreads();
    for (int d = tid; d < nk; d += NT2D) {
        const int kp = d / Nx, i = d - kp * Nx;
        const int k = (kp + R) * Nx + i;
        int c = -1;
        unsigned m[4] = {0u, 0u, 0u, 0u};
        if (s.t[k] == PDGPU_OUTLET) {
            c = 0;
            for (int o = 0; o < q.n_off; ++o) {
                const Off2 e = s_off[o];
                const int ni = i + e.di, k2 = k + e.ds;
                if ((unsigned)ni >= (unsigned)Nx || k2 < 0 || k2 >= ns) continue;
                const uint8_t tj = s.t[k2];
                if (tj == PDGPU_FLUID || (tj == PDGPU_OUTLET && o >= q.n_early)) { m[o >> 5] |= 1u << (o & 31); ++c; }
                else if (tj == PDGPU_OUTLET) ++c;
            }
        }
        for (int w = 0; w < MW; ++w) s.mask[d * MW + w] = m[w];
        s.cnt[d] = c;
        s.rc[d] = c > 0 ? 1.0 / (double)c : 0.0;
        s.dn[d] = c > 0 ? (double)c : 1.0;
    }
}

__device__ void outlet_phase(const Ns2dParams& q, const OutSmem& s, const Off2* s_off, int S, bool last, double* Cb) {
    const int tid = threadIdx.x, Nx = q.Nx, R = q.R, PW = Nx + 2 * R;
    const long long g0 = (long long)(q.out_row0 - R) * Nx;   // global index of the first staged node
    const int ns = q.out_rows_staged * Nx, nk = q.KP * Nx, MW = mask_words(q);
    const double* vy = q.vy[S];
    for (int k = tid; k < ns; k += NT2D) {
        s.v[k] = __ldcg(vy + g0 + k);
        s.c[k] = __ldcg(Cb + g0 + k);
    }
    __syncthreads();
    stamp(q, last, 5);
    // pre-pass (:96-111)
    for (int d = tid; d < nk; d += NT2D) {
        const int kp = d / Nx;
        const int k = d + R * Nx;
        double sv = 0.0, sc = 0.0;
        for (int w = 0; w < MW; ++w) {
            unsigned m = s.mask[d * MW + w];
            const int lim = min(32, q.n_off - 32 * w);
            for (int u = 0; u < lim; ++u) {
                const int k2 = k + s_off[32 * w + u].ds;
                if ((m >> u) & 1u) { sv += s.v[k2]; sc += s.c[k2]; }
            }
        }
        (void)kp;
        s.base_v[d] = sv; s.base_c[d] = sc;
    }
    __syncthreads();
    stamp(q, last, 6);
    // sweep
    const int LAG = R + SWB;
    const int steps = Nx + LAG * (q.KP - 1);
    const int n_below = q.n_early - R;      // the earlier half ends with the R in-row offsets
    // T phase roles (fixed for the whole sweep): PARTS adjacent lanes per (row, step of the batch, field)
    const int PARTS = q.sweep_parts, n_items = q.KP * SWB * 2 * PARTS;
    const bool t_valid = tid < n_items;
    const int t_part = tid % PARTS, t_rest = tid / PARTS;
    const int t_fld = t_rest & 1, t_u = (t_rest >> 1) % SWB, t_kp = t_valid ? (t_rest >> 1) / SWB : 0;
    const double* t_base = (t_fld ? s.base_c : s.base_v) + t_kp * Nx;
    const double* t_rc = s.rc + t_kp * Nx;
    const double* t_dn = s.dn + t_kp * Nx;
    const double* t_row = (t_fld ? s.new_c : s.new_v) + (t_kp + R) * PW + R;
    const int t_i0 = t_u - LAG * t_kp;
    double* t_T = s.T + (t_fld * SWB + t_u) * MAXKP2D + t_kp;
    double* t_brc = s.brc + t_u * MAXKP2D + t_kp;
    double* t_bdn = s.bdn + t_u * MAXKP2D + t_kp;
    const int ne = (n_below + PARTS - 1) / PARTS;
    int rel[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int o = t_part + e * PARTS;
        rel[e] = (o < n_below) ? s.rel[o] : -(t_kp + 1) * PW;      // else: a zero row
    }
    // recurrence roles: warp 0 = velocity, warp 1 = concentration, lane = outlet row
    double prev[MAXR2D];
#pragma unroll
    for (int u = 0; u < MAXR2D; ++u) prev[u] = 0.0;
    const int fld_w = tid >> 5, kp_w = tid & 31;
    double* c_row = (fld_w ? s.new_c : s.new_v) + (min(kp_w, q.KP - 1) + R) * PW + R;
    long long c_help = 0, c_chain = 0;
    for (int b0 = 0; b0 < steps; b0 += SWB) {
        const long long c0 = clock64();
        {
            // (addresses of steps outside [0, Nx) stay inside the CTA's shared memory; what they read is dropped)
            const int i = b0 + t_i0;
            const bool in = t_valid && i >= 0 && i < Nx;
            const double* p = t_row + i;
            const double bse = t_base[i], rcv = t_rc[i], dnv = t_dn[i];
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                if (e < ne) a0 += p[rel[e]];
                if (e + 1 < ne) a1 += p[rel[e + 1]];
            }
            double a = a0 + a1;
            if (PARTS == 4) { a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2); }
            else a += __shfl_xor_sync(0xffffffffu, a, 1);
            if (t_valid && t_part == 0) {
                t_T[0] = in ? bse + a : 0.0;
                if (t_fld == 0) { t_brc[0] = in ? rcv : 0.0; t_bdn[0] = in ? dnv : 1.0; }
            }
        }
        __syncthreads();
        const long long c1 = clock64();
        if (tid < 64 && kp_w < q.KP) {
            const int i0 = b0 - LAG * kp_w;
            switch (R) {
                case 1: chain_batch<1>(s, c_row, i0, Nx, fld_w, kp_w, prev); break;
                case 2: chain_batch<2>(s, c_row, i0, Nx, fld_w, kp_w, prev); break;
                case 3: chain_batch<3>(s, c_row, i0, Nx, fld_w, kp_w, prev); break;
                case 4: chain_batch<4>(s, c_row, i0, Nx, fld_w, kp_w, prev); break;
                default: chain_batch<5>(s, c_row, i0, Nx, fld_w, kp_w, prev); break;
            }
        }
        __syncthreads();
        c_help += c1 - c0; c_chain += clock64() - c1;
    }
    stamp(q, last, 7);
    if (last && q.prof && tid == 0) {
        q.prof[8 * (blockIdx.x + 1)] = c_help; q.prof[8 * (blockIdx.x + 1) + 1] = c_chain;
    }
    double* rho = q.rho[S]; double* pr = q.p[S]; double* vx = q.vx[S]; double* vyw = q.vy[S];
    const long long l0 = (long long)q.out_row0 * Nx;
    for (int d = tid; d < nk; d += NT2D) {
        const int n = s.cnt[d];
        if (n < 0) continue;
        const long long l = l0 + d;
        const int kp = d / Nx, pd = (kp + R) * PW + R + (d - kp * Nx);
        rho[l] = q.rho_f;
        pr[l] = 0.0;            // EOS(rho_f) = 0 exactly
        vx[l] = 0.0;
        vyw[l] = n > 0 ? s.new_v[pd] : q.U_in;   // n = 0: an OUTLET node nobody is a neighbour of
        Cb[l] = n > 0 ? s.new_c[pd] : 0.0;
    }
}

// -------------------------------------------------------------------- step CTAs -----------
struct StepSmem {
    double *rho, *p, *vx, *vy;   // staged band + halo
    double* nrho;                // new density of the owned nodes (channel corrections)
    double* red;                 // channel corrections: 8 warp partials + mean + pressure of the mean
    int* src;                    // staging source: >= 0 node, -1 wall without mirror, <= -2: mirror -2-src (negated v)
    int* redc;
    uint8_t* t;
};

__device__ __forceinline__ StepSmem step_smem(int ns_max, unsigned char* raw) {
    StepSmem s;
    double* d = (double*)raw;
    s.rho = d; d += ns_max;
    s.p = d; d += ns_max;
    s.vx = d; d += ns_max;
    s.vy = d; d += ns_max;
    s.nrho = d; d += ns_max;
    s.red = d; d += 16;
    s.src = (int*)d;
    s.redc = s.src + ns_max;
    s.t = (uint8_t*)(s.redc + 16);
    return s;
}

__device__ __forceinline__ double eos2d(const Ns2dParams& q, double rho) {
    return eos_pressure(rho, q.rho_f, q.gamma, q.B);
}

__device__ void bc_phase(const Ns2dParams& q, const Off2* s_off, int S, double* Cb, bool solid_bc) {
    // apply_inlet_bc (src/boundary.cpp:31-75): warp per node, neighbour loads in parallel, sum in CSR order
    const int lane = threadIdx.x & 31, wpc = NT2D / 32;
    const int gw = blockIdx.x * wpc + (threadIdx.x >> 5), nw = q.n_step * wpc;
    double* rho = q.rho[S];
    for (int t = gw; t < q.n_inlet; t += nw) {
        const int l = q.l_inlet[t];
        const int i = l % q.Nx;
        double s = 0.0;
        int cnt = 0;
        for (int o0 = 0; o0 < q.n_off; o0 += 32) {
            const int o = o0 + lane;
            bool ok = false;
            double rv = 0.0;
            if (o < q.n_off) {
                const Off2 e = s_off[o];
                const int ni = i + e.di;
                if ((unsigned)ni < (unsigned)q.Nx && q.type[l + e.ds] == PDGPU_FLUID) { ok = true; rv = __ldcg(rho + l + e.ds); }
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            const int lim = min(32, q.n_off - o0);
            for (int u = 0; u < lim; ++u) {
                const double v = __shfl_sync(0xffffffffu, rv, u);
                if ((m >> u) & 1u) { s += v; ++cnt; }
            }
        }
        if (lane == 0) {
            const double r = cnt > 0 ? s / cnt : q.rho_f;
            rho[l] = r;
            q.p[S][l] = eos2d(q, r);
            q.vx[S][l] = 0.0;
            q.vy[S][l] = q.inlet_vax[t];
            Cb[l] = q.C_in;
        }
    }
    // apply_solid_surface_bc (src/boundary.cpp:381-390)
    if (solid_bc)
    for (int t = blockIdx.x * NT2D + threadIdx.x; t < q.n_solid; t += q.n_step * NT2D) {
        const int l = q.l_solid[t];
        q.vx[S][l] = 0.0;
        q.vy[S][l] = 0.0;
    }
}

template <int SPLIT>
__device__ void step_phase(const Ns2dParams& q, const StepSmem& s, const Off2* s_off, int S, int r0, int r1) {
    const int tid = threadIdx.x, Nx = q.Nx, R = q.R, D = 1 - S;
    const int ns = (r1 - r0 + 2 * R) * Nx;
    const long long g0 = (long long)(r0 - R) * Nx;
    const int own0 = R * Nx, n_own = (r1 - r0) * Nx;
    const double *rho = q.rho[S], *pr = q.p[S], *vx = q.vx[S], *vy = q.vy[S];
    // ---- staging; WALL nodes take the mirror node's values (apply_wall_mirror_proper, src/boundary.cpp:266-283)
    for (int k = tid; k < ns; k += NT2D) {
        const int sc = s.src[k];
        double a, b, c, d;
        if (sc >= 0) { a = __ldcg(rho + sc); b = __ldcg(pr + sc); c = __ldcg(vx + sc); d = __ldcg(vy + sc); }
        else if (sc == -1) { a = q.rho_f; b = 0.0; c = 0.0; d = 0.0; }
        else {
            const int m = -2 - sc;
            a = __ldcg(rho + m); b = __ldcg(pr + m); c = -__ldcg(vx + m); d = -__ldcg(vy + m);
        }
        s.rho[k] = a; s.p[k] = b; s.vx[k] = c; s.vy[k] = d;
        if (sc < 0 && k >= own0 && k < own0 + n_own) {          // owned WALL node: the buffer holds the BC value
            const long long l = g0 + k;
            q.rho[S][l] = a; q.p[S][l] = b; q.vx[S][l] = c; q.vy[S][l] = d;
        }
    }
    __syncthreads();
    // ---- bond sums (PD_NS_Solver::step, src/pd_ns.cpp:86-179), SPLIT lanes per node
    const double dt = *q.dt;
    const int total = n_own * SPLIT;
    const int per = (q.n_off + SPLIT - 1) / SPLIT;
    for (int w0 = 0; w0 < total; w0 += NT2D) {
        const int w = w0 + tid;
        const bool in = w < total;
        const int node = in ? w / SPLIT : 0, part = w & (SPLIT - 1);
        const int k = own0 + node;
        const bool fluid = in && s.t[k] == PDGPU_FLUID;
        const double rho_i = s.rho[k], p_i = s.p[k], vi0 = s.vx[k], vi1 = s.vy[k];
        double mass_conv = 0.0, mass_diff = 0.0, mc0 = 0.0, mc1 = 0.0, mp0 = 0.0, mp1 = 0.0, mv0 = 0.0, mv1 = 0.0;
        if (fluid) {
            const int i = node % Nx;
            const int o_end = min(q.n_off, (part + 1) * per);
            const double mi0 = rho_i * vi0, mi1 = rho_i * vi1;
            for (int o = part * per; o < o_end; ++o) {
                const Off2 e = s_off[o];
                const int ni = i + e.di, k2 = k + e.ds;
                if ((unsigned)ni >= (unsigned)Nx) continue;
                if (s.t[k2] == PDGPU_OUTSIDE) continue;
                const double rho_j = s.rho[k2], p_j = s.p[k2], vj0 = s.vx[k2], vj1 = s.vy[k2];
                const double mj0 = rho_j * vj0, mj1 = rho_j * vj1;
                mass_conv += ((mj0 - mi0) * e.ex + (mj1 - mi1) * e.ey) * e.w1;                 // :128-130
                mass_diff += (rho_j - rho_i) * e.w2;                                           // :133
                const double c0 = (mj0 * vj0 - mi0 * vi0) * e.ex + (mj0 * vj1 - mi0 * vi1) * e.ey;   // :136-143
                const double c1 = (mj1 * vj0 - mi1 * vi0) * e.ex + (mj1 * vj1 - mi1 * vi1) * e.ey;
                mc0 += c0 * e.w1; mc1 += c1 * e.w1;
                const double dp = (p_j - p_i) * e.w1;                                          // :146-148
                mp0 += dp * e.ex; mp1 += dp * e.ey;
                mv0 += (vj0 - vi0) * e.w2; mv1 += (vj1 - vi1) * e.w2;                          // :151-153
            }
        }
#pragma unroll
        for (int x = 1; x < SPLIT; x <<= 1) {
            mass_conv += __shfl_xor_sync(0xffffffffu, mass_conv, x);
            mass_diff += __shfl_xor_sync(0xffffffffu, mass_diff, x);
            mc0 += __shfl_xor_sync(0xffffffffu, mc0, x);
            mc1 += __shfl_xor_sync(0xffffffffu, mc1, x);
            mp0 += __shfl_xor_sync(0xffffffffu, mp0, x);
            mp1 += __shfl_xor_sync(0xffffffffu, mp1, x);
            mv0 += __shfl_xor_sync(0xffffffffu, mv0, x);
            mv1 += __shfl_xor_sync(0xffffffffu, mv1, x);
        }
        if (in && part == 0) {
            const long long l = g0 + k;
            double rn = rho_i, pn = p_i, v0 = vi0, v1 = vi1;     // :93-97 copy-through of the other node types
            if (fluid) {
                rn = rho_i + dt * (-q.c_div * mass_conv + q.dens_diff * mass_diff);    // :160-168
                rn = fmin(fmax(rn, q.rho_lo), q.rho_hi);
                pn = eos2d(q, rn);
                const double sdt = dt / rho_i;                                        // :171-178
                v0 = vi0 + sdt * (-q.c_div * mc0 - q.c_div * mp0 + q.visc * mv0);
                v1 = vi1 + sdt * (-q.c_div * mc1 - q.c_div * mp1 + q.visc * mv1);
            }
            q.rho[D][l] = rn; q.p[D][l] = pn; q.vx[D][l] = v0; q.vy[D][l] = v1;
            if (q.channel) s.nrho[k] = rn;
        }
    }
}

// Channel-flow corrections (src/pd_ns.cpp:209-270): row mean of the new rho over the FLUID nodes of each owned
// row, reduction shaped like k_channel_corrections (ns.cu).  They follow the wall mirror of the new buffers,
// which reads the UNcorrected fluid values: on the last iteration of a launch (the only one whose mirror is
// observable) the caller runs barrier, mirror, barrier before this.
__device__ void channel_write(const Ns2dParams& q, const StepSmem& s, int S, int r0, int r1) {
    const int tid = threadIdx.x, Nx = q.Nx, R = q.R, D = 1 - S;
    const long long g0 = (long long)(r0 - R) * Nx;
    for (int r = r0; r < r1; ++r) {
        const int kb = (r - r0 + R) * Nx;
        double sum = 0.0;
        int cnt = 0;
        if (tid < 256)
            for (int x = tid; x < Nx; x += 256)
                if (s.t[kb + x] == PDGPU_FLUID) { sum += s.nrho[kb + x]; ++cnt; }
        sum = warp_sum(sum);
        cnt = warp_sum_i(cnt);
        if (tid < 256 && (tid & 31) == 0) { s.red[tid >> 5] = sum; s.redc[tid >> 5] = cnt; }
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            int c = 0;
            for (int w = 0; w < 8; ++w) { t += s.red[w]; c += s.redc[w]; }
            const double avg = c > 0 ? t / c : 0.0;
            s.redc[8] = c;
            s.red[8] = avg;
            s.red[9] = c > 0 ? eos2d(q, avg) : 0.0;
        }
        __syncthreads();
        const bool have = s.redc[8] > 0;
        const double avg = s.red[8], pavg = s.red[9];
        for (int x = tid; x < Nx; x += NT2D) {
            if (s.t[kb + x] != PDGPU_FLUID) continue;
            const long long l = g0 + kb + x;
            q.vx[D][l] = 0.0;
            if (have) { q.rho[D][l] = avg; q.p[D][l] = pavg; }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NT2D, 1) k_ns2d_loop(const __grid_constant__ Ns2dParams q) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int tid = threadIdx.x, b = blockIdx.x, Nx = q.Nx, R = q.R;
    Off2* s_off = (Off2*)raw;
    unsigned char* rest = raw + sizeof(Off2) * q.n_off;
    for (int o = tid; o < q.n_off; o += NT2D) {
        const OffEntry e = q.off[o];
        Off2 f;
        f.di = e.di; f.dj = e.dj; f.ds = e.dj * Nx + e.di; f.pad = 0;
        f.ex = e.ex; f.ey = e.ey; f.w1 = e.w1; f.w2 = e.w2;
        s_off[o] = f;
    }
    const bool is_out = q.KP > 0 && b == q.n_step;
    const unsigned nctas = gridDim.x;
    unsigned target = 0;
    OutSmem so;
    StepSmem ss;
    int r0 = 0, r1 = 0;
    if (is_out) {
        so = out_smem(q, rest);
        __syncthreads();          // s_off
        outlet_setup(q, so, s_off);
    } else {
        r0 = row_lo(q, b); r1 = row_lo(q, b + 1);
        int rows_max = (q.Na + q.n_step - 1) / q.n_step + 2 * R;
        ss = step_smem(rows_max * Nx, rest);
        const long long g0 = (long long)(r0 - R) * Nx;
        const int ns = (r1 - r0 + 2 * R) * Nx;
        for (int k = tid; k < ns; k += NT2D) {
            const uint8_t t = q.type[g0 + k];
            ss.t[k] = t;
            int sc = (int)(g0 + k);
            if (t == PDGPU_WALL) {
                const int m = q.wsrc[g0 + k];
                sc = m >= 0 ? -2 - m : -1;
            }
            ss.src[k] = sc;
        }
    }
    __syncthreads();

    for (int it = 0; it < q.iters; ++it) {
        const int S = q.src0 ^ (it & 1);
        const bool last = it == q.iters - 1;
        stamp(q, last, 0);
        if (is_out) outlet_phase(q, so, s_off, S, last, q.C);
        else bc_phase(q, s_off, S, q.C, true);
        stamp(q, last, 1);
        target += nctas;
        grid_barrier(q.bar, target);
        stamp(q, last, 2);
        if (!is_out) {
            if (q.split == 4) step_phase<4>(q, ss, s_off, S, r0, r1);
            else if (q.split == 2) step_phase<2>(q, ss, s_off, S, r0, r1);
            else step_phase<1>(q, ss, s_off, S, r0, r1);
            if (q.channel && !last) { __syncthreads(); channel_write(q, ss, S, r0, r1); }
        }
        stamp(q, last, 3);
        target += nctas;
        grid_barrier(q.bar, target);
        stamp(q, last, 4);
        if (last) {
            // wall mirror of the new buffers (src/pd_ns.cpp:205)
            const int D = 1 - S;
            for (int t = b * NT2D + tid; t < q.n_wall; t += (int)nctas * NT2D) {
                const int l = q.l_wall[t], m = q.mirror[t];
                if (m >= 0) {
                    q.vx[D][l] = -__ldcg(q.vx[D] + m);
                    q.vy[D][l] = -__ldcg(q.vy[D] + m);
                    q.rho[D][l] = __ldcg(q.rho[D] + m);
                    q.p[D][l] = __ldcg(q.p[D] + m);
                } else {
                    q.vx[D][l] = 0.0; q.vy[D][l] = 0.0; q.rho[D][l] = q.rho_f; q.p[D][l] = 0.0;
                }
            }
            if (q.channel) {
                target += nctas;
                grid_barrier(q.bar, target);
                if (!is_out) channel_write(q, ss, S, r0, r1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// The explicit corrosion loop body (src/coupling.cpp:232-240: inlet, outlet, wall-concentration BCs, ARD step,
// swap C) with the same structure: `steps` bodies per launch, two grid barriers per step.
//   phase 1: inlet BC; salt-layer flag and interface diffusivity of the surface solids (src/pd_ard.cpp:61-73,
//            140-162) from the FLUID concentrations; the outlet CTA sweeps axial velocity and concentration.
//   phase 2: the band + halo of C is staged through L2 together with what changes per step in the second staged
//            value -- |v| of the OUTLET nodes (their velocity was just swept) and the interface diffusivity of the
//            SOLID_MG nodes -- while |v| of the other fluid-like nodes, the node types and the own velocities of the
//            band stay in shared memory for the whole launch (velocities are frozen during the corrosion phase);
//            PD_ARD_Solver::step (src/pd_ard.cpp:81-190) per FLUID / SOLID_MG node in CSR order, copy-through else.
// The wall-concentration BC stays lazy (no bond reads WALL C, :120): the caller keeps it owed as in the
// per-operator loop.
struct ArdSmem {
    double *C, *a;      // staged concentration; |v| (fluid-like) or interface diffusivity (SOLID_MG), -1 otherwise
    double *vx, *vy;    // own velocities of the band (static)
    uint8_t* t;
};
__device__ __forceinline__ ArdSmem ard_smem(int ns_max, int n_own_max, unsigned char* raw) {
    ArdSmem s;
    double* d = (double*)raw;
    s.C = d; d += ns_max;
    s.a = d; d += ns_max;
    s.vx = d; d += n_own_max;
    s.vy = d; d += n_own_max;
    s.t = (uint8_t*)d;
    return s;
}

__device__ void ard_bc_phase(const Ns2dParams& q, const Off2* s_off, int S, double* Cb) {
    bc_phase(q, s_off, S, Cb, false);                     // apply_inlet_bc
    // salt flags / interface diffusivities of the surface solids: one thread per node, order-free
    for (int t = blockIdx.x * NT2D + threadIdx.x; t < q.n_ssolid; t += q.n_step * NT2D) {
        const int l = q.l_ssolid[t];
        const int i = l % q.Nx;
        bool blocked = false;
        for (int o = 0; o < q.n_off && !blocked; ++o) {
            const Off2 e = s_off[o];
            if ((unsigned)(i + e.di) >= (unsigned)q.Nx) continue;
            if (q.type[l + e.ds] == PDGPU_FLUID && __ldcg(Cb + l + e.ds) >= q.C_sat) blocked = true;
        }
        double ds = 0.0;
        if (!blocked) {
            double D_s = q.is_gb[l] ? q.D_gb : (q.is_precip[l] ? q.D_precip : q.D_grain);
            D_s *= q.decay;
            ds = 2.0 * q.D_liquid * D_s / (q.D_liquid + D_s + 1e-30);
        }
        q.salt[l] = blocked ? 1 : 0;
        q.dsol[l] = ds;
        q.wpack[l] = -ds;
    }
}

template <int SPLIT>
__device__ void ard_step_phase(const Ns2dParams& q, const ArdSmem& s, const Off2* s_off, const double* Cs, double* Cd,
                               int r0, int r1) {
    const int tid = threadIdx.x, Nx = q.Nx, R = q.R;
    const int ns = (r1 - r0 + 2 * R) * Nx;
    const long long g0 = (long long)(r0 - R) * Nx;
    const int own0 = R * Nx, n_own = (r1 - r0) * Nx;
    const double* vxg = q.vx[q.fb];
    const double* vyg = q.vy[q.fb];
    for (int k = tid; k < ns; k += NT2D) {
        s.C[k] = __ldcg(Cs + g0 + k);
        const uint8_t t = s.t[k];
        if (t == PDGPU_SOLID_MG) s.a[k] = __ldcg(q.dsol + g0 + k);
        else if (t == PDGPU_OUTLET || t == PDGPU_INLET) {    // velocities (re)written by the BCs of phase 1
            const double a = __ldcg(vxg + g0 + k), b = __ldcg(vyg + g0 + k);
            s.a[k] = sqrt(a * a + b * b);
        }
    }
    __syncthreads();
    const double dt = *q.dt_ard;
    const int total = n_own * SPLIT;
    const int per = (q.n_off + SPLIT - 1) / SPLIT;
    for (int w0 = 0; w0 < total; w0 += NT2D) {
        const int w = w0 + tid;
        const bool in = w < total;
        const int node = in ? w / SPLIT : 0, part = w & (SPLIT - 1);
        const int k = own0 + node;
        const uint8_t ti = s.t[k];
        const bool i_fl = in && ti == PDGPU_FLUID, i_so = in && ti == PDGPU_SOLID_MG;
        const double C_i = s.C[k];
        double diff = 0.0, adv = 0.0;
        if (i_fl || i_so) {
            const int i = node % Nx;
            const double vi0 = i_fl ? s.vx[node] : 0.0, vi1 = i_fl ? s.vy[node] : 0.0;
            const double vi_mag = i_fl ? s.a[k] : 0.0, ds_i = i_so ? s.a[k] : 0.0;
            const int o_end = min(q.n_off, (part + 1) * per);
            for (int o = part * per; o < o_end; ++o) {
                const Off2 e = s_off[o];
                const int k2 = k + e.ds;
                if ((unsigned)(i + e.di) >= (unsigned)Nx) continue;
                const uint8_t tj = s.t[k2];
                if (tj == PDGPU_OUTSIDE || tj == PDGPU_WALL) continue;                       // :120
                const bool j_fl = tj == PDGPU_FLUID || tj == PDGPU_INLET || tj == PDGPU_OUTLET;
                if (!i_fl && !j_fl) continue;                                                // solid-solid :134
                const double dC = s.C[k2] - C_i;
                double D;
                if (i_fl && j_fl) {                                                          // :137-139, :166-181
                    D = q.D_liquid + q.alpha_dx * fmax(vi_mag, s.a[k2]);
                    const double vde = vi0 * e.ex + vi1 * e.ey;
                    adv += dC * vde * e.w1;
                } else {
                    D = i_fl ? s.a[k2] : ds_i;                                               // interface :140-162
                }
                diff += q.beta * D * dC * e.w2;                                              // :173
            }
        }
#pragma unroll
        for (int x = 1; x < SPLIT; x <<= 1) {
            diff += __shfl_xor_sync(0xffffffffu, diff, x);
            adv += __shfl_xor_sync(0xffffffffu, adv, x);
        }
        if (in && part == 0) {
            double cn = C_i;
            if (i_fl || i_so) {
                cn = C_i + dt * (diff - adv * q.div_coeff);
                if (cn < 0.0) cn = 0.0;
            }
            Cd[g0 + k] = cn;
        }
    }
}

__global__ void __launch_bounds__(NT2D, 1) k_ard2d_loop(const __grid_constant__ Ns2dParams q) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int tid = threadIdx.x, b = blockIdx.x, Nx = q.Nx, R = q.R;
    Off2* s_off = (Off2*)raw;
    unsigned char* rest = raw + sizeof(Off2) * q.n_off;
    for (int o = tid; o < q.n_off; o += NT2D) {
        const OffEntry e = q.off[o];
        Off2 f;
        f.di = e.di; f.dj = e.dj; f.ds = e.dj * Nx + e.di; f.pad = 0;
        f.ex = e.ex; f.ey = e.ey; f.w1 = e.w1; f.w2 = e.w2;
        s_off[o] = f;
    }
    const bool is_out = q.KP > 0 && b == q.n_step;
    const unsigned nctas = gridDim.x;
    unsigned target = 0;
    OutSmem so;
    ArdSmem sa;
    int r0 = 0, r1 = 0;
    const int S = q.fb;
    if (is_out) {
        so = out_smem(q, rest);
        __syncthreads();
        outlet_setup(q, so, s_off);
    } else {
        r0 = row_lo(q, b); r1 = row_lo(q, b + 1);
        const int rows_own_max = (q.Na + q.n_step - 1) / q.n_step;
        sa = ard_smem((rows_own_max + 2 * R) * Nx, rows_own_max * Nx, rest);
        const long long g0 = (long long)(r0 - R) * Nx;
        const int ns = (r1 - r0 + 2 * R) * Nx, own0 = R * Nx, n_own = (r1 - r0) * Nx;
        for (int k = tid; k < ns; k += NT2D) {
            const uint8_t t = q.type[g0 + k];
            sa.t[k] = t;
            double a = -1.0;                                  // k_ard_vmag (ard.cu): |v| of the fluid-like nodes
            if (t == PDGPU_FLUID || t == PDGPU_INLET || t == PDGPU_OUTLET) {
                const double u = q.vx[S][g0 + k], v = q.vy[S][g0 + k];
                a = sqrt(u * u + v * v);
            }
            sa.a[k] = a;
        }
        for (int n = tid; n < n_own; n += NT2D) { sa.vx[n] = q.vx[S][g0 + own0 + n]; sa.vy[n] = q.vy[S][g0 + own0 + n]; }
    }
    __syncthreads();
    for (int it = 0; it < q.iters; ++it) {
        double* Cs = q.Cb[q.cs0 ^ (it & 1)];
        double* Cd = q.Cb[q.cs0 ^ (it & 1) ^ 1];
        if (is_out) outlet_phase(q, so, s_off, S, false, Cs);
        else ard_bc_phase(q, s_off, S, Cs);
        target += nctas;
        grid_barrier(q.bar, target);
        if (!is_out) {
            if (q.split == 4) ard_step_phase<4>(q, sa, s_off, Cs, Cd, r0, r1);
            else if (q.split == 2) ard_step_phase<2>(q, sa, s_off, Cs, Cd, r0, r1);
            else ard_step_phase<1>(q, sa, s_off, Cs, Cd, r0, r1);
        }
        target += nctas;
        grid_barrier(q.bar, target);
    }
}

__global__ void k_wsrc(long long NL, int* __restrict__ wsrc) {
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l < NL) wsrc[l] = -2;
}
__global__ void k_wsrc_walls(const int* __restrict__ l_wall, const int* __restrict__ mirror, long long n,
                             const uint8_t* __restrict__ type, int* __restrict__ wsrc, int* __restrict__ n_solid_mirror) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int l = l_wall[t], m = mirror[t];
    wsrc[l] = m >= 0 ? m : -1;
    if (m >= 0 && type[m] == PDGPU_SOLID_MG) atomicAdd(n_solid_mirror, 1);
}

size_t step_bytes(int ns_max, int n_off) {
    return sizeof(Off2) * n_off + sizeof(double) * (5 * (size_t)ns_max + 16) + sizeof(int) * ((size_t)ns_max + 16) + ns_max + 16;
}
size_t out_bytes(int ns, int nk, int npad, int n_off) {
    return sizeof(Off2) * n_off + sizeof(double) * (2 * (size_t)ns + 4 * (size_t)nk + 2 * (size_t)npad + 4 * SWB * MAXKP2D) +
           sizeof(int) * ((size_t)nk + n_off) + sizeof(unsigned) * (size_t)nk * ((n_off + 31) / 32) + ns + 16;
}

}   // namespace

void pd_ns2d_free(pdgpu_ctx* c) {
    Ns2dState* st = (Ns2dState*)c->ns2d_state;
    if (!st) return;
    cudaFree(st->wsrc);
    cudaFree(st->bar);
    cudaFree(st->prof);
    delete st;
    c->ns2d_state = nullptr;
}

// Called from pd_rebuild_tables (may allocate and synchronise): staging sources of the WALL nodes, the band
// partition, applicability.
int pd_ns2d_prepare(pdgpu_ctx* c) {
    if (c->dim != 2 || c->nranks > 1) return 0;
    Ns2dState* st = (Ns2dState*)c->ns2d_state;
    if (!st) { st = new Ns2dState; c->ns2d_state = st; }
    st->ok = false;
    st->epoch = c->types_epoch;
    if (!st->sm_count) {
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, c->device));
        st->sm_count = prop.multiProcessorCount;
    }
    if (!st->bar) CUDA_OK(cudaMalloc(&st->bar, sizeof(unsigned) * 4));
    if (st->wsrc) { CUDA_OK(cudaFree(st->wsrc)); st->wsrc = nullptr; }
    CUDA_OK(cudaMalloc(&st->wsrc, sizeof(int) * c->NL));
    k_wsrc<<<nblocks(c->NL, 256), 256, 0, c->stream>>>(c->NL, st->wsrc);
    CUDA_OK(cudaMemsetAsync(c->d_int, 0, sizeof(int), c->stream));
    if (c->n_wall)
        k_wsrc_walls<<<nblocks(c->n_wall, 256), 256, 0, c->stream>>>(c->l_wall, c->l_wall_mirror, c->n_wall, c->type,
                                                                     st->wsrc, c->d_int);
    int solid_mirrors = 0;
    CUDA_OK(cudaMemcpyAsync(&solid_mirrors, c->d_int, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    int ends[2] = {0, 0};
    if (c->n_outlet) {
        CUDA_OK(cudaMemcpyAsync(&ends[0], c->l_outlet, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaMemcpyAsync(&ends[1], c->l_outlet + (c->n_outlet - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (solid_mirrors > 0) return 0;
    const int Nx = c->Nx, R = c->R, Na = c->a1 - c->a0;
    st->KP = 0; st->out_row0 = 0; st->out_rows_staged = 0;
    if (c->n_outlet) {
        st->out_row0 = ends[0] / Nx;
        st->KP = ends[1] / Nx - st->out_row0 + 1;
        if (st->KP > MAXKP2D || st->out_row0 < R) return 0;
        st->out_rows_staged = std::min(st->KP + 2 * R, Na + 2 * R - (st->out_row0 - R));
    }
    if (R > MAXR2D) return 0;
    // the earlier half of the stencil must end with the R in-row offsets (dj = 0, di = -R..-1)
    const int n_early = c->n_off / 2;
    for (int u = 0; u < R; ++u) {
        const OffEntry& e = c->h_off[n_early - R + u];
        if (e.dj != 0 || e.di != -R + u) return 0;
    }
    for (int o = 0; o < n_early - R; ++o)
        if (c->h_off[o].dj >= 0) return 0;
    if (c->n_off > 128) return 0;   // pre-pass masks: <= 4 words per node
    if (st->KP > 0 && (n_early - R + (st->KP <= 8 ? 4 : 2) - 1) / (st->KP <= 8 ? 4 : 2) > 8) return 0;   // sweep: <= 8 offsets per lane
    st->n_step = std::min(Na, st->sm_count - (st->KP > 0 ? 1 : 0));
    if (st->n_step < 1) return 0;
    const int rows_max = (Na + st->n_step - 1) / st->n_step + 2 * R;
    const int n_own_max = ((Na + st->n_step - 1) / st->n_step) * Nx;
    st->split = (n_own_max * 4 <= NT2D) ? 4 : (n_own_max * 2 <= NT2D) ? 2 : 1;
    size_t smem = step_bytes(rows_max * Nx, c->n_off);
    if (st->KP > 0) smem = std::max(smem, out_bytes(st->out_rows_staged * Nx, st->KP * Nx, (st->KP + R) * (Nx + 2 * R), c->n_off));
    if (smem > 200 * 1024) return 0;
    st->smem = smem;
    if (!st->attr_set || st->attr_smem < smem) {
        CUDA_OK(cudaFuncSetAttribute(k_ns2d_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        st->attr_set = true; st->attr_smem = smem;
    }
    int coop = 0;
    CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
    int per_sm = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ns2d_loop, NT2D, smem));
    if (!coop || per_sm < 1) return 0;
    st->ok = true;
    return 0;
}

bool pd_ns2d_ok(const pdgpu_ctx* c) {
    const Ns2dState* st = (const Ns2dState*)c->ns2d_state;
    return c->opt_ns2d && c->dim == 2 && c->nranks == 1 && st && st->ok && st->epoch == c->types_epoch;
}

// `iters` loop bodies starting from flow buffer `src`; the state afterwards is that of the per-operator path
// after the same bodies (last body read buffer src ^ ((iters-1) & 1), no swap after it).
int pd_enqueue_ns2d(pdgpu_ctx* c, int src, int iters) {
    Ns2dState* st = (Ns2dState*)c->ns2d_state;
    const PdConsts k = pd_consts(c->cfg, c->dim);
    Ns2dParams q;
    q.Nx = c->Nx; q.R = c->R; q.Na = c->a1 - c->a0;
    q.n_step = st->n_step; q.n_off = c->n_off; q.n_early = c->n_off / 2; q.split = st->split;
    q.out_row0 = st->out_row0; q.KP = st->KP; q.out_rows_staged = st->out_rows_staged;
    q.channel = c->cfg.channel_flow_corrections ? 1 : 0;
    q.iters = iters; q.src0 = src;
    q.rho_f = c->cfg.rho_f; q.gamma = c->cfg.gamma_eos; q.B = k.B_eos;
    q.c_div = k.alpha * k.inv_VH; q.dens_diff = k.dens_diff_coeff; q.visc = c->cfg.mu_f * k.beta_lap;
    q.rho_lo = 0.5 * c->cfg.rho_f; q.rho_hi = 2.0 * c->cfg.rho_f;
    q.C_in = c->cfg.C_liquid_init; q.U_in = c->cfg.U_in;
    q.dt = c->d_dt;
    for (int bfr = 0; bfr < 2; ++bfr) {
        q.rho[bfr] = c->rho[bfr]; q.p[bfr] = c->p[bfr]; q.vx[bfr] = c->v[bfr][0]; q.vy[bfr] = c->v[bfr][1];
    }
    q.C = c->C[c->curC];
    q.type = c->type; q.wsrc = st->wsrc; q.off = c->d_off;
    q.l_inlet = c->l_inlet; q.inlet_vax = c->inlet_vax; q.n_inlet = (int)c->n_inlet;
    q.l_solid = c->l_solid; q.n_solid = (int)c->n_solid;
    q.l_wall = c->l_wall; q.mirror = c->l_wall_mirror; q.n_wall = (int)c->n_wall;
    q.bar = st->bar;
    q.sweep_parts = st->KP <= 8 ? 4 : 2;
    static const bool want_prof = getenv("PDGPU_NS2D_PROF") != nullptr;
    q.prof = nullptr;
    if (want_prof) {
        if (!st->prof) CUDA_OK(cudaMalloc(&st->prof, sizeof(unsigned long long) * 8 * (st->sm_count + 2)));
        q.prof = st->prof;
    }
    CUDA_OK(cudaMemsetAsync(st->bar, 0, sizeof(unsigned), c->stream));
    void* args[] = {(void*)&q};
    const unsigned grid = (unsigned)(st->n_step + (st->KP > 0 ? 1 : 0));
    CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_ns2d_loop, dim3(grid), dim3(NT2D), args, st->smem, c->stream));
    c->launches++;
    if (want_prof && iters > 1) {   // phase times of the last iteration: first / middle step CTA, outlet CTA
        std::vector<unsigned long long> h(8 * (grid + 1));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        CUDA_OK(cudaMemcpy(h.data(), st->prof, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
        if (st->KP > 0) {
            fprintf(stderr, "[ns2d prof] sweep: T phases %llu cycles, chain phases %llu cycles\n", h[8 * grid], h[8 * grid + 1]);
            const unsigned long long* t = &h[8 * (grid - 1)];
            fprintf(stderr, "[ns2d prof] outlet cta: staging %6.2f us  pre-pass %6.2f  recurrence %6.2f  write %6.2f\n",
                    (t[5] - t[0]) * 1e-3, (t[6] - t[5]) * 1e-3, (t[7] - t[6]) * 1e-3, (t[1] - t[7]) * 1e-3);
        }
        for (unsigned b : {0u, grid / 2, grid - 1}) {
            const unsigned long long* t = &h[8 * b];
            fprintf(stderr, "[ns2d prof] cta %3u: phase1 %6.2f us  barrier %6.2f  phase2 %6.2f  barrier %6.2f\n", b,
                    (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3, (t[3] - t[2]) * 1e-3, (t[4] - t[3]) * 1e-3);
        }
    }
    return 0;
}

// `steps` corrosion loop bodies (src/coupling.cpp:232-240) from concentration buffer `srcC`, flow buffer `buf`
// frozen; the caller flips curC per step and keeps the lazy wall-concentration BC owed.
int pd_enqueue_ard2d(pdgpu_ctx* c, int buf, int srcC, int steps) {
    Ns2dState* st = (Ns2dState*)c->ns2d_state;
    const PdConsts k = pd_consts(c->cfg, c->dim);
    Ns2dParams q;
    memset(&q, 0, sizeof(q));
    q.Nx = c->Nx; q.R = c->R; q.Na = c->a1 - c->a0;
    q.n_step = st->n_step; q.n_off = c->n_off; q.n_early = c->n_off / 2; q.split = st->split;
    q.out_row0 = st->out_row0; q.KP = st->KP; q.out_rows_staged = st->out_rows_staged;
    q.sweep_parts = st->KP <= 8 ? 4 : 2;
    q.iters = steps; q.src0 = buf;
    q.rho_f = c->cfg.rho_f; q.gamma = c->cfg.gamma_eos; q.B = k.B_eos;
    q.C_in = c->cfg.C_liquid_init; q.U_in = c->cfg.U_in;
    for (int bfr = 0; bfr < 2; ++bfr) {
        q.rho[bfr] = c->rho[bfr]; q.p[bfr] = c->p[bfr]; q.vx[bfr] = c->v[bfr][0]; q.vy[bfr] = c->v[bfr][1];
        q.Cb[bfr] = c->C[bfr];
    }
    q.C = nullptr;
    q.type = c->type; q.wsrc = st->wsrc; q.off = c->d_off;
    q.l_inlet = c->l_inlet; q.inlet_vax = c->inlet_vax; q.n_inlet = (int)c->n_inlet;
    q.l_solid = c->l_solid; q.n_solid = (int)c->n_solid;
    q.bar = st->bar;
    q.prof = nullptr;
    q.cs0 = srcC; q.fb = buf;
    q.l_ssolid = c->l_ssolid; q.n_ssolid = (int)c->n_ssolid;
    q.is_gb = c->is_gb; q.is_precip = c->is_precip; q.salt = c->salt; q.dsol = c->dsol; q.wpack = c->wpack;
    q.dt_ard = c->d_dt + 1;
    q.D_liquid = c->cfg.D_liquid; q.D_grain = c->cfg.D_grain; q.D_gb = c->cfg.D_gb; q.D_precip = c->cfg.D_precip;
    q.decay = c->cfg.corrosion_decay_l > 0.0 ? std::pow(10.0, -c->volume_loss / c->cfg.corrosion_decay_l) : 1.0;
    q.alpha_dx = c->cfg.alpha_art_diff * c->cfg.dx;
    q.beta = k.beta_lap; q.div_coeff = k.alpha / k.V_H; q.C_sat = c->cfg.C_sat;
    if (!st->attr_set_ard || st->attr_smem_ard < st->smem) {
        CUDA_OK(cudaFuncSetAttribute(k_ard2d_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem));
        st->attr_set_ard = true; st->attr_smem_ard = st->smem;
    }
    CUDA_OK(cudaMemsetAsync(st->bar, 0, sizeof(unsigned), c->stream));
    void* args[] = {(void*)&q};
    const unsigned grid = (unsigned)(st->n_step + (st->KP > 0 ? 1 : 0));
    CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_ard2d_loop, dim3(grid), dim3(NT2D), args, st->smem, c->stream));
    c->launches++;
    return 0;
}
