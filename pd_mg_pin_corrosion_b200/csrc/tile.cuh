// tile.cuh -- shared geometry of the tiled bond kernels (ns_tile.cu, ard_tile.cu):
// 3D, m_ratio = 3 (reach 3), full FLUID rows.
//
// A CTA (16 x 8 x 4 threads) stages a haloed (16+6) x (8+6) x (8+6) block in shared memory;
// thread (tx,ty,tz) owns the 2 nodes (x0+tx, y0+ty, z0+2tz..z0+2tz+1): 16 warps per SM hide the
// FP64 and shared-memory latencies (8 warps with 4 nodes per thread reached 48 % FP64-pipe
// utilisation, profiles/r1_notes.md).  A warp is two x-rows of 16 nodes: on the circular tube
// cross-section 88 % of the lanes of a working warp own a FLUID node (82 % with 32-wide rows), and
// 8 planes per tile amortise the 6 halo planes (measured: 32x8x4 tiles 3.24 ms, 16x16x4 2.97 ms,
// 16x8x8 2.92 ms for the NS kernel).  The horizon sphere (di^2+dj^2+dk^2 <= 12, 178
// offsets) is walked as 37 (di,dj) COLUMNS in a runtime loop; inside a column the window
// slides along z, so a staged neighbour value is read from shared memory once and used for
// up to 2 bonds.  Only the half-height H of the column (1, 2 or 3) is a compile-time
// parameter: three unrolled bodies of 6/10/14 bonds keep the instruction footprint at
// ~20 KB (a fully unrolled 712-bond body is 230 KB and stalls on instruction fetch: ncu
// "no_instruction" 3.2 per issue, profiles/r1_notes.md).
//
// Bond weights: e w1 = (d dx / r)(V/r) = d (dx V / r^2) = d * kappa with
// kappa = dx * w2, so ONE weight per (column, |dk|) serves the gradient sums and the
// Laplacians (which are accumulated with kappa and rescaled by 1/dx at the end).
#pragma once
#include "common.cuh"

namespace tile {

constexpr int TR = 3;
constexpr int TX = 16, TY = 8;             // threads in x, y
constexpr int RZ = 2;                      // z-nodes per thread (sliding window length)
#ifndef PD_TILE_NZT
#define PD_TILE_NZT 4
#endif
constexpr int NZT = PD_TILE_NZT;           // thread layers in z (6 = 768 threads / 12 planes measured slower for NS: 78 registers)
constexpr int TZ = RZ * NZT;               // z-nodes per tile
constexpr int SX = TX + 2 * TR, SY = TY + 2 * TR, SZ = TZ + 2 * TR;
constexpr int SPLANE = SX * SY, SN = SPLANE * SZ;
constexpr int NTHREADS = TX * TY * NZT;
constexpr int NCOL = 37;

struct ColTable {
    int off[NCOL];        // dj*SX + di (ns_tile / ard_tile block pitch)
    int di_i[NCOL], dj_i[NCOL];
    int h[NCOL];          // half height of the column: |dk| <= h
    double di[NCOL], dj[NCOL];
    double kap[NCOL][4];  // kappa(|dk|), 0 where the offset does not exist (incl. the node itself)
    double kz[NCOL][4];   // |dk| * kappa(|dk|)
    double aux[NCOL][4];  // kernel specific (NS: mu*beta/dx * kappa); kept in the table so that the
                          // weights reach the DFMAs as uniform-register operands (two register
                          // sources per DFMA instead of three, profiles/r1_notes.md)
};

struct TileGeom {
    int Nx, Ny, nlp, z_lo, z_hi;   // in-plane extents, local planes, owned local plane range
    long long P;
};

// Host: build the column table from the context's stencil; false if the stencil is not the
// m=3 sphere.
inline bool build_columns(const pdgpu_ctx* c, ColTable* T, double* sum_kappa) {
    if (c->dim != 3 || c->cfg.m_ratio != 3 || c->n_off != 178) return false;
    int n = 0;
    double sk = 0.0;
    for (int pass = 3; pass >= 1; --pass)          // columns ordered by H: uniform branch pattern
        for (int dj = -3; dj <= 3; ++dj)
            for (int di = -3; di <= 3; ++di) {
                int r2 = 12 - di * di - dj * dj;
                if (r2 < 0) continue;
                int H = r2 >= 9 ? 3 : r2 >= 4 ? 2 : r2 >= 1 ? 1 : 0;
                if (H != pass) continue;
                if (n >= NCOL) return false;
                T->off[n] = dj * SX + di;
                T->di_i[n] = di;
                T->dj_i[n] = dj;
                T->h[n] = H;
                T->di[n] = di;
                T->dj[n] = dj;
                for (int k = 0; k < 4; ++k) T->kap[n][k] = T->kz[n][k] = T->aux[n][k] = 0.0;
                ++n;
            }
    if (n != NCOL) return false;
    for (const OffEntry& e : c->h_off) {
        if (e.di * e.di + e.dj * e.dj + e.dk * e.dk > 12) return false;
        int col = -1;
        for (int q = 0; q < NCOL; ++q)
            if ((int)T->di[q] == e.di && (int)T->dj[q] == e.dj) col = q;
        int ak = e.dk < 0 ? -e.dk : e.dk;
        if (col < 0 || ak > T->h[col]) return false;
        double kappa = c->cfg.dx * e.w2;
        T->kap[col][ak] = kappa;
        T->kz[col][ak] = ak * kappa;
        sk += kappa;
    }
    *sum_kappa = sk;
    return true;
}

inline TileGeom make_geom(const pdgpu_ctx* c) {
    TileGeom g;
    g.Nx = c->Nx; g.Ny = c->Ny; g.nlp = c->nlp; g.z_lo = c->R; g.z_hi = c->R + (c->a1 - c->a0);
    g.P = c->P;
    return g;
}

#ifdef __CUDACC__
// Asynchronous global -> shared copy of one double (LDGSTS); src_bytes = 0 zero-fills the
// destination (elements outside the box). No registers are held while the copy is in flight,
// so a thread keeps all of its ~52 staging copies outstanding at once.
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int src_bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// local index of staged element idx of the block at (x0,y0,z0), or -1 outside the box
__device__ __forceinline__ long long staged_index(const TileGeom& g, int idx, int x0, int y0, int z0) {
    const int sz = idx / SPLANE;
    const int rem = idx - sz * SPLANE;
    const int sy = rem / SX;
    const int sx = rem - sy * SX;
    const int ax = x0 - TR + sx, ay = y0 - TR + sy, az = z0 - TR + sz;
    if (ax < 0 || ax >= g.Nx || ay < 0 || ay >= g.Ny || az >= g.nlp) return -1;
    return (long long)az * g.P + (long long)ay * g.Nx + ax;
}
#endif

}  // namespace tile
