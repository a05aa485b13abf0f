// stream.cuh -- machinery of the z-streaming NS bond kernel (ns_stream.cu):
// 3D, m_ratio = 3 (reach 3), full FLUID rows.
//
// A persistent CTA owns a 16 x 8 column of lattice nodes and STREAMS along z through a chunk of
// planes in steps of 4 planes.  The haloed planes ((16+6) x (8+6) values per field) live in a
// ring of 14 plane slots in shared memory: 10 planes are read by the current step while the 4
// planes of the next step arrive.  Planes are moved by 16-byte cp.async (LDGSTS.128 in SASS; every
// thread requests its share, ~6.6 copies per step, in the middle of its bond loop) with completion
// counted on an mbarrier (cp.async.mbarrier.arrive.noinc), so the transfer of step s+1 overlaps
// the FP64 work of step s.  (Bulk row copies -- cp.async.bulk / UBLKCP -- were measured and lost:
// they issue through uniform registers, ~50 cycles per 192-byte row; profiles/r2_notes.md.)  Every staged plane is read from
// L2 once per chunk: 2.4x read amplification (in-plane halo only) instead of 4.2x for the
// block-per-CTA kernels (tile.cuh).
//
// Row alignment.  16-byte copies need 16-byte aligned addresses; with an odd lattice pitch (157
// at params_fine) the first element of a staged row is 16-byte aligned only for every other
// (row, plane).  Each row copy therefore starts at the even element at or below the row start and
// moves 24 doubles; the row's data begin at element `par` = (element index & 1) of the slot row.
// par alternates like a checkerboard in (row, plane) when the pitches are odd, so the bond loop
// keeps two column bases (even / odd window plane) and all offsets stay immediates.
// Rows that stick out of the lattice box read the neighbouring row / plane (or the zeroed pad in
// front of / behind the arrays, pd_alloc_fields): such values only reach sums of non-FLUID lanes,
// which are discarded (full rows: a FLUID row never leaves the box).
//
// 16 compute warps per CTA: the 256 node-owning threads (16 x 8 x 2 layers, 2 z-nodes each)
// exist twice; group 0 walks one half of the 37 (di,dj) columns of the horizon sphere, group 1
// the other half, and the halves are combined through shared memory (group g finalises z-node g
// of every thread pair).
#pragma once
#include "tile.cuh"

namespace stream {

using tile::TR;
using tile::TX;
using tile::TY;
using tile::RZ;
using tile::NCOL;

constexpr int MS = 4;                       // planes per step
constexpr int MLAYERS = MS / RZ;            // thread layers in z
constexpr int MROWS = TY + 2 * TR;          // 14 staged rows
constexpr int MPITCH = 24;                  // doubles per staged row (22 + alignment slack; 192 B)
constexpr int MFS = MROWS * MPITCH;         // doubles per field per plane
constexpr int MRING = 14;                   // plane slots
constexpr int MWIN = MS + 2 * TR;           // planes a step reads
constexpr int MGROUP = TX * TY * MLAYERS;   // threads per column group
constexpr int MTHREADS = 2 * MGROUP;
constexpr int ROWBYTES = MPITCH * 8;

// column table ordered by group (columns [beg[g], end[g]) belong to group g), half height
// descending inside a group
struct StreamCols {
    int off[NCOL];        // dj * MPITCH + di
    int djodd[NCOL];      // dj & 1
    int h[NCOL];
    int beg[2], mid[2], end[2];   // [beg, mid): half height 3, [mid, end): the rest
    double di[NCOL], dj[NCOL];
    double kap[NCOL][4], kz[NCOL][4], aux[NCOL][4];
};

// per-context state of the streaming kernels
struct TileState {
    long long epoch = -1;       // tables_epoch the tile list was built for
    int* d_tiles = nullptr;     // active (bx | by << 16) tiles: columns with a non-OUTSIDE node
    int ntiles = 0;
    int* d_work = nullptr;      // work-item counters of the persistent kernels (one per launch in flight)
    unsigned work_seq = 0;
    int sm_count = 0;
    bool attr_ns = false, attr_ard = false, attr_tile_ns = false, attr_tile_ard = false;
    bool cols_ok = false;
    double sum_kappa = 0.0;
    StreamCols cols;
    tile::ColTable tcols;       // block-per-CTA kernels (tile.cuh)
};

inline bool build_stream_cols(const tile::ColTable& T0, StreamCols* S) {
    int n = 0;
    for (int g = 0; g < 2; ++g) {
        S->beg[g] = n;
        for (int pass = 3; pass >= 1; --pass) {
            if (pass == 2) S->mid[g] = n;
            for (int c = 0; c < NCOL; ++c) {
                if (T0.h[c] != pass) continue;
                const int di = T0.di_i[c], dj = T0.dj_i[c];
                int grp = (dj > 0 || (dj == 0 && di >= 0)) ? 0 : 1;
                if (di == 3 && dj == 0) grp = 1;          // balance: 180 / 178 bond slots per thread
                if (grp != g) continue;
                S->off[n] = dj * MPITCH + di;
                S->djodd[n] = dj & 1;
                S->h[n] = pass;
                S->di[n] = di;
                S->dj[n] = dj;
                for (int k = 0; k < 4; ++k) {
                    S->kap[n][k] = T0.kap[c][k];
                    S->kz[n][k] = T0.kz[c][k];
                    S->aux[n][k] = T0.aux[c][k];
                }
                ++n;
            }
        }
        S->end[g] = n;
    }
    return n == NCOL;
}

// lattice, plane range and work items of one launch of a streaming kernel
struct StreamGeom {
    int Nx, Ny;
    long long P;
    int zb, ze;            // local plane range of this launch
    int zc;                // planes per work item
    int nchunks, ntiles;   // work items = ntiles * nchunks, chunk-major
    int par_p, par_x;      // P & 1, Nx & 1 (row alignment parity)
    const int* tiles;      // (x0 / TX) | (y0 / TY) << 16 of the active tiles
    int* work;             // work-item counter of this launch (zeroed before the launch)
};

// one work item: tile column (x0, y0) x planes [z0, z0 + len)
struct StreamItem {
    int x0, y0, z0, len, nsteps, np;
    long long ebase;       // element index of staged (row 0, plane 0) = (x0-3, y0-3, z0-3)
};

// One 16-byte piece of a staged row that a thread copies for every plane of the item.
struct CopyDesc {
    const double* src;   // field + first element of the piece in staged plane 0 (not yet aligned down)
    int dst;             // offset inside a plane slot; < 0: no piece
    int par;             // (element index of the row start in plane 0) & 1
};

#ifdef __CUDACC__
__device__ __forceinline__ StreamItem stream_item(const StreamGeom& g, int item) {
    StreamItem it;
    const int ch = item / g.ntiles;
    const int tl = g.tiles[item - ch * g.ntiles];
    it.x0 = (tl & 0xffff) * TX;
    it.y0 = (tl >> 16) * TY;
    it.z0 = g.zb + ch * g.zc;
    it.len = min(g.zc, g.ze - it.z0);
    it.nsteps = (it.len + MS - 1) / MS;
    it.np = it.len + 2 * TR;       // staged planes z0-3 .. z0+len+2
    it.ebase = (long long)(it.z0 - TR) * g.P + (long long)(it.y0 - TR) * g.Nx + (it.x0 - TR);
    return it;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Spin on the phase with the given parity.  A protocol error would otherwise hang the GPU: after
// ~2^28 failed polls (tens of seconds) the kernel traps and the launch reports an error instead.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned done = 0;
    for (unsigned spins = 0; !done; ++spins) {
        // try_wait suspends the thread in hardware for up to the hinted time (ns) before it reports
        // failure: a waiting warp leaves the issue slots to the warps that feed the FP64 pipe
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(2000u)
            : "memory");
        if (!done && spins > (1u << 24)) asm volatile("trap;\n");
    }
}
// one non-blocking poll of the phase with the given parity
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier `id` (1..15) over `nthreads` threads (whole warps)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
// 16-byte asynchronous copy global -> shared (LDGSTS.128, L2 only), per-thread commit groups
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// arrival on `bar` (not counted as an additional pending arrival) once all earlier cp.async of this thread are done
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// bulk copy global -> shared (UBLKCP); bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif

}  // namespace stream

stream::TileState* pd_tile_state(pdgpu_ctx* c);    // ns_stream.cu: created on first use
int pd_stream_prepare(pdgpu_ctx* c);               // ns_stream.cu: column tables + active tile list (cached per tables_epoch)
void pd_tile_state_free(pdgpu_ctx* c);
