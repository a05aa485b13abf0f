// host_step.cu -- one coupling-loop pass (NS loop body + ARD loop body) on HOST-resident state.
//
// The reference keeps Fields in host memory (src/fields.h:28-58) and its loop bodies
// (src/pd_ns.cpp:196-205, src/coupling.cpp:232-240) read and write those vectors.  A drop-in call
// with host arrays therefore moves rho, vel and C over PCIe in both directions every step, and
// those copies -- not the bond kernels -- bound the step (5 doubles per node each way).
//
// pdgpu_step_host hides the compute and one of the two copy directions: the axial planes are cut
// into chunks that are processed from the outlet end downwards,
//
//     upload(k)        copy engine 1   rho, vel (AoS), C of chunk k
//     A(k)             compute         EOS, inlet/outlet/wall/solid BCs on chunk k
//     N(k+1), W(k+1)   compute         NS bond kernel and wall mirror of the new buffers
//     B(k+1)           compute         ARD-body BCs, |v|, wall concentration, salt pre-pass
//     D(k+2)           compute         ARD bond kernel
//     download(k+2)    copy engine 2   new rho, vel, C of chunk k+2
//
// so that every bond kernel sees its +-reach planes in the state the sequential order gives
// them.  The top chunk T is the thin outlet region: its chain (outlet sweep -> N(T) -> outlet sweep
// -> D(T), D(T-1)) is 4-5 ms of dependent single-CTA work and runs on the high-priority side
// stream with its own download stream, while the main stream streams the chunks below it.  The arithmetic per node is that of pdgpu_ns_iterate(1) + pdgpu_ard_iterate(1); results
// are bit-identical (tests/test_gpu_parity.py::test_step_host_*).  Geometries that do not meet
// the chunk invariants (checked when the plan is built) run the same operators unchunked.
//
// Slab contexts (one rank per GPU): every rank pipelines its own slab the same way; the ghost planes
// of the CURRENT state come straight from the host arrays, so the only exchanges left are the ones
// of the classical loop bodies, issued once all NS chunks are done: new flow fields -> |v| of the
// ghost planes -> salt/dsol/wpack of ghost solids; the ARD kernels of the two chunks that touch a
// neighbour's planes (and their downloads) run after them.
#include <algorithm>

#include "common.cuh"

struct HostStep {
    long long epoch = -1;
    int n_req = 0;
    int n = 1;                          // chunks (1 = unchunked)
    std::vector<int> zb;                // local plane boundaries [n+1]
    std::vector<long long> wall_lo, solid_lo, ssolid_lo;   // list offsets per chunk [n+1]
    long long gwall_split = 0;          // ghost-plane WALL entries below / above the owned planes
    long long wall_split = 0;           // first WALL entry in an outlet plane (entries of the top chunk
                                        // before it do not depend on the outlet sweep)
    cudaStream_t s_up = nullptr, s_down = nullptr, s_down2 = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_a = nullptr, ev_x = nullptr, ev_end2 = nullptr, ev_sb = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_cmp, ev_down;   // per chunk: upload done, kernels done, download done
    cudaEvent_t ev_start = nullptr, ev_end = nullptr;
    double *stage_up = nullptr, *stage_down = nullptr;
    size_t stage_elems = 0;
    char why[160] = {0};                // why the plan fell back to one chunk
};

void pd_host_step_free(pdgpu_ctx* c) {
    HostStep* h = c->hs;
    if (!h) return;
    for (cudaEvent_t e : h->ev_up) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_cmp) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_down) cudaEventDestroy(e);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_end) cudaEventDestroy(h->ev_end);
    if (h->s_up) cudaStreamDestroy(h->s_up);
    if (h->s_down) cudaStreamDestroy(h->s_down);
    if (h->s_down2) cudaStreamDestroy(h->s_down2);
    for (cudaEvent_t e : {h->ev_t0, h->ev_a, h->ev_x, h->ev_end2, h->ev_sb})
        if (e) cudaEventDestroy(e);
    if (h->stage_up) cudaFree(h->stage_up);
    if (h->stage_down) cudaFree(h->stage_down);
    delete h;
    c->hs = nullptr;
}

namespace {

// narrows the WALL and SOLID lists of the context to the entries of one chunk
struct ListWindow {
    pdgpu_ctx* c;
    int *w, *wm, *s, *ss;
    long long nw, ns, nss;
    ListWindow(pdgpu_ctx* ctx, long long w0, long long w1, long long s0, long long s1, long long q0 = 0, long long q1 = 0)
        : c(ctx), w(ctx->l_wall), wm(ctx->l_wall_mirror), s(ctx->l_solid), ss(ctx->l_ssolid), nw(ctx->n_wall),
          ns(ctx->n_solid), nss(ctx->n_ssolid) {
        c->l_ssolid = ss + q0;
        c->n_ssolid = q1 - q0;
        c->l_wall = w + w0;
        c->l_wall_mirror = wm + w0;
        c->n_wall = w1 - w0;
        c->l_solid = s + s0;
        c->n_solid = s1 - s0;
    }
    ListWindow(pdgpu_ctx* ctx, const HostStep& h, int k)
        : ListWindow(ctx, h.wall_lo[k], h.wall_lo[k + 1], h.solid_lo[k], h.solid_lo[k + 1], h.ssolid_lo[k],
                     h.ssolid_lo[k + 1]) {}
    ~ListWindow() {
        c->l_wall = w; c->l_wall_mirror = wm; c->n_wall = nw;
        c->l_solid = s; c->n_solid = ns;
        c->l_ssolid = ss; c->n_ssolid = nss;
    }
};

// narrows the ghost-plane WALL list (slab contexts)
struct GhostWallWindow {
    pdgpu_ctx* c;
    int *w, *wm;
    long long n;
    GhostWallWindow(pdgpu_ctx* ctx, long long first, long long end)
        : c(ctx), w(ctx->l_gwall), wm(ctx->l_gwall_mirror), n(ctx->n_gwall) {
        c->l_gwall = w + first;
        c->l_gwall_mirror = wm + first;
        c->n_gwall = end - first;
    }
    ~GhostWallWindow() { c->l_gwall = w; c->l_gwall_mirror = wm; c->n_gwall = n; }
};

int fallback(HostStep* h, const char* why) {
    h->n = 1;
    snprintf(h->why, sizeof(h->why), "%s", why);
    return 0;
}

int build_plan(pdgpu_ctx* c, HostStep* h, int n_req) {
    h->epoch = c->tables_epoch;
    h->n_req = n_req;
    h->why[0] = 0;
    const int lo = c->R, nz = c->a1 - c->a0, hi = lo + nz;
    const int TZ = 8;   // tile::TZ: chunk boundaries stay tile aligned
    h->zb.assign(2, lo);
    h->zb[1] = hi;
    h->n = 1;
    if (n_req < 2) return fallback(h, "one chunk requested");
    if (c->nranks > 1 && !c->comm) return fallback(h, "slab context without communicator");
    if (c->opt_ns_kernel == 3 || c->opt_ard_kernel == 3) return fallback(h, "CSR kernels take no plane range");
    if (c->cfg.channel_flow_corrections) return fallback(h, "channel_flow_corrections");
    if (c->n_outlet > 0 && !(c->out_fast && c->opt_outlet_kernel > 0)) {
        // the level-list sweep is fine too, it only needs the outlet planes; nothing to check
    }
    // top chunk: every OUTLET node and its reach must lie inside it; with an outlet it is just that
    // region (its kernels wait for the sequential outlet sweeps)
    int thick = ((nz + n_req - 1) / n_req + TZ - 1) / TZ * TZ;
    thick = std::max(thick, 2 * c->R + TZ);
    int zt = hi - thick;
    int zo_first = hi, zi_last = -1;
    if (c->n_outlet) {
        int first = 0;
        CUDA_OK(cudaMemcpy(&first, c->l_outlet, sizeof(int), cudaMemcpyDeviceToHost));
        zo_first = (int)(first / c->P);
        zt = zo_first - c->R;
    }
    if (c->n_inlet) {
        int last = 0;
        CUDA_OK(cudaMemcpy(&last, c->l_inlet + (c->n_inlet - 1), sizeof(int), cudaMemcpyDeviceToHost));
        zi_last = (int)(last / c->P);
    }
    zt = lo + (zt - lo) / TZ * TZ;
    if (zt - lo < 2 * c->R + TZ || hi - zt <= c->R) return fallback(h, "domain too short for two chunks");
    // lower chunks, tile aligned: graded thickness -- thin at both ends (the first download can start
    // after two thin chunks, the drain after the last upload is two thin chunks), thick in the middle
    // (fewer, larger copies and kernel launches). Weights 1,1,2,2,3,4,4,... from either end.
    int n_low = std::max(1, std::min(n_req - 1, (zt - lo) / std::max(2 * c->R + TZ, TZ)));
    std::vector<int> zb;
    zb.push_back(lo);
    {
        static const int ramp[] = {1, 1, 2, 2, 3};
        std::vector<double> wgt(n_low);
        double tot = 0.0;
        for (int k = 0; k < n_low; ++k) {
            const int e = std::min(k, n_low - 1 - k);   // distance from the nearer end
            wgt[k] = (c->opt_host_step_graded && n_low >= 8) ? (e < 5 ? ramp[e] : 4) : 1;
            tot += wgt[k];
        }
        double acc = 0.0;
        for (int k = 1; k < n_low; ++k) {
            acc += wgt[k - 1];
            int z = lo + (int)((double)(zt - lo) * acc / tot) / TZ * TZ;
            if (z - zb.back() >= 2 * c->R + TZ && zt - z >= 2 * c->R + TZ) zb.push_back(z);
        }
    }
    zb.push_back(zt);
    zb.push_back(hi);
    const int n = (int)zb.size() - 1;
    if (zi_last >= zb[1]) return fallback(h, "INLET nodes above the first chunk");

    std::vector<int> w(c->n_wall), wm(c->n_wall), s(c->n_solid);
    if (c->n_wall) {
        CUDA_OK(cudaMemcpy(w.data(), c->l_wall, sizeof(int) * c->n_wall, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(wm.data(), c->l_wall_mirror, sizeof(int) * c->n_wall, cudaMemcpyDeviceToHost));
    }
    if (c->n_solid) CUDA_OK(cudaMemcpy(s.data(), c->l_solid, sizeof(int) * c->n_solid, cudaMemcpyDeviceToHost));
    std::vector<int> q(c->n_ssolid);
    if (c->n_ssolid) CUDA_OK(cudaMemcpy(q.data(), c->l_ssolid, sizeof(int) * c->n_ssolid, cudaMemcpyDeviceToHost));
    std::vector<long long> wl(n + 1), sl(n + 1), ql(n + 1);
    for (int k = 0; k <= n; ++k) {
        long long first = (long long)zb[k] * c->P;
        wl[k] = std::lower_bound(w.begin(), w.end(), first, [](int a, long long b) { return (long long)a < b; }) - w.begin();
        sl[k] = std::lower_bound(s.begin(), s.end(), first, [](int a, long long b) { return (long long)a < b; }) - s.begin();
        ql[k] = std::lower_bound(q.begin(), q.end(), first, [](int a, long long b) { return (long long)a < b; }) - q.begin();
    }
    wl[0] = 0; sl[0] = 0; ql[0] = 0; wl[n] = c->n_wall; sl[n] = c->n_solid; ql[n] = c->n_ssolid;
    // a WALL node must find its mirror node inside its own chunk (same BC state as unchunked)
    for (int k = 0; k < n; ++k)
        for (long long t = wl[k]; t < wl[k + 1]; ++t) {
            if (wm[t] < 0) continue;
            int zm = (int)(wm[t] / c->P);
            if (zm < zb[k] || zm >= zb[k + 1]) return fallback(h, "a wall mirror crosses a chunk boundary");
        }
    // WALL entries of the top chunk below the first outlet plane are mirrored before the outlet sweep
    // has run: they must not mirror an OUTLET node
    long long split = std::lower_bound(w.begin(), w.end(), (long long)zo_first * c->P,
                                       [](int a, long long b) { return (long long)a < b; }) - w.begin();
    split = std::max(split, wl[n - 1]);
    if (c->n_outlet) {
        const long long t0 = (long long)zb[n - 1] * c->P, tn = (long long)(hi - zb[n - 1]) * c->P;
        std::vector<uint8_t> ty(tn);
        CUDA_OK(cudaMemcpy(ty.data(), c->type + t0, (size_t)tn, cudaMemcpyDeviceToHost));
        for (long long t = wl[n - 1]; t < split; ++t)
            if (wm[t] >= 0 && ty[wm[t] - t0] == PDGPU_OUTLET)
                return fallback(h, "a wall below the outlet planes mirrors an OUTLET node");
        // ... and the walls of the outlet planes are mirrored next to the solid no-slip BC of the chunk
        for (long long t = split; t < wl[n]; ++t)
            if (wm[t] >= 0 && ty[wm[t] - t0] == PDGPU_SOLID_MG)
                return fallback(h, "a wall of the outlet planes mirrors a SOLID_MG node");
    }
    h->gwall_split = 0;
    if (c->n_gwall) {
        std::vector<int> gw(c->n_gwall);
        CUDA_OK(cudaMemcpy(gw.data(), c->l_gwall, sizeof(int) * c->n_gwall, cudaMemcpyDeviceToHost));
        h->gwall_split = std::lower_bound(gw.begin(), gw.end(), c->own_lo,
                                          [](int a, long long b) { return (long long)a < b; }) - gw.begin();
    }
    h->n = n;
    h->zb = zb;
    h->wall_lo = wl;
    h->solid_lo = sl;
    h->ssolid_lo = ql;
    h->wall_split = split;
    return 0;
}

int ensure(pdgpu_ctx* c, int n_req) {
    if (!c->hs) c->hs = new HostStep();
    HostStep* h = c->hs;
    if (!h->s_up) {
        CUDA_OK(cudaStreamCreateWithFlags(&h->s_up, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&h->s_down, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&h->s_down2, cudaStreamNonBlocking));
        for (cudaEvent_t* e : {&h->ev_t0, &h->ev_a, &h->ev_x, &h->ev_end2, &h->ev_sb})
            CUDA_OK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreate(&h->ev_start));
        CUDA_OK(cudaEventCreate(&h->ev_end));
    }
    if (h->epoch != c->tables_epoch || h->n_req != n_req) PD_TRY(build_plan(c, h, n_req));
    while ((int)h->ev_up.size() < h->n) {
        cudaEvent_t a, b, d;   // timing enabled: pdgpu_step_host_trace reads them
        CUDA_OK(cudaEventCreate(&a));
        CUDA_OK(cudaEventCreate(&b));
        CUDA_OK(cudaEventCreate(&d));
        h->ev_up.push_back(a);
        h->ev_cmp.push_back(b);
        h->ev_down.push_back(d);
    }
    size_t need = (size_t)c->NL * c->dim;   // indexed by local node (ghost planes included)
    if (h->n > 1 && h->stage_elems < need) {
        if (h->stage_up) CUDA_OK(cudaFree(h->stage_up));
        if (h->stage_down) CUDA_OK(cudaFree(h->stage_down));
        h->stage_up = h->stage_down = nullptr;
        CUDA_OK(cudaMalloc(&h->stage_up, sizeof(double) * need));
        CUDA_OK(cudaMalloc(&h->stage_down, sizeof(double) * need));
        h->stage_elems = need;
    }
    return 0;
}

}  // namespace

extern "C" int pdgpu_step_host_chunks(pdgpu_ctx* c, int n_chunks, int* n_used, char* why, int why_len) {
    NEED_GRID(c);
    PD_TRY(ensure(c, n_chunks));
    if (n_used) *n_used = c->hs->n;
    if (why && why_len > 0) snprintf(why, (size_t)why_len, "%s", c->hs->why);
    return 0;
}

extern "C" int pdgpu_step_host(pdgpu_ctx* c, double dt_ns, double dt_ard, double* rho, double* vel, double* C,
                               int n_chunks) {
    NEED_GRID(c);
    if (!rho || !vel || !C) PD_FAIL("pdgpu_step_host: null host array");
    PD_TRY(ensure(c, n_chunks));
    HostStep* h = c->hs;
    if (h->n <= 1) {   // same operators, whole domain at once
        PD_TRY(pdgpu_fields_upload(c, PDGPU_F_RHO, rho));
        PD_TRY(pdgpu_fields_upload(c, PDGPU_F_VEL, vel));
        PD_TRY(pdgpu_fields_upload(c, PDGPU_F_C, C));
        PD_TRY(pdgpu_ns_iterate(c, 1, dt_ns));
        PD_TRY(pdgpu_ard_iterate(c, 1, dt_ard));
        PD_TRY(pdgpu_fields_download(c, PDGPU_F_RHO, rho));
        PD_TRY(pdgpu_fields_download(c, PDGPU_F_VEL, vel));
        PD_TRY(pdgpu_fields_download(c, PDGPU_F_C, C));
        return 0;
    }
    c->wallC_pending = false;   // every WALL concentration is rewritten below
    c->fields_ready = true;
    const int n = h->n, dim = c->dim;
    const int cur = c->cur, nw = 1 - c->cur, sC = c->curC, dC = 1 - c->curC;
    const long long P = c->P;
    const long long gshift = (long long)(c->a0 - c->R) * P;   // global index = local + gshift
    cudaStream_t cs = c->stream;
    PD_TRY(pd_set_dt(c, 0, dt_ns));
    PD_TRY(pd_set_dt(c, 1, dt_ard));
    CUDA_OK(cudaEventRecord(h->ev_start, cs));
    CUDA_OK(cudaStreamWaitEvent(h->s_up, h->ev_start, 0));
    CUDA_OK(cudaStreamWaitEvent(h->s_down, h->ev_start, 0));

    const int T = n - 1;
    const bool has_lo = c->nranks > 1 && c->rank > 0, has_hi = c->nranks > 1 && c->rank < c->nranks - 1;
    // plane range a chunk moves/initialises: the end chunks of a slab take the ghost planes along
    auto span_x = [&](int k, long long* l0, long long* cnt) {
        int z0 = h->zb[k], z1 = h->zb[k + 1];
        if (k == 0 && has_lo) z0 -= c->R;
        if (k == T && has_hi) z1 += c->R;
        *l0 = (long long)z0 * P;
        *cnt = (long long)(z1 - z0) * P;
    };
    // uploads, outlet end first
    for (int k = n - 1; k >= 0; --k) {
        long long l0, cnt;
        span_x(k, &l0, &cnt);
        const long long g0 = l0 + gshift, s0 = l0 * dim;
        CUDA_OK(cudaMemcpyAsync(c->rho[cur] + l0, rho + g0, sizeof(double) * cnt, cudaMemcpyHostToDevice, h->s_up));
        CUDA_OK(cudaMemcpyAsync(h->stage_up + s0, vel + g0 * dim, sizeof(double) * cnt * dim, cudaMemcpyHostToDevice,
                                h->s_up));
        CUDA_OK(cudaMemcpyAsync(c->C[sC] + l0, C + g0, sizeof(double) * cnt, cudaMemcpyHostToDevice, h->s_up));
        CUDA_OK(cudaEventRecord(h->ev_up[k], h->s_up));
    }

    const bool use_side = c->n_outlet > 0;   // the outlet chain lives on the rank that owns the outlet
    cudaStream_t side = c->stream2;
    CUDA_OK(cudaStreamWaitEvent(h->s_down2, h->ev_start, 0));
    const bool tiles = (c->opt_ard_kernel == 1 || c->opt_ard_kernel == 2);
    auto span = [&](int k, long long* l0, long long* cnt) {
        *l0 = (long long)h->zb[k] * P;
        *cnt = (long long)(h->zb[k + 1] - h->zb[k]) * P;
    };
    // D(kd) + download of the finished chunk on (stream, download stream)
    auto finish_chunk = [&](int kd, cudaStream_t st, cudaStream_t sd) -> int {
        StreamSwap sw(c, st);
        {
            ListWindow win(c, *h, kd);   // src/coupling.cpp:236-238
            PD_TRY(pd_enqueue_ard_main(c, nw, sC, c->d_dt + 1, h->zb[kd], h->zb[kd + 1], true, tiles));
        }
        long long l0, cnt;
        span(kd, &l0, &cnt);
        const long long g0 = l0 + gshift, s0 = l0 * dim;
        PD_TRY(pd_enqueue_interleave(c, h->stage_down + s0, l0, cnt, nw));
        CUDA_OK(cudaEventRecord(h->ev_cmp[kd], st));
        CUDA_OK(cudaStreamWaitEvent(sd, h->ev_cmp[kd], 0));
        CUDA_OK(cudaMemcpyAsync(rho + g0, c->rho[nw] + l0, sizeof(double) * cnt, cudaMemcpyDeviceToHost, sd));
        CUDA_OK(cudaMemcpyAsync(vel + g0 * dim, h->stage_down + s0, sizeof(double) * cnt * dim,
                                cudaMemcpyDeviceToHost, sd));
        CUDA_OK(cudaMemcpyAsync(C + g0, c->C[dC] + l0, sizeof(double) * cnt, cudaMemcpyDeviceToHost, sd));
        CUDA_OK(cudaEventRecord(h->ev_down[kd], sd));
        return 0;
    };

    for (int k = n - 1; k >= -2; --k) {
        if (k >= 0) {   // A(k): src/pd_ns.cpp:197-200 restricted to chunk k
            long long l0, cnt;
            span_x(k, &l0, &cnt);
            CUDA_OK(cudaStreamWaitEvent(cs, h->ev_up[k], 0));
            PD_TRY(pd_enqueue_eos_range(c, cur, l0, cnt));
            PD_TRY(pd_enqueue_deinterleave(c, h->stage_up + l0 * dim, l0, cnt, cur));
            if (k == T && use_side) {
                // outlet sweep and the walls of the outlet planes: side stream; the walls below them
                // (all that the chunk underneath reads) and the solids: main stream
                CUDA_OK(cudaEventRecord(h->ev_t0, cs));
                CUDA_OK(cudaStreamWaitEvent(side, h->ev_t0, 0));
                {
                    StreamSwap sw(c, side);
                    ListWindow win(c, h->wall_split, h->wall_lo[T + 1], h->solid_lo[T], h->solid_lo[T]);
                    PD_TRY(pd_enqueue_bc_outlet(c, cur, sC));
                    PD_TRY(pd_enqueue_bc_wall(c, cur));
                }
                ListWindow win(c, h->wall_lo[T], h->wall_split, h->solid_lo[T], h->solid_lo[T + 1]);
                PD_TRY(pd_enqueue_bc_wall(c, cur));
                PD_TRY(pd_enqueue_bc_solid(c, cur));
            } else {
                ListWindow win(c, *h, k);
                if (k == 0) PD_TRY(pd_enqueue_bc_inlet(c, cur, sC));
                if (k == T) PD_TRY(pd_enqueue_bc_outlet(c, cur, sC));   // (no fast sweep on this geometry)
                PD_TRY(pd_enqueue_bc_wall(c, cur));
                if (k == 0 && has_lo) { GhostWallWindow gw(c, 0, h->gwall_split); PD_TRY(pd_enqueue_bc_wall(c, cur, 3)); }
                if (k == T && has_hi) { GhostWallWindow gw(c, h->gwall_split, c->n_gwall); PD_TRY(pd_enqueue_bc_wall(c, cur, 3)); }
                PD_TRY(pd_enqueue_bc_solid(c, cur));
            }
        }
        const int kn = k + 1;
        if (kn >= 0 && kn < n) {   // N, W: src/pd_ns.cpp:201-204; B: src/coupling.cpp:232-235 + ARD pre-passes
            cudaStream_t st = cs;
            if (kn == T && use_side) {   // after A(T-1) on the main stream, behind the sweep on the side stream
                CUDA_OK(cudaEventRecord(h->ev_a, cs));
                CUDA_OK(cudaStreamWaitEvent(side, h->ev_a, 0));
                st = side;
            }
            StreamSwap sw(c, st);
            PD_TRY(pd_enqueue_ns_step(c, cur, c->d_dt, h->zb[kn], h->zb[kn + 1]));
            ListWindow win(c, *h, kn);
            PD_TRY(pd_enqueue_bc_wall(c, nw));
            if (kn == 0) PD_TRY(pd_enqueue_bc_inlet(c, nw, sC));
            if (kn == T) PD_TRY(pd_enqueue_bc_outlet(c, nw, sC));
            PD_TRY(pd_enqueue_ard_vmag_range(c, nw, (long long)h->zb[kn] * P, (long long)h->zb[kn + 1] * P));
            PD_TRY(pd_enqueue_bc_wall_conc(c, sC, true));
            PD_TRY(pd_enqueue_ard_prepass_solids(c, sC));
            if (kn == T && use_side) CUDA_OK(cudaEventRecord(h->ev_sb, side));
        }
        const int kd = k + 2;
        const bool deferred = (kd == T && has_hi) || (kd == 0 && has_lo);   // reads a neighbour rank's planes
        if (kd >= 0 && kd < n && !deferred) {
            if (kd >= T - 1 && use_side) {     // reads |v| of the top chunk: side stream, behind everything the main
                                   // stream has enqueued up to here (B of the chunks around it)
                CUDA_OK(cudaEventRecord(h->ev_x, cs));
                CUDA_OK(cudaStreamWaitEvent(side, h->ev_x, 0));
                PD_TRY(finish_chunk(kd, side, h->s_down2));
            } else {
                PD_TRY(finish_chunk(kd, cs, h->s_down));
            }
        }
    }
    if (c->nranks > 1) {
        // the exchanges of the classical loop bodies (src order: after wall_bc_new; inside the ARD step;
        // after it), once per call and in the same order on every rank, chunked or not
        if (use_side) CUDA_OK(cudaStreamWaitEvent(cs, h->ev_sb, 0));
        PD_TRY(pd_enqueue_halo(c, 0, nw, sC));
        if (has_lo) PD_TRY(pd_enqueue_ard_vmag_range(c, nw, 0, c->own_lo));
        if (has_hi) PD_TRY(pd_enqueue_ard_vmag_range(c, nw, c->own_hi, c->NL));
        PD_TRY(pd_enqueue_halo(c, 3, nw, sC));
        if (has_hi) PD_TRY(finish_chunk(T, cs, h->s_down));
        if (has_lo) PD_TRY(finish_chunk(0, cs, h->s_down));
        PD_TRY(pd_enqueue_halo(c, 1, nw, dC));
    }
    CUDA_OK(cudaEventRecord(h->ev_end2, h->s_down2));
    CUDA_OK(cudaStreamWaitEvent(cs, h->ev_end2, 0));
    CUDA_OK(cudaEventRecord(h->ev_end, h->s_down));
    CUDA_OK(cudaStreamWaitEvent(cs, h->ev_end, 0));
    // std::swap(rho, rho_new) ... (src/pd_ns.cpp:325) and std::swap(C, C_new) (src/coupling.cpp:239)
    c->p_input = cur; pd_pressure_recomputed(c);
    c->cur = nw;
    c->curC = dC;
    pd_touch_flow(c);
    CUDA_OK(cudaStreamSynchronize(cs));
    CUDA_OK(cudaGetLastError());
    return 0;
}

// Timeline of the last chunked pdgpu_step_host call: per chunk (axial order) the milliseconds from
// the start of the call to { upload done, kernels done, download done }. out[3*n_chunks].
extern "C" int pdgpu_step_host_trace(pdgpu_ctx* c, double* out, int cap, int* n_chunks) {
    NEED_GRID(c);
    HostStep* h = c->hs;
    if (!h || h->n <= 1) { if (n_chunks) *n_chunks = h ? h->n : 0; return 0; }
    if (n_chunks) *n_chunks = h->n;
    for (int k = 0; k < h->n && 3 * k + 2 < cap; ++k) {
        float a = 0, b = 0, d = 0;
        CUDA_OK(cudaEventElapsedTime(&a, h->ev_start, h->ev_up[k]));
        CUDA_OK(cudaEventElapsedTime(&b, h->ev_start, h->ev_cmp[k]));
        CUDA_OK(cudaEventElapsedTime(&d, h->ev_start, h->ev_down[k]));
        out[3 * k] = a; out[3 * k + 1] = b; out[3 * k + 2] = d;
    }
    return 0;
}

extern "C" int pdgpu_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) PD_FAIL("pdgpu_host_register: null argument");
    CUDA_OK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}
extern "C" int pdgpu_host_unregister(void* ptr) {
    if (!ptr) PD_FAIL("pdgpu_host_unregister: null argument");
    CUDA_OK(cudaHostUnregister(ptr));
    return 0;
}
