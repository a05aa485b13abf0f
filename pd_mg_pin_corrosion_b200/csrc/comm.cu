// comm.cu -- z-slab halo exchange and scalar all-reduces over NCCL (NVLink 5 / NVSwitch).
//
// One process per GPU; the caller (bench.py / host driver) creates the unique id on rank 0
// with pdgpu_comm_get_uid, broadcasts it with whatever it has (torch.distributed, a file),
// and every rank calls pdgpu_comm_init.  The reference has no distributed layer at all
// (SURVEY.md 5): this is the new axis.  NCCL is resolved lazily with dlopen so that
// single-GPU use and the CPU-only symbol checks never need libnccl.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace {
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) PD_FAIL("cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                   \
    g_nccl.field = (decltype(g_nccl.field))dlsym(h, name);                 \
    if (!g_nccl.field) PD_FAIL("libnccl: missing symbol %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.handle = h;
    return 0;
}
}  // namespace

#define NCCL_OK(expr)                                                                          \
    do {                                                                                       \
        ncclResult_t _r = (expr);                                                              \
        if (_r != ncclSuccess) PD_FAIL("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
    } while (0)

extern "C" int pdgpu_comm_uid_bytes(void) { return (int)sizeof(ncclUniqueId); }

extern "C" int pdgpu_comm_get_uid(void* uid_out) {
    if (!uid_out) PD_FAIL("pdgpu_comm_get_uid: null output");
    PD_TRY(load_nccl());
    ncclUniqueId id;
    NCCL_OK(g_nccl.GetUniqueId(&id));
    memcpy(uid_out, &id, sizeof(id));
    return 0;
}

extern "C" int pdgpu_comm_init(pdgpu_ctx* c, const void* uid, int rank, int nranks) {
    CHECK_CTX(c);
    if (!uid) PD_FAIL("pdgpu_comm_init: null uid");
    if (rank != c->rank || nranks != c->nranks) PD_FAIL("pdgpu_comm_init: rank/nranks differ from pdgpu_create_slab");
    if (nranks == 1) return 0;
    PD_TRY(load_nccl());
    ncclUniqueId id;
    memcpy(&id, uid, sizeof(id));
    ncclComm_t comm = nullptr;
    NCCL_OK(g_nccl.CommInitRank(&comm, nranks, id, rank));
    c->comm = comm;
    pd_invalidate_graphs(c);
    if (c->grid_built) PD_TRY(pd_rebuild_tables(c));   // ghost-wall mirrors need the communicator (collective)
    return 0;
}

int pd_comm_destroy(pdgpu_ctx* c) {
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
    return 0;
}

int pd_comm_allreduce(pdgpu_ctx* c, double* d_buf, int n, int op) {
    if (!c->comm) return 0;
    ncclRedOp_t rop = op == 0 ? ncclSum : (op == 1 ? ncclMax : ncclMin);
    NCCL_OK(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclFloat64, rop, (ncclComm_t)c->comm, c->stream));
    return 0;
}

// In-place merge over ranks of a device buffer in which every element is non-zero on at most one rank
// (elements of `elem` = 1, 4 or 8 bytes): an INTEGER sum, so that the merged array is the owners' bytes
// exactly (a floating-point sum would turn -0.0 into +0.0).
int pd_comm_allreduce_bytes(pdgpu_ctx* c, void* d_buf, size_t bytes, int elem) {
    if (!c->comm) return 0;
    ncclDataType_t ty = elem == 8 ? ncclUint64 : elem == 4 ? ncclUint32 : ncclUint8;
    NCCL_OK(g_nccl.AllReduce(d_buf, d_buf, bytes / (size_t)elem, ty, ncclSum, (ncclComm_t)c->comm, c->stream));
    return 0;
}

// Host scalars combined over the ranks of a slab run (op 0 sum, 1 max, 2 min); no-op for one rank.
extern "C" int pdgpu_comm_allreduce(pdgpu_ctx* c, double* host_vals, int n, int op) {
    CHECK_CTX(c);
    if (!host_vals || n < 0 || n > 1024) PD_FAIL("pdgpu_comm_allreduce: bad arguments");
    if (c->nranks == 1 || !c->comm || n == 0) return 0;
    CUDA_OK(cudaMemcpyAsync(c->d_red, host_vals, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    PD_TRY(pd_comm_allreduce(c, c->d_red, n, op));
    CUDA_OK(cudaMemcpyAsync(host_vals, c->d_red, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

// Exchange the `reach` boundary planes of the listed arrays with both axial neighbours.
//   which 0: rho, p, v[dim] of flow buffer `buf`        (after an NS step + wall_new)
//   which 1: C of buffer `bufC`                          (after an ARD step)
//   which 2: node types, phase, flags and every field of both buffers (after upload / phase change)
//   which 3: salt-layer flags                            (inside an ARD step)
int pd_enqueue_halo(pdgpu_ctx* c, int which, int buf, int bufC) {
    if (!c->comm || c->nranks == 1) return 0;
    if (c->opt_debug_no_halo && which != 4 && which != 2) return 0;   // timing experiments only (wrong results)
    ncclComm_t comm = (ncclComm_t)c->comm;
    long long hp = (long long)c->R * c->P;     // nodes per halo block
    struct Arr { void* p; int bytes; };
    std::vector<Arr> arrs;
    auto add_d = [&](double* p) { if (p) arrs.push_back({p, 8}); };
    auto add_b = [&](uint8_t* p) { if (p) arrs.push_back({p, 1}); };
    if (which == 0 || which == 2) {
        add_d(c->rho[buf]); add_d(c->p[buf]);
        for (int d = 0; d < c->dim; ++d) add_d(c->v[buf][d]);
    }
    if (which == 1 || which == 2) add_d(c->C[bufC]);
    if (which == 2) {
        int ob = 1 - buf, oc = 1 - bufC;
        add_d(c->rho[ob]); add_d(c->p[ob]);
        for (int d = 0; d < c->dim; ++d) add_d(c->v[ob][d]);
        add_d(c->C[oc]);
        add_b(c->type); add_b(c->phase); add_b(c->is_gb); add_b(c->is_precip);
    }
    if (which == 3) { add_b(c->salt); add_d(c->dsol); add_d(c->wpack); }   // wpack: -dsol of ghost solids
    if (which == 4 && c->moff) arrs.push_back({c->moff, 4});
    int lo = c->rank - 1, hi = c->rank + 1;
    NCCL_OK(g_nccl.GroupStart());
    // a failing call inside the group must still close it (an open group swallows every later call)
    ncclResult_t bad = ncclSuccess;
    auto keep = [&](ncclResult_t r) { if (r != ncclSuccess && bad == ncclSuccess) bad = r; };
    for (const Arr& a : arrs) {
        char* base = (char*)a.p;
        size_t nb = (size_t)hp * a.bytes;
        // offsets from pdgpu_slab_layout: send_lo, recv_lo, send_hi, recv_hi
        if (lo >= 0) {
            keep(g_nccl.Send(base + (size_t)c->halo_off[0] * a.bytes, nb, ncclUint8, lo, comm, c->stream));
            keep(g_nccl.Recv(base + (size_t)c->halo_off[1] * a.bytes, nb, ncclUint8, lo, comm, c->stream));
        }
        if (hi < c->nranks) {
            keep(g_nccl.Send(base + (size_t)c->halo_off[2] * a.bytes, nb, ncclUint8, hi, comm, c->stream));
            keep(g_nccl.Recv(base + (size_t)c->halo_off[3] * a.bytes, nb, ncclUint8, hi, comm, c->stream));
        }
    }
    keep(g_nccl.GroupEnd());
    if (bad != ncclSuccess) PD_FAIL("halo exchange (which=%d) failed: %s", which, g_nccl.GetErrorString(bad));
    return 0;
}

extern "C" int pdgpu_halo_exchange(pdgpu_ctx* c, int which) {
    NEED_GRID(c);
    if (which < 0 || which > 2) PD_FAIL("pdgpu_halo_exchange: which must be 0, 1 or 2");
    PD_TRY(pd_enqueue_halo(c, which, c->cur, c->curC));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (which == 2) PD_TRY(pd_rebuild_tables(c));
    return 0;
}
