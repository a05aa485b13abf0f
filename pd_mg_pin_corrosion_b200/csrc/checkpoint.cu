// checkpoint.cu -- binary checkpoint / resume of the device-resident state (the
// reference has no resume; a multi-day 3D run restarts from t = 0 after any interruption).
//
// File = header (magic, version, dim, lattice, PdConfig, buffer flags, volume loss) followed by the
// raw arrays of the OWNED nodes: node types, phase / grain flags, and BOTH ping-pong copies of
// rho, p, v and C.  Everything a later step reads is in there (p is the Horner-evaluated shadow
// of rho, not a derived quantity at the bit level), so a run continued from a checkpoint is
// bit-identical to the uninterrupted one (tests/test_gpu_checkpoint.py).  Data moves through the
// pinned staging pool of iopool.cuh in both directions.
#include <cstring>
#include <string>

#include "common.cuh"
#include "iopool.cuh"

namespace {

struct CkHeader {
    char magic[8];              // "PDGPUCK1"
    int version, dim, Nx, Ny, Nz, R;
    long long N;                // owned nodes in the file
    int cur, curC, p_input, wallC_pending, wallC_src, pad;
    double volume_loss;
    PdConfig cfg;
};

struct Arr { void* ptr; size_t bytes; };

// Run-control members (how long, how often) may differ between the run that wrote the checkpoint
// and the one that resumes it; everything that shapes the lattice or the physics must match.
PdConfig physics_only(PdConfig k) {
    k.T_final = 0.0;
    k.flow_max_iters = 0;
    k.corrosion_steps_per_check = 0;
    k.output_every_flow = 0;
    k.output_every_corr = 0;
    k.reserved = 0;
    return k;
}

std::vector<Arr> arrays(pdgpu_ctx* c) {
    const long long lo = c->own_lo, n = c->own_hi - c->own_lo;
    std::vector<Arr> a;
    auto b8 = [&](uint8_t* p) { a.push_back({p + lo, (size_t)n}); };
    auto f64 = [&](double* p) { a.push_back({p + lo, (size_t)n * 8}); };
    b8(c->type); b8(c->phase); b8(c->is_gb); b8(c->is_precip);
    for (int k = 0; k < 2; ++k) {
        f64(c->rho[k]); f64(c->p[k]); f64(c->C[k]);
        for (int d = 0; d < c->dim; ++d) f64(c->v[k][d]);
    }
    return a;
}

}  // namespace

extern "C" int pdgpu_checkpoint_save(pdgpu_ctx* c, const char* path, long long* bytes_out) {
    NEED_FIELDS(c);
    if (!path) PD_FAIL("pdgpu_checkpoint_save: null path");
    if (c->nranks > 1) PD_FAIL("pdgpu_checkpoint_save: slab contexts are not supported yet (one file per rank needed)");
    PD_TRY(pd_flush_wall_c(c));   // an owed wall-concentration BC becomes part of the saved state
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CkHeader h;
    std::memset(&h, 0, sizeof(h));
    std::memcpy(h.magic, "PDGPUCK1", 8);
    h.version = 1; h.dim = c->dim; h.Nx = c->Nx; h.Ny = c->Ny; h.Nz = c->Nz; h.R = c->R;
    h.N = c->own_hi - c->own_lo;
    h.cur = c->cur; h.curC = c->curC; h.p_input = c->p_input;
    h.wallC_pending = c->wallC_pending ? 1 : 0; h.wallC_src = c->wallC_src;
    h.volume_loss = c->volume_loss;
    h.cfg = c->cfg;
    // written under a temporary name, synced, then renamed: a crash never leaves a torn file under `path`
    const std::string tmp = std::string(path) + ".tmp";
    const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) PD_FAIL("cannot open checkpoint file '%s'", tmp.c_str());
    pdio::IoPool* io = nullptr;
    int rc = pdio::io_pool(c, &io);
    off_t off = 0;
    if (!rc && ::pwrite(fd, &h, sizeof(h), 0) != (ssize_t)sizeof(h)) rc = 1;
    off += (off_t)sizeof(h);
    {
        pdio::IoRun run(io, fd, c->device);
        for (const Arr& a : arrays(c)) {
            if (rc || run.failed()) break;
            for (size_t done = 0; done < a.bytes; done += pdio::kIoChunk) {
                const size_t len = std::min(pdio::kIoChunk, a.bytes - done);
                const int k = run.acquire();
                if (k < 0) break;
                if (cudaMemcpyAsync(io->buf[k], (const char*)a.ptr + done, len, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                    cudaEventRecord(io->ev[k], c->stream) != cudaSuccess) { run.fail(); break; }
                run.submit(k, len, off + (off_t)done);
            }
            off += (off_t)a.bytes;
        }
        run.finish();
        if (run.failed()) rc = 1;
    }
    if (!rc && ::fsync(fd) != 0) rc = 1;
    ::close(fd);
    if (!rc && ::rename(tmp.c_str(), path) != 0) rc = 1;
    if (rc) { ::unlink(tmp.c_str()); PD_FAIL("pdgpu_checkpoint_save: writing '%s' failed", path); }
    if (bytes_out) *bytes_out = (long long)off;
    return 0;
}

extern "C" int pdgpu_checkpoint_load(pdgpu_ctx* c, const char* path) {
    NEED_GRID(c);
    if (!path) PD_FAIL("pdgpu_checkpoint_load: null path");
    if (c->nranks > 1) PD_FAIL("pdgpu_checkpoint_load: slab contexts are not supported yet");
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) PD_FAIL("cannot open checkpoint file '%s'", path);
    CkHeader h;
    if (::pread(fd, &h, sizeof(h), 0) != (ssize_t)sizeof(h) || std::memcmp(h.magic, "PDGPUCK1", 8) != 0 || h.version != 1) {
        ::close(fd);
        PD_FAIL("'%s' is not a pdgpu checkpoint (version 1)", path);
    }
    const PdConfig want = physics_only(c->cfg), have = physics_only(h.cfg);
    if (h.dim != c->dim || h.Nx != c->Nx || h.Ny != c->Ny || h.Nz != c->Nz || h.R != c->R ||
        h.N != c->own_hi - c->own_lo || std::memcmp(&have, &want, sizeof(PdConfig)) != 0) {
        ::close(fd);
        PD_FAIL("checkpoint '%s' was written for another configuration or lattice (%dD %dx%dx%d)", path, h.dim, h.Nx,
                h.Ny, h.Nz);
    }
    pdio::IoPool* io = nullptr;
    if (pdio::io_pool(c, &io)) { ::close(fd); return 1; }
    std::lock_guard<std::mutex> pool_guard(io->busy);
    // from here on the device arrays are being overwritten: a failure leaves the context without fields
    c->fields_ready = false;
    pd_invalidate_graphs(c);
    off_t off = (off_t)sizeof(h);
    int rc = 0, k = 0;
    for (const Arr& a : arrays(c)) {
        for (size_t done = 0; done < a.bytes && !rc; done += pdio::kIoChunk) {
            const size_t len = std::min(pdio::kIoChunk, a.bytes - done);
            // buffer k is free again when the copy recorded with it has finished
            if (cudaEventSynchronize(io->ev[k]) != cudaSuccess) { rc = 1; break; }
            size_t r = 0;
            while (r < len) {
                ssize_t got = ::pread(fd, io->buf[k] + r, len - r, off + (off_t)(done + r));
                if (got <= 0) { rc = 1; break; }
                r += (size_t)got;
            }
            if (rc) break;
            if (cudaMemcpyAsync((char*)a.ptr + done, io->buf[k], len, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                cudaEventRecord(io->ev[k], c->stream) != cudaSuccess) { rc = 1; break; }
            k = (k + 1) % pdio::kIoBufs;
        }
        if (rc) break;
        off += (off_t)a.bytes;
    }
    ::close(fd);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = 1;
    if (rc) PD_FAIL("pdgpu_checkpoint_load: reading '%s' failed (truncated file?)", path);
    c->cur = h.cur; c->curC = h.curC; c->p_input = h.p_input;
    c->wallC_pending = h.wallC_pending != 0; c->wallC_src = h.wallC_src;
    c->volume_loss = h.volume_loss;
    c->fields_ready = true;
    PD_TRY(pd_rebuild_tables(c));   // node types may differ from the freshly built grid (dissolved nodes)
    return 0;
}
