// host/main.cpp -- pd_corrosion_gpu: the reference's main() (src/main.cpp:129-177) over
// libpdgpu.so. The dimension is a run-time argument instead of the PD_DIM compile-time switch.
//
//   pd_corrosion_gpu [config/params.cfg] [--dim 2|3] [--device N] [--dump fields.bin] [--no-vti]
//                    [--checkpoint prefix --checkpoint-every N] [--resume prefix_cNNNN] [--host-grains]
//
// z-slab runs (BASELINE config 4: the coupled run at 1/2/4/8 GPUs): start one process per GPU with
// RANK / WORLD_SIZE / LOCAL_RANK in the environment, e.g.
//   python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 ...
//       ... host/pd_corrosion_gpu cfg --dim 3 --no-vti
// Rank 0 creates the NCCL id and publishes it as <output_dir>/.pdgpu_uid.<MASTER_PORT>; every rank runs
// the coupling loop on its slab, rank 0 writes diagnostics.csv / mass_loss.csv.
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "config.h"
#include "coupling.h"
#include "grains.h"

int run_amr(const HostConfig& cfg, int device, bool write_vtu);   // amr_run.cpp

#define PD(call)                                                                  \
    do {                                                                          \
        if ((call) != 0) {                                                        \
            std::fprintf(stderr, "libpdgpu: %s failed: %s\n", #call, pdgpu_last_error()); \
            return 2;                                                             \
        }                                                                         \
    } while (0)

static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

// NCCL unique id from rank 0 to the others through a file (no MPI / torch in this driver)
static int exchange_uid(const std::string& path, int rank, std::vector<unsigned char>& uid) {
    if (rank == 0) {
        if (pdgpu_comm_get_uid(uid.data()) != 0) return 1;
        const std::string tmp = path + ".tmp";
        FILE* f = std::fopen(tmp.c_str(), "wb");
        if (!f) return 1;
        std::fwrite(uid.data(), 1, uid.size(), f);
        std::fclose(f);
        return std::rename(tmp.c_str(), path.c_str()) != 0;
    }
    for (int tries = 0; tries < 6000; ++tries) {   // up to 10 minutes
        FILE* f = std::fopen(path.c_str(), "rb");
        if (f) {
            size_t got = std::fread(uid.data(), 1, uid.size(), f);
            std::fclose(f);
            if (got == uid.size()) return 0;
        }
        usleep(100000);
    }
    return 1;
}

int main(int argc, char** argv) {
    std::setvbuf(stdout, nullptr, _IONBF, 0);
    const int rank = env_int("RANK", 0), nranks = env_int("WORLD_SIZE", 1);
    std::string cfg_path = "configs/params.cfg", dump;
    int dim = 2, device = env_int("LOCAL_RANK", 0);
    bool no_vti = false, host_grains = false;
    std::string ck_prefix, resume;
    int ck_every = 0;
    for (int a = 1; a < argc; ++a) {
        if (!std::strcmp(argv[a], "--dim") && a + 1 < argc) dim = std::atoi(argv[++a]);
        else if (!std::strcmp(argv[a], "--device") && a + 1 < argc) device = std::atoi(argv[++a]);
        else if (!std::strcmp(argv[a], "--dump") && a + 1 < argc) dump = argv[++a];
        else if (!std::strcmp(argv[a], "--no-vti")) no_vti = true;
        else if (!std::strcmp(argv[a], "--host-grains")) host_grains = true;
        else if (!std::strcmp(argv[a], "--checkpoint") && a + 1 < argc) ck_prefix = argv[++a];
        else if (!std::strcmp(argv[a], "--checkpoint-every") && a + 1 < argc) ck_every = std::atoi(argv[++a]);
        else if (!std::strcmp(argv[a], "--resume") && a + 1 < argc) resume = argv[++a];
        else cfg_path = argv[a];
    }
    if (rank != 0 && !std::freopen("/dev/null", "w", stdout)) return 1;   // rank 0 reports
    std::printf("=== Peridynamic Mg-Pin Corrosion Simulation (B200 path) ===\n  Dimension: %dD\n", dim);
    if (nranks > 1) std::printf("  z-slabs: %d ranks (one GPU each)\n", nranks);
    std::printf("\n");
    auto t0 = std::chrono::steady_clock::now();
    HostConfig cfg;
    cfg.load(cfg_path);
    cfg.print(dim);
    if (cfg.use_amr) {      // two-level AMR cloud (2D, one GPU): its own grid / solver entry points (amr_run.cpp)
        if (dim != 2 || nranks > 1) {
            std::fprintf(stderr, "use_amr = 1 runs in 2D on one GPU (like the reference's AMR grid)\n");
            return 1;
        }
        return run_amr(cfg, device, !no_vti);
    }
    if (cfg.use_implicit && nranks > 1) {
        std::fprintf(stderr, "the implicit branch runs on one GPU only (set use_implicit = 0 for z-slab runs)\n");
        return 1;
    }
    PdConfig pod = cfg.to_pod();
    pdgpu_ctx* ctx = nullptr;
    if (nranks > 1 && (!no_vti || !ck_prefix.empty() || !resume.empty())) {
        std::fprintf(stderr, "z-slab runs write no VTI snapshots / checkpoints yet: pass --no-vti\n");
        return 1;
    }
    PD(pdgpu_create_slab(&pod, dim, device, rank, nranks, &ctx));
    std::printf("Building grid...\n");
    PD(pdgpu_grid_build(ctx));
    if (nranks > 1) {
        mkdir(cfg.output_dir.c_str(), 0755);
        std::vector<unsigned char> uid((size_t)pdgpu_comm_uid_bytes());
        const char* port = std::getenv("MASTER_PORT");
        const std::string uid_path = cfg.output_dir + "/.pdgpu_uid." + (port ? port : "0");
        if (rank == 0) std::remove(uid_path.c_str());
        if (exchange_uid(uid_path, rank, uid)) {
            std::fprintf(stderr, "rank %d: cannot exchange the NCCL id through %s\n", rank, uid_path.c_str());
            return 2;
        }
        PD(pdgpu_comm_init(ctx, uid.data(), rank, nranks));
    }
    PdGridInfo gi;
    PD(pdgpu_grid_info(ctx, &gi));
    std::printf("Grid: Nx=%d Ny=%d Nz=%d  N_total=%lld\n", gi.Nx, gi.Ny, gi.Nz, gi.N_total);
    std::printf("Node types: FLUID=%lld SOLID_MG=%lld WALL=%lld INLET=%lld OUTLET=%lld OUTSIDE=%lld\n", gi.counts[0],
                gi.counts[1], gi.counts[2], gi.counts[3], gi.counts[4], gi.counts[5]);
    std::printf("Neighbor stencil size: %d   bond-updates per NS / ARD step: %lld / %lld\n", gi.n_off, gi.ns_bonds,
                gi.ard_bonds);
    HostState st;
    st.dim = dim; st.Nx = gi.Nx; st.Ny = gi.Ny; st.Nz = gi.Nz; st.N = gi.N_total;
    st.node_type.resize(st.N);
    PD(pdgpu_fields_download_all(ctx, PDGPU_F_NODE_TYPE, st.node_type.data()));   // whole grid on every rank

    std::printf("Generating grain structure...\n");
    GrainParams gp;
    gp.grain_size_mean = cfg.grain_size_mean; gp.precip_fraction = cfg.precip_fraction;
    gp.gb_width_cells = cfg.gb_width_cells; gp.precip_cluster_cells = cfg.precip_cluster_cells;
    GrainStructure grains;
    auto tg = std::chrono::steady_clock::now();
    if (host_grains) {
        grains.generate(pod, gp, dim, st.node_type.data());
    } else if (grains.generate_device(ctx, pod, gp, dim, st.node_type.data()) != 0) {   // lattice passes on the device
        std::fprintf(stderr, "libpdgpu: grain generation failed: %s\n", pdgpu_last_error());
        return 2;
    }
    std::printf("Grain generation: %d grains (%s passes)\n  [Timer] grain_generation: %.3f s\n", grains.n_grains,
                host_grains ? "host" : "device",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - tg).count());
    st.grain_id = grains.grain_id;

    std::printf("Initializing fields...\n");
    PD(pdgpu_fields_init(ctx, grains.is_grain_boundary.data(), grains.is_precipitate.data()));
    st.D_map.assign(st.N, 0.0);   // host-only output field (src/main.cpp:19-112)
    for (long long i = 0; i < st.N; ++i) {
        uint8_t t = st.node_type[i];
        if (t == PDGPU_FLUID || t == PDGPU_INLET || t == PDGPU_OUTLET) st.D_map[i] = cfg.D_liquid;
        else if (t == PDGPU_SOLID_MG)
            st.D_map[i] = grains.is_grain_boundary[i] ? cfg.D_gb : (grains.is_precipitate[i] ? cfg.D_precip : cfg.D_grain);
    }
    std::printf("  [Timer] initialization: %.3f s\n",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());

    CoupledSolver solver;
    solver.rank = rank; solver.nranks = nranks;
    solver.write_vti = !no_vti;
    solver.checkpoint_prefix = ck_prefix; solver.checkpoint_every = ck_every; solver.resume_prefix = resume;
    auto t1 = std::chrono::steady_clock::now();
    solver.run(ctx, st, cfg);
    std::printf("  [Timer] total_simulation: %.3f s\n",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());

    if (!dump.empty()) {   // raw binary state: N, dim, then rho, vel[N][dim], C (FP64)
        std::vector<double> rho(st.N), vel((size_t)st.N * dim), C(st.N);
        PD(pdgpu_fields_download_all(ctx, PDGPU_F_RHO, rho.data()));   // collective for slab runs
        PD(pdgpu_fields_download_all(ctx, PDGPU_F_VEL, vel.data()));
        PD(pdgpu_fields_download_all(ctx, PDGPU_F_C, C.data()));
        if (rank != 0) { PD(pdgpu_destroy(ctx)); return 0; }
        std::ofstream f(dump, std::ios::binary);
        long long hdr[2] = {st.N, dim};
        f.write((const char*)hdr, sizeof(hdr));
        f.write((const char*)rho.data(), sizeof(double) * rho.size());
        f.write((const char*)vel.data(), sizeof(double) * vel.size());
        f.write((const char*)C.data(), sizeof(double) * C.size());
    }
    PD(pdgpu_destroy(ctx));
    return 0;
}
