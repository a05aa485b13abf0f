// host/amr_run.cpp -- the reference's main() + CoupledSolver::run with use_amr = 1 (src/main.cpp:151-174,
// src/coupling.cpp:82-302, explicit ARD branch) over the pdamr_* entry points of libpdgpu.so: two-level grid and
// cell-list neighbours (host code inside the library, bit-identical to the reference), grains on the cloud,
// initialize_fields, then flow solve + IDW refresh / corrosion cycles (explicit, or implicit with use_implicit = 1:
// src/coupling.cpp:154-216 over pdamr_implicit_*) / phase change on the device.
// Writes diagnostics.csv, mass_loss.csv and -- like the reference (src/coupling.cpp:117-121,142-147,242-246,292-296) --
// the state_/flow_/corr_/final_ VTU snapshots of the cloud with simulation.pvd / flow.pvd (host/vtu.cpp: the text of
// VTKWriter::write_vtu, src/vtk_writer.cpp:199-346); --no-vti switches the snapshots off.
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

#include "config.h"
#include "coupling.h"
#include "grains.h"
#include "vtu.h"

#define PDA(call)                                                                         \
    do {                                                                                  \
        if ((call) != 0) {                                                                \
            std::fprintf(stderr, "libpdgpu: %s failed: %s\n", #call, pdgpu_last_error()); \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

int run_amr(const HostConfig& cfg, int device, bool write_vtu) {
    auto t0 = std::chrono::steady_clock::now();
    PdConfig pod = cfg.to_pod();
    pdamr_ctx* a = nullptr;
    PDA(pdamr_create(&pod, cfg.amr_ratio, cfg.amr_buffer, &a));
    std::printf("Building grid...\n");
    PDA(pdamr_build(a));
    PDA(pdamr_build_neighbors(a));
    PdAmrInfo in;
    PDA(pdamr_info(a, &in));
    const int N = (int)in.N_total;
    std::printf("AMR Node types: FLUID=%lld SOLID_MG=%lld WALL=%lld INLET=%lld OUTLET=%lld OUTSIDE=%lld FICT=%lld\n",
                in.counts[0], in.counts[1], in.counts[2], in.counts[3], in.counts[4], in.counts[5], in.counts[6]);
    std::printf("AMR total: %d nodes (fine=%lld, coarse=%lld, fict=%lld)\nCell-list neighbors: %lld total entries\n", N,
                in.n_fine, in.n_coarse, in.n_fict, in.nnz);
    std::vector<double> pos(2 * (size_t)N), dist((size_t)in.nnz);
    std::vector<uint8_t> type(N);
    std::vector<int> off(N + 1), idx((size_t)in.nnz);
    PDA(pdamr_get(a, "pos", pos.data())); PDA(pdamr_get(a, "node_type", type.data()));
    PDA(pdamr_get(a, "nbr_offset", off.data())); PDA(pdamr_get(a, "nbr_index", idx.data()));
    PDA(pdamr_get(a, "nbr_dist", dist.data()));
    std::printf("Generating grain structure...\n");
    std::vector<int> grain_id(N);
    std::vector<uint8_t> is_gb(N), is_precip(N);
    int n_grains = 0;
    if (pdhost_generate_grains_cloud(&pod, cfg.grain_size_mean, cfg.precip_fraction, cfg.gb_width_cells,
                                     cfg.precip_cluster_cells, N, pos.data(), type.data(), off.data(), idx.data(),
                                     dist.data(), 42, grain_id.data(), is_gb.data(), is_precip.data(), &n_grains) != 0) {
        std::fprintf(stderr, "grain generation failed\n");
        return 2;
    }
    std::printf("Grain generation: %d grains, %d boundary nodes\n", n_grains, (int)std::count(is_gb.begin(), is_gb.end(), 1));
    // initialize_fields (src/main.cpp:9-126)
    std::printf("Initializing fields...\n");
    std::vector<double> rho(N), vel(2 * (size_t)N, 0.0), C(N, 0.0);
    std::vector<uint8_t> phase(N, 1);
    const double R2 = cfg.R_tube * cfg.R_tube;
    for (int i = 0; i < N; ++i) {
        const double px = pos[2 * i];
        double rr = (px * px) / R2;
        if (rr > 1.0) rr = 1.0;
        const double v_ax = 1.5 * cfg.U_in * (1.0 - rr);
        rho[i] = type[i] == PDGPU_OUTSIDE ? 0.0 : cfg.rho_f;
        switch (type[i]) {
            case PDGPU_FLUID: case PDGPU_INLET: vel[2 * i + 1] = v_ax; C[i] = cfg.C_liquid_init; break;
            case PDGPU_OUTLET: C[i] = cfg.C_liquid_init; break;
            case PDGPU_SOLID_MG: C[i] = cfg.C_solid_init; phase[i] = 0; break;
            default: break;
        }
    }
    PDA(pdamr_device_init(a, device));
    for (const char* n : {"rho", "rho_new"}) PDA(pdamr_field_set(a, n, rho.data()));
    for (const char* n : {"vel", "vel_new"}) PDA(pdamr_field_set(a, n, vel.data()));
    for (const char* n : {"C", "C_new"}) PDA(pdamr_field_set(a, n, C.data()));
    PDA(pdamr_field_set(a, "phase", phase.data()));
    PDA(pdamr_field_set(a, "is_gb", is_gb.data()));
    PDA(pdamr_field_set(a, "is_precip", is_precip.data()));

    // CoupledSolver::run, explicit branch
    mkdir(cfg.output_dir.c_str(), 0755);
    // snapshots: what the VTU writer reads besides the device fields lives on the host (D_map and
    // grain_id are never read by the solvers); D_map is patched from the node types after every phase change
    std::vector<double> D_map(N), dx_local(N), pressure(N);
    std::vector<int> grid_level(N);
    pdhost_init_dmap(N, type.data(), is_gb.data(), is_precip.data(), cfg.D_liquid, cfg.D_grain, cfg.D_gb, cfg.D_precip,
                     D_map.data());
    PDA(pdamr_get(a, "dx_local", dx_local.data())); PDA(pdamr_get(a, "grid_level", grid_level.data()));
    PvdSeries writer, flow_writer;
    writer.set_path(cfg.output_dir + "/simulation.pvd");
    flow_writer.set_path(cfg.output_dir + "/flow.pvd");
    int frame = 0;
    auto snapshot = [&](const char* prefix, double t, PvdSeries& series, bool count_frame) -> int {
        if (!write_vtu) return 0;
        std::ostringstream ss;                  // make_filename (src/coupling.cpp:10-18) with use_amr = 1
        ss << cfg.output_dir << "/" << prefix << "_" << std::setw(6) << std::setfill('0') << frame << "_t" << std::fixed
           << std::setprecision(1) << t << "s.vtu";
        const std::string fname = ss.str();
        PDA(pdamr_field_get(a, "vel", vel.data())); PDA(pdamr_field_get(a, "pressure", pressure.data()));
        PDA(pdamr_field_get(a, "C", C.data())); PDA(pdamr_field_get(a, "phase", phase.data()));
        PDA(pdamr_field_get(a, "node_type", type.data()));
        if (pdhost_write_vtu(fname.c_str(), N, pos.data(), type.data(), vel.data(), pressure.data(), C.data(), phase.data(),
                             grid_level.data(), dx_local.data(), grain_id.data(), D_map.data(), is_gb.data(),
                             is_precip.data()) != 0)
            return 2;
        series.add_timestep(t, fname);
        if (count_frame) ++frame;
        return 0;
    };
    if (snapshot("state", 0.0, writer, true) != 0) return 2;
    { std::ofstream csv(cfg.output_dir + "/diagnostics.csv", std::ios::trunc);
      csv << "time_s,time_h,pin_mass_loss_pct,solid_nodes,v_max,C_max_fluid\n"; }
    { std::ofstream ml(cfg.output_dir + "/mass_loss.csv", std::ios::trunc); ml << "time_h,pin_mass_loss_pct\n"; }
    std::vector<int> solid0;
    for (int i = 0; i < N; ++i)
        if (type[i] == PDGPU_SOLID_MG) solid0.push_back(i);
    std::printf("Initial solid nodes: %d\nUsing %s ARD solver\n", (int)solid0.size(), cfg.use_implicit ? "IMPLICIT" : "EXPLICIT");
    auto solid_sum = [&](const std::vector<double>& Cc) {
        double s = 0.0;
        for (int i : solid0) s += Cc[i];            // ordered sum (src/coupling.cpp:32-38)
        return s;
    };
    auto diagnostics = [&](double t) -> int {
        PDA(pdamr_field_get(a, "C", C.data())); PDA(pdamr_field_get(a, "vel", vel.data()));
        PDA(pdamr_field_get(a, "node_type", type.data()));
        double loss = (1.0 - solid_sum(C) / (solid0.size() + 1e-30)) * 100.0;
        if (loss < 0.0) loss = 0.0;
        long long solid = 0;
        double vmax = 0.0, cmax = 0.0;
        for (int i = 0; i < N; ++i) {
            if (type[i] == PDGPU_SOLID_MG) ++solid;
            if (type[i] != PDGPU_FLUID) continue;
            vmax = std::max(vmax, std::sqrt(vel[2 * i] * vel[2 * i] + vel[2 * i + 1] * vel[2 * i + 1]));
            cmax = std::max(cmax, C[i]);
        }
        std::printf("  t=%.1f s (%.2f h)  pin_mass_loss=%.2f%%  solid=%lld  v_max=%.3e  C_max_fluid=%.4f\n", t, t / 3600.0, loss,
                    solid, vmax, cmax);
        std::ofstream csv(cfg.output_dir + "/diagnostics.csv", std::ios::app);
        csv << std::scientific << std::setprecision(6) << t << "," << t / 3600.0 << "," << loss << "," << solid << "," << vmax
            << "," << cmax << "\n";
        std::ofstream ml(cfg.output_dir + "/mass_loss.csv", std::ios::app);
        ml << std::fixed << std::setprecision(6) << t / 3600.0 << "," << loss << "\n";
        return 0;
    };
    double t_corr = 0.0;
    int cycle = 0, total_dissolved = 0, total_implicit_steps = 0;
    bool need_flow = true;
    while (t_corr < cfg.T_final) {
        ++cycle;
        std::printf("\n=== Coupling cycle %d, t=%.1f s (%.2f h) ===\n", cycle, t_corr, t_corr / 3600.0);
        if (need_flow) {
            PdSteadyResult r;
            PDA(pdamr_ns_solve_steady(a, &r, 1));
            PDA(pdamr_update_fictitious(a));        // src/coupling.cpp:139
            need_flow = false;
            if (snapshot("flow", t_corr, flow_writer, true) != 0) return 2;
        }
        PDA(pdamr_field_get(a, "C", C.data()));
        double vl = 1.0 - solid_sum(C) / (solid0.size() + 1e-30);
        PDA(pdamr_ard_set_volume_loss(a, vl < 0.0 ? 0.0 : vl));
        if (cfg.use_implicit) {                         // src/coupling.cpp:154-216 on the cloud (pdamr_implicit_*)
            PDA(pdamr_implicit_assemble(a));
            int implicit_step = 0;
            bool dissolved = false;
            const double t_start = t_corr;
            while (implicit_step < cfg.corrosion_steps_per_check && t_corr < cfg.T_final && !dissolved) {
                double dt_impl = 0.0;
                PDA(pdamr_implicit_compute_dt(a, cfg.implicit_dt_fraction, cfg.implicit_dt_max, &dt_impl));
                PDA(pdamr_bc(a, 0)); PDA(pdamr_bc(a, 1)); PDA(pdamr_bc(a, 4));
                PdLinSolveInfo info;
                // true relative residual 1e-12 (the reference: 1e-10 on its preconditioned residual), axial-sweep GMRES(50)
                PDA(pdamr_implicit_step(a, dt_impl, 1e-12, 50, 2000, 1, &info));
                std::printf("    Linear solve: GMRES %d iters, |res|=%.2e\n", info.iters, info.rel_res);
                PDA(pdamr_bc(a, 6));                    // smooth_boundary_concentration
                PDA(pdamr_update_fictitious(a));
                t_corr += dt_impl;
                ++implicit_step;
                ++total_implicit_steps;
                if (total_implicit_steps % cfg.diagnostic_every == 0 && diagnostics(t_corr) != 0) return 2;
                if (total_implicit_steps % cfg.implicit_output_every == 0 && snapshot("corr", t_corr, writer, true) != 0) return 2;
                PDA(pdamr_field_get(a, "C", C.data())); PDA(pdamr_field_get(a, "node_type", type.data()));
                for (int i = 0; i < N && !dissolved; ++i) dissolved = type[i] == PDGPU_SOLID_MG && C[i] < cfg.C_thresh;
            }
            std::printf("  Implicit cycle: %d steps, t=%.2f to %.2f s (%.4f h)\n", implicit_step, t_start, t_corr, t_corr / 3600.0);
        }
        double dtc = 0.0;
        if (!cfg.use_implicit) {
            PDA(pdamr_ard_compute_dt(a, &dtc));
            std::printf("  Corrosion dt = %.4e s\n", dtc);
        }
        int step = 0;
        const int n_steps = cfg.use_implicit ? 0 : cfg.corrosion_steps_per_check, every = cfg.output_every_corr;
        while (step < n_steps) {
            const int chunk = std::min(n_steps - step, every - step % every);
            int done = 0;
            while (done < chunk) {                  // t_corr advances per step and ends the cycle (:237, :248)
                t_corr += dtc;
                ++done;
                if (t_corr >= cfg.T_final) break;
            }
            PDA(pdamr_ard_iterate(a, done, dtc));
            step += done;
            if (done == chunk && step % every == 0) {       // snapshot, then the diagnostics row (:242-248)
                if (snapshot("corr", t_corr, writer, true) != 0 || diagnostics(t_corr) != 0) return 2;
            }
            if (t_corr >= cfg.T_final) break;
        }
        int n = 0;
        PDA(pdamr_phase_change(a, &n));
        total_dissolved += n;
        if (n > 0) {
            std::printf("  Phase change: %d nodes dissolved (total: %d)\n", n, total_dissolved);
            need_flow = true;
        } else {
            std::printf("  No phase changes this cycle\n");
        }
        if (n > 0) {                                // apply_phase_change sets D_map = D_liquid (src/pd_ard.cpp:202)
            std::vector<uint8_t> before = type;
            PDA(pdamr_field_get(a, "node_type", type.data()));
            for (int i = 0; i < N; ++i)
                if (before[i] == PDGPU_SOLID_MG && type[i] == PDGPU_FLUID) D_map[i] = cfg.D_liquid;
        }
        PDA(pdamr_field_get(a, "node_type", type.data()));
        if (std::count(type.begin(), type.end(), (uint8_t)PDGPU_SOLID_MG) == 0) {
            std::printf("\n=== All solid nodes dissolved at t=%.1f s (%.2f h) ===\n", t_corr, t_corr / 3600.0);
            break;
        }
    }
    if (snapshot("final", t_corr, writer, false) != 0) return 2;
    std::printf("\n=== Simulation complete ===\n  Final time: %.1f s (%.2f h)\n  [Timer] total_simulation: %.3f s\n", t_corr,
                t_corr / 3600.0, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    pdamr_destroy(a);
    return 0;
}
