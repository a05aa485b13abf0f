// host/coupling.cpp -- see coupling.h. Control flow follows src/coupling.cpp:82-302 (explicit
// branch); every solver call is one C-ABI call into libpdgpu.so, and the corrosion steps
// between two output points run device resident (pdgpu_ard_iterate) instead of one host call
// per step. VTI snapshots are formatted on the device (pdgpu_vti_write); the PVD files are host text.
#include "coupling.h"

#include <sys/stat.h>

#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>

#define PD(call)                                                                  \
    do {                                                                          \
        if ((call) != 0) {                                                        \
            std::fprintf(stderr, "libpdgpu: %s failed: %s\n", #call, pdgpu_last_error()); \
            std::exit(2);                                                         \
        }                                                                         \
    } while (0)

// ordered host sum over the initially solid nodes: (1 - sum/n) cancels catastrophically
// (SURVEY.md 7.2-4), so the values are gathered and added in index order like
// src/coupling.cpp:32-38.
double CoupledSolver::solid_C_sum(pdgpu_ctx* ctx) {
    std::vector<double> vals(initial_solid_indices_.size());
    PD(pdgpu_gather(ctx, PDGPU_F_C, initial_solid_indices_.data(), (long long)vals.size(), vals.data()));
    double s = 0.0;
    for (double v : vals) s += v;
    return s;
}

void CoupledSolver::write_diagnostics(pdgpu_ctx* ctx, double t_corr, const HostConfig& cfg) {   // :20-68
    PdDiag d;
    PD(pdgpu_diag(ctx, &d));
    double n0 = (double)initial_solid_indices_.size();
    double loss = (1.0 - solid_C_sum(ctx) / (n0 + 1e-30)) * 100.0;   // collective gather: same sum on every rank
    if (loss < 0.0) loss = 0.0;
    if (rank != 0) return;
    std::printf("  t=%.1f s (%.2f h)  pin_mass_loss=%.2f%%  solid=%lld  v_max=%.3e  C_max_fluid=%.4f\n", t_corr,
                t_corr / 3600.0, loss, d.solid_count, d.v_max, d.C_max_fluid);
    std::ofstream csv(cfg.output_dir + "/diagnostics.csv", std::ios::app);
    csv << std::scientific << std::setprecision(6) << t_corr << "," << t_corr / 3600.0 << "," << loss << ","
        << d.solid_count << "," << d.v_max << "," << d.C_max_fluid << "\n";
    std::ofstream ml(cfg.output_dir + "/mass_loss.csv", std::ios::app);
    ml << std::fixed << std::setprecision(6) << t_corr / 3600.0 << "," << loss << "\n";
}

static std::string make_filename(const HostConfig& cfg, const std::string& prefix, double time_s, int frame) {   // :10-18
    std::ostringstream ss;
    ss << cfg.output_dir << "/" << prefix << "_" << std::setw(6) << std::setfill('0') << frame << "_t" << std::fixed
       << std::setprecision(1) << time_s << "s.vti";
    return ss.str();
}

void PvdSeries::add_timestep(double time, const std::string& file) {   // src/vtk_writer.cpp:150-186
    entries_.push_back({time, file});
    if (path_.empty()) return;
    std::ofstream out(path_);
    if (!out.is_open()) {
        std::fprintf(stderr, "Error: Cannot open PVD file '%s'\n", path_.c_str());
        return;
    }
    std::string dir;
    auto slash = path_.find_last_of('/');
    if (slash != std::string::npos) dir = path_.substr(0, slash + 1);
    out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"Collection\" version=\"1.0\" byte_order=\"LittleEndian\">\n  <Collection>\n";
    for (auto& e : entries_) {
        std::string rel = e.second;
        if (!dir.empty() && rel.find(dir) == 0) rel = rel.substr(dir.size());
        out << "    <DataSet timestep=\"" << std::scientific << std::setprecision(6) << e.first << "\" file=\"" << rel
            << "\"/>\n";
    }
    out << "  </Collection>\n</VTKFile>\n";
    out.close();
    std::printf("  Wrote PVD file: %s (%zu timesteps)\n", path_.c_str(), entries_.size());
}

// ---- driver checkpoint: plain binary, same machine ------------------------------------------------
template <typename T>
static void put(std::ostream& o, const T& v) { o.write((const char*)&v, sizeof(T)); }
template <typename T>
static void get(std::istream& i, T& v) { i.read((char*)&v, sizeof(T)); }

void PvdSeries::save(std::ostream& out) const {
    put(out, (long long)entries_.size());
    for (auto& e : entries_) {
        put(out, e.first);
        put(out, (long long)e.second.size());
        out.write(e.second.data(), (std::streamsize)e.second.size());
    }
}
void PvdSeries::load(std::istream& in) {
    long long n = 0;
    get(in, n);
    entries_.clear();
    if (!in || n < 0 || n > (1LL << 24)) { in.setstate(std::ios::failbit); return; }   // corrupt file
    for (long long k = 0; k < n; ++k) {
        double t = 0.0; long long len = 0;
        get(in, t); get(in, len);
        if (!in || len < 0 || len > 4096) { in.setstate(std::ios::failbit); return; }   // a file name, not a blob
        std::string f((size_t)len, '\0');
        in.read(&f[0], (std::streamsize)len);
        entries_.push_back({t, f});
    }
}

void CoupledSolver::save_driver_state(const std::string& path, const HostState& st, double t_corr, int cycle,
                                      bool need_flow) const {
    // temporary name + rename: the .drv file appears only complete, and after its .pdck partner
    const std::string tmp = path + ".tmp";
    std::ofstream o(tmp, std::ios::binary | std::ios::trunc);
    const char magic[8] = {'P', 'D', 'D', 'R', 'V', 'C', 'K', '1'};
    o.write(magic, 8);
    put(o, t_corr); put(o, cycle); put(o, (int)need_flow); put(o, frame_count_); put(o, total_dissolved_);
    put(o, dissolved_since_flow_);
    put(o, (long long)initial_solid_indices_.size());
    o.write((const char*)initial_solid_indices_.data(), (std::streamsize)(sizeof(int) * initial_solid_indices_.size()));
    put(o, st.N);
    o.write((const char*)st.node_type.data(), (std::streamsize)st.N);
    o.write((const char*)st.D_map.data(), (std::streamsize)(sizeof(double) * st.N));
    writer_.save(o);
    flow_writer_.save(o);
    o.flush();
    o.close();
    if (!o || std::rename(tmp.c_str(), path.c_str()) != 0) std::fprintf(stderr, "cannot write driver checkpoint %s\n", path.c_str());
}

bool CoupledSolver::load_driver_state(const std::string& path, HostState& st, double* t_corr, int* cycle, bool* need_flow) {
    std::ifstream in(path, std::ios::binary);
    char magic[8];
    if (!in.read(magic, 8) || std::string(magic, 8) != "PDDRVCK1") return false;
    int nf = 0;
    long long n = 0, N = 0;
    get(in, *t_corr); get(in, *cycle); get(in, nf); get(in, frame_count_); get(in, total_dissolved_);
    get(in, dissolved_since_flow_);
    *need_flow = nf != 0;
    get(in, n);
    if (!in || n < 0 || n > st.N) return false;   // corrupt file: never size a vector from unchecked input
    initial_solid_indices_.resize((size_t)n);
    in.read((char*)initial_solid_indices_.data(), (std::streamsize)(sizeof(int) * n));
    get(in, N);
    if (N != st.N) return false;
    in.read((char*)st.node_type.data(), (std::streamsize)N);
    in.read((char*)st.D_map.data(), (std::streamsize)(sizeof(double) * N));
    writer_.load(in);
    flow_writer_.load(in);
    return (bool)in;
}

void CoupledSolver::snapshot(pdgpu_ctx* ctx, const HostState& st, const HostConfig& cfg, const char* prefix, double t,
                             PvdSeries& series, bool count_frame) {
    if (!write_vti) return;
    std::string fname = make_filename(cfg, prefix, t, frame_count_);
    PD(pdgpu_vti_write(ctx, fname.c_str(), st.grain_id.data(), st.D_map.data(), nullptr, nullptr));
    series.add_timestep(t, fname);
    if (count_frame) frame_count_++;
}

double CoupledSolver::run(pdgpu_ctx* ctx, HostState& st, const HostConfig& cfg, bool verbose) {
    mkdir(cfg.output_dir.c_str(), 0755);
    if (rank == 0) {
        writer_.set_path(cfg.output_dir + "/simulation.pvd");
        flow_writer_.set_path(cfg.output_dir + "/flow.pvd");
    }
    const bool resuming = !resume_prefix.empty();
    if (!resuming && rank == 0) {
        std::ofstream csv(cfg.output_dir + "/diagnostics.csv", std::ios::trunc);
        csv << "time_s,time_h,pin_mass_loss_pct,solid_nodes,v_max,C_max_fluid\n";
        std::ofstream ml(cfg.output_dir + "/mass_loss.csv", std::ios::trunc);
        ml << "time_h,pin_mass_loss_pct\n";
    }
    double t_corr = 0.0;
    int cycle = 0;
    bool need_flow_solve = true;
    dissolved_since_flow_ = 0;
    if (resuming) {
        if (!load_driver_state(resume_prefix + ".drv", st, &t_corr, &cycle, &need_flow_solve)) {
            std::fprintf(stderr, "cannot read driver checkpoint %s.drv\n", resume_prefix.c_str());
            std::exit(2);
        }
        PD(pdgpu_checkpoint_load(ctx, (resume_prefix + ".pdck").c_str()));
        std::printf("Resumed from %s: cycle %d, t=%.6e s, %zu initial solid nodes\n", resume_prefix.c_str(), cycle, t_corr,
                    initial_solid_indices_.size());
    } else {
        initial_solid_indices_.clear();
        for (long long i = 0; i < st.N; ++i)
            if (st.node_type[i] == PDGPU_SOLID_MG) initial_solid_indices_.push_back((int)i);
    }
    const double n0 = (double)initial_solid_indices_.size();
    std::printf("Initial solid nodes: %zu\nUsing %s ARD solver\n", initial_solid_indices_.size(),
                cfg.use_implicit ? "IMPLICIT (matrix-free GMRES)" : "EXPLICIT");

    if (!resuming) snapshot(ctx, st, cfg, "state", 0.0, writer_, true);   // :117-122
    std::vector<int> dissolved(std::max<size_t>(initial_solid_indices_.size(), 1));
    while (t_corr < cfg.T_final) {
        ++cycle;
        std::printf("\n=== Coupling cycle %d, t=%.1f s (%.2f h) ===\n", cycle, t_corr, t_corr / 3600.0);
        if (need_flow_solve) {   // phase 1 :136-151
            std::printf("  Flow re-solve triggered (%d nodes dissolved since last flow solve)\n", dissolved_since_flow_);
            PdSteadyResult r;
            PD(pdgpu_ns_solve_steady(ctx, &r, verbose && rank == 0 ? 1 : 0));
            dissolved_since_flow_ = 0;
            need_flow_solve = false;
            snapshot(ctx, st, cfg, "flow", t_corr, flow_writer_, true);   // :143-148
        } else {
            std::printf("  Skipping flow solve (no dissolution since last flow solve)\n");
        }
        double vol_loss = 1.0 - solid_C_sum(ctx) / (n0 + 1e-30);
        if (vol_loss < 0.0) vol_loss = 0.0;
        PD(pdgpu_ard_set_volume_loss(ctx, vol_loss));
        if (cfg.use_implicit) {   // phase 2, implicit :154-216 (operator matrix-free on the device, GMRES)
            PD(pdgpu_implicit_assemble(ctx));                       // once per coupling cycle
            int implicit_step = 0, below = 0;
            const double t_cycle_start = t_corr;
            PdLinSolveInfo info = {0, 0, 0.0, 0};
            while (implicit_step < cfg.corrosion_steps_per_check && t_corr < cfg.T_final && below == 0) {
                double dt_impl = 0.0;
                PD(pdgpu_implicit_compute_dt(ctx, cfg.implicit_dt_fraction, cfg.implicit_dt_max, &dt_impl));
                PD(pdgpu_bc_inlet(ctx));
                PD(pdgpu_bc_outlet(ctx));
                PD(pdgpu_bc_wall_conc(ctx));
                PD(pdgpu_implicit_step(ctx, dt_impl, 1e-10, 50, 2000, 2, &info));
                std::printf("    Linear solve: GMRES %d iters, |res|=%.2e%s\n", info.iters, info.rel_res,
                            info.converged ? "" : "  (NOT converged)");
                PD(pdgpu_bc_smooth_conc(ctx));
                t_corr += dt_impl;
                ++implicit_step;
                ++total_implicit_steps_;
                if (total_implicit_steps_ % cfg.diagnostic_every == 0) write_diagnostics(ctx, t_corr, cfg);
                if (total_implicit_steps_ % cfg.implicit_output_every == 0) snapshot(ctx, st, cfg, "corr", t_corr, writer_, true);
                PD(pdgpu_solid_below_thresh(ctx, &below));
            }
            std::printf("  Implicit cycle: %d steps, t=%.2f to %.2f s (%.4f h)\n", implicit_step, t_cycle_start, t_corr,
                        t_corr / 3600.0);
        } else {
        // phase 2, explicit :217-253
        double dt_corr = 0.0;
        PD(pdgpu_ard_compute_dt(ctx, &dt_corr));
        std::printf("  Corrosion dt = %.4e s\n", dt_corr);
        int step = 0;
        const int n_steps = cfg.corrosion_steps_per_check, every = cfg.output_every_corr;
        while (step < n_steps) {
            // device-resident run up to the next output point; the reference leaves the cycle
            // as soon as t_corr >= T_final (:251), so count the steps on the host first
            int chunk = std::min(n_steps - step, every - step % every), done = 0;
            double t_probe = t_corr;
            while (done < chunk) {
                t_probe += dt_corr;
                ++done;
                if (t_probe >= cfg.T_final) break;
            }
            PD(pdgpu_ard_iterate(ctx, done, dt_corr));
            for (int s = 0; s < done; ++s) t_corr += dt_corr;   // same additions as the reference
            step += done;
            if (step % every == 0) {
                snapshot(ctx, st, cfg, "corr", t_corr, writer_, true);   // :242-247
                write_diagnostics(ctx, t_corr, cfg);
            }
            if (t_corr >= cfg.T_final) break;
        }
        }
        // phase 3 :255-290
        int n_dissolved = 0;
        PD(pdgpu_phase_change(ctx, &n_dissolved, dissolved.data(), (int)dissolved.size()));   // this rank's slab
        const int n_local = n_dissolved;
        if (nranks > 1) {   // all ranks take the same branch below
            double tot = (double)n_dissolved;
            PD(pdgpu_comm_allreduce(ctx, &tot, 1, 0));
            n_dissolved = (int)tot;
        }
        total_dissolved_ += n_dissolved;
        dissolved_since_flow_ += n_dissolved;
        if (n_dissolved > 0) {
            std::printf("  Phase change: %d nodes dissolved (total: %d, since flow: %d)\n", n_dissolved,
                        total_dissolved_, dissolved_since_flow_);
            for (int t = 0; t < n_local; ++t) {   // host mirrors used for output only (this rank's nodes)
                st.node_type[dissolved[t]] = PDGPU_FLUID;
                st.D_map[dissolved[t]] = cfg.D_liquid;
            }
            need_flow_solve = true;   // neighbour tables were rebuilt inside pdgpu_phase_change
        } else {
            std::printf("  No phase changes this cycle\n");
        }
        if (checkpoint_every > 0 && !checkpoint_prefix.empty() && cycle % checkpoint_every == 0) {
            char tag[32];
            std::snprintf(tag, sizeof(tag), "_c%04d", cycle);
            const std::string base = checkpoint_prefix + tag;
            PD(pdgpu_checkpoint_save(ctx, (base + ".pdck").c_str(), nullptr));
            save_driver_state(base + ".drv", st, t_corr, cycle, need_flow_solve);
            std::printf("  Checkpoint written: %s.pdck / .drv\n", base.c_str());
        }
        PdDiag d;
        PD(pdgpu_diag(ctx, &d));
        if (d.solid_count == 0) {
            std::printf("\n=== All solid nodes dissolved at t=%.1f s (%.2f h) ===\n", t_corr, t_corr / 3600.0);
            break;
        }
    }
    snapshot(ctx, st, cfg, "final", t_corr, writer_, false);   // :292-296
    std::printf("\n=== Simulation complete ===\n  Final time: %.1f s (%.2f h)\n", t_corr, t_corr / 3600.0);
    return t_corr;
}
