// host/config.cpp -- see config.h. Parsing rules follow src/config.cpp:16-96: '#' starts a
// comment, lines without '=' or with an empty key/value are skipped, later keys win, unknown
// keys give a warning on stderr, numbers go through std::stod / std::stoi.
#include "config.h"

#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>

namespace {
std::string strip(const std::string& s) {
    const char* ws = " \t\r\n";
    size_t a = s.find_first_not_of(ws);
    if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(ws);
    return s.substr(a, b - a + 1);
}
}  // namespace

bool HostConfig::set(const std::string& key, const std::string& v) {
    static const std::map<std::string, double HostConfig::*> dbl = {
        {"dx", &HostConfig::dx}, {"R_wire", &HostConfig::R_wire}, {"L_wire", &HostConfig::L_wire},
        {"R_tube", &HostConfig::R_tube}, {"L_upstream", &HostConfig::L_upstream},
        {"L_downstream", &HostConfig::L_downstream}, {"rho_f", &HostConfig::rho_f}, {"mu_f", &HostConfig::mu_f},
        {"gamma_eos", &HostConfig::gamma_eos}, {"c0", &HostConfig::c0}, {"eta_density", &HostConfig::eta_density},
        {"Q_flow", &HostConfig::Q_flow}, {"rho_m", &HostConfig::rho_m}, {"D_liquid", &HostConfig::D_liquid},
        {"D_grain", &HostConfig::D_grain}, {"D_gb", &HostConfig::D_gb}, {"D_precip", &HostConfig::D_precip},
        {"precip_fraction", &HostConfig::precip_fraction}, {"C_solid_init", &HostConfig::C_solid_init},
        {"C_liquid_init", &HostConfig::C_liquid_init}, {"C_thresh", &HostConfig::C_thresh},
        {"C_sat", &HostConfig::C_sat}, {"alpha_art_diff", &HostConfig::alpha_art_diff},
        {"corrosion_decay_l", &HostConfig::corrosion_decay_l}, {"grain_size_mean", &HostConfig::grain_size_mean},
        {"grain_size_std", &HostConfig::grain_size_std}, {"cfl_factor", &HostConfig::cfl_factor},
        {"cfl_factor_corr", &HostConfig::cfl_factor_corr}, {"flow_conv_tol", &HostConfig::flow_conv_tol},
        {"T_final", &HostConfig::T_final}, {"implicit_dt_fraction", &HostConfig::implicit_dt_fraction},
        {"implicit_dt_max", &HostConfig::implicit_dt_max}, {"newton_tol", &HostConfig::newton_tol},
        {"amr_buffer", &HostConfig::amr_buffer}};
    static const std::map<std::string, int HostConfig::*> itg = {
        {"m_ratio", &HostConfig::m_ratio}, {"gb_width_cells", &HostConfig::gb_width_cells},
        {"precip_cluster_cells", &HostConfig::precip_cluster_cells}, {"flow_max_iters", &HostConfig::flow_max_iters},
        {"corrosion_steps_per_check", &HostConfig::corrosion_steps_per_check},
        {"output_every_flow", &HostConfig::output_every_flow}, {"output_every_corr", &HostConfig::output_every_corr},
        {"use_implicit", &HostConfig::use_implicit}, {"implicit_output_every", &HostConfig::implicit_output_every},
        {"diagnostic_every", &HostConfig::diagnostic_every}, {"newton_max_iter", &HostConfig::newton_max_iter},
        {"channel_flow_corrections", &HostConfig::channel_flow_corrections}, {"use_amr", &HostConfig::use_amr},
        {"amr_ratio", &HostConfig::amr_ratio}};
    if (key == "output_dir") { output_dir = v; return true; }
    auto d = dbl.find(key);
    if (d != dbl.end()) { this->*(d->second) = std::stod(v); return true; }
    auto i = itg.find(key);
    if (i != itg.end()) { this->*(i->second) = std::stoi(v); return true; }
    return false;
}

void HostConfig::load(const std::string& filename) {
    std::ifstream f(filename);
    if (!f.is_open()) {
        std::cerr << "Warning: Cannot open config file '" << filename << "', using defaults.\n";
        compute_derived();
        return;
    }
    std::string line;
    while (std::getline(f, line)) {
        size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        line = strip(line);
        size_t eq = line.find('=');
        if (line.empty() || eq == std::string::npos) continue;
        std::string key = strip(line.substr(0, eq)), val = strip(line.substr(eq + 1));
        if (key.empty() || val.empty()) continue;
        if (!set(key, val)) std::cerr << "Warning: Unknown config key '" << key << "'\n";
    }
    compute_derived();
}

void HostConfig::compute_derived() {   // src/config.cpp:98-112
    const double PI = 3.14159265358979323846;
    delta = m_ratio * dx;
    U_in = Q_flow / (PI * R_tube * R_tube);
    if (c0 < 25.0 * U_in) {
        c0 = 25.0 * U_in;
        std::printf("NOTE: Increased c0 to %.4e (25x U_in) for stability.\n", c0);
    }
}

void HostConfig::print(int dim) const {
    std::printf("=== Configuration ===\n  DIM          = %d\n  dx           = %.2e m\n  delta        = %.2e m (m=%d)\n",
                dim, dx, delta, m_ratio);
    std::printf("  R_wire       = %.2e m\n  L_wire       = %.2e m\n  R_tube       = %.2e m\n  U_in         = %.4e m/s\n",
                R_wire, L_wire, R_tube, U_in);
    std::printf("  c0           = %.2f m/s (Mach ~ %.4f)\n  D_liquid     = %.2e m2/s\n  D_grain      = %.2e m2/s\n"
                "  D_gb         = %.2e m2/s\n  T_final      = %.1f s\n  output_dir   = %s\n=====================\n\n",
                c0, U_in / c0, D_liquid, D_grain, D_gb, T_final, output_dir.c_str());
}

PdConfig HostConfig::to_pod() const {
    PdConfig p{};
    p.dx = dx; p.R_wire = R_wire; p.L_wire = L_wire; p.R_tube = R_tube; p.L_upstream = L_upstream;
    p.L_downstream = L_downstream; p.rho_f = rho_f; p.mu_f = mu_f; p.gamma_eos = gamma_eos; p.c0 = c0;
    p.eta_density = eta_density; p.Q_flow = Q_flow; p.rho_m = rho_m; p.D_liquid = D_liquid; p.D_grain = D_grain;
    p.D_gb = D_gb; p.D_precip = D_precip; p.C_solid_init = C_solid_init; p.C_liquid_init = C_liquid_init;
    p.C_thresh = C_thresh; p.C_sat = C_sat; p.alpha_art_diff = alpha_art_diff;
    p.corrosion_decay_l = corrosion_decay_l; p.cfl_factor = cfl_factor; p.cfl_factor_corr = cfl_factor_corr;
    p.flow_conv_tol = flow_conv_tol; p.T_final = T_final; p.delta = delta; p.U_in = U_in;
    p.m_ratio = m_ratio; p.flow_max_iters = flow_max_iters; p.corrosion_steps_per_check = corrosion_steps_per_check;
    p.output_every_flow = output_every_flow; p.output_every_corr = output_every_corr;
    p.channel_flow_corrections = channel_flow_corrections; p.use_implicit = use_implicit;
    return p;
}

extern "C" int pdhost_load_config(const char* path, PdConfig* pod, double* host_members, char* output_dir, int output_dir_len) {
    if (!path || !pod || !host_members) return 1;
    HostConfig c;
    c.load(path);
    *pod = c.to_pod();
    const double v[14] = {c.implicit_dt_fraction, c.implicit_dt_max, (double)c.implicit_output_every, (double)c.diagnostic_every,
                          c.newton_tol, (double)c.newton_max_iter, (double)c.use_amr, (double)c.amr_ratio, c.amr_buffer,
                          c.precip_fraction, c.grain_size_mean, (double)c.gb_width_cells, (double)c.precip_cluster_cells, c.C_sat};
    for (int i = 0; i < 14; ++i) host_members[i] = v[i];
    if (output_dir && output_dir_len > 0) {
        std::snprintf(output_dir, (size_t)output_dir_len, "%s", c.output_dir.c_str());
    }
    return 0;
}
