// host/config.h -- host-side Config of the GPU driver: same `key = value` file format, keys,
// defaults and derived quantities as the reference's Config (src/config.h:4-99,
// src/config.cpp:16-112), written for this repository. to_pod() fills the PdConfig that
// crosses the C ABI (include/pdgpu.h) AFTER compute_derived.
#pragma once
#include <string>

#include "../include/pdgpu.h"

struct HostConfig {
    // grid / geometry
    double dx = 5.0e-6; int m_ratio = 3;
    double R_wire = 40.0e-6, L_wire = 400.0e-6, R_tube = 150.0e-6, L_upstream = 80.0e-6, L_downstream = 80.0e-6;
    // fluid / flow / solid
    double rho_f = 1000.0, mu_f = 1.0e-3, gamma_eos = 7.0, c0 = 0.5, eta_density = 0.1, Q_flow = 1.667e-8;
    double rho_m = 1738.0;
    // transport
    double D_liquid = 1.0e-9, D_grain = 5.0e-11, D_gb = 5.0e-9, D_precip = 5.0e-15, precip_fraction = 0.05;
    double C_solid_init = 1.0, C_liquid_init = 0.0, C_thresh = 0.2, C_sat = 0.9, alpha_art_diff = 0.1;
    double corrosion_decay_l = 0.0;
    // grains
    double grain_size_mean = 40.0e-6, grain_size_std = 5.0e-6; int gb_width_cells = 1, precip_cluster_cells = 0;
    // time stepping / coupling
    double cfl_factor = 0.25, cfl_factor_corr = 0.25;
    int flow_max_iters = 50000; double flow_conv_tol = 5.0e-6, T_final = 32400.0;
    int corrosion_steps_per_check = 200, output_every_flow = 2000, output_every_corr = 100;
    std::string output_dir = "output";
    // implicit branch (lattice: pdgpu_implicit_*, cloud: pdamr_implicit_*) and two-level AMR grid (pdamr_*)
    int use_implicit = 1; double implicit_dt_fraction = 0.5, implicit_dt_max = 60.0;
    int implicit_output_every = 10, diagnostic_every = 1; double newton_tol = 1.0e-8; int newton_max_iter = 20;
    int channel_flow_corrections = 0, use_amr = 0, amr_ratio = 3; double amr_buffer = 50.0e-6;
    // derived
    double delta = 0.0, U_in = 0.0;

    void load(const std::string& filename);     // missing file -> warning + defaults
    bool set(const std::string& key, const std::string& value);   // false: unknown key
    void compute_derived();
    void print(int dim) const;
    PdConfig to_pod() const;
};

// C view for the parity tests: load `path` like the driver does and return the PdConfig that crosses the C ABI plus
// the members that stay on the host, in the order implicit_dt_fraction, implicit_dt_max, implicit_output_every,
// diagnostic_every, newton_tol, newton_max_iter, use_amr, amr_ratio, amr_buffer, precip_fraction, grain_size_mean,
// gb_width_cells, precip_cluster_cells, C_sat (14 doubles).
extern "C" int pdhost_load_config(const char* path, PdConfig* pod, double* host_members, char* output_dir, int output_dir_len);
