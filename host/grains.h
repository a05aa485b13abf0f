// host/grains.h -- Voronoi grain structure of the Mg wire (host, one-off initialisation).
// Same algorithm and the same libstdc++ random-number calls as the reference's
// GrainStructure::generate (src/grains.cpp:9-179), so is_grain_boundary / is_precipitate are
// reproduced bit for bit; the lattice is walked directly instead of through the CSR.
#pragma once
#include <cstdint>
#include <vector>

#include "../include/pdgpu.h"

struct GrainParams {
    double grain_size_mean = 40.0e-6;
    double precip_fraction = 0.05;
    int gb_width_cells = 1;
    int precip_cluster_cells = 0;
};

struct GrainStructure {
    std::vector<int> grain_id;
    std::vector<uint8_t> is_grain_boundary, is_precipitate;
    int n_grains = 0;
    void generate(const PdConfig& cfg, const GrainParams& gp, int dim, const uint8_t* node_type, int seed = 42);
    // the same structure with the O(N_solid * n_grains) Voronoi pass, the boundary passes and the cluster
    // growth on the device; 0 on success
    int generate_device(pdgpu_ctx* ctx, const PdConfig& cfg, const GrainParams& gp, int dim, const uint8_t* node_type,
                        int seed = 42);

private:
    int generate_impl(const PdConfig& cfg, const GrainParams& gp, int dim, const uint8_t* node_type, int seed,
                      pdgpu_ctx* ctx);
};
int pdhost_copy_out(const GrainStructure& gs, int* grain_id, uint8_t* is_gb, uint8_t* is_precip, int* n_grains);

extern "C" int pdhost_generate_grains_device(pdgpu_ctx* ctx, const PdConfig* cfg, double grain_size_mean,
                                             double precip_fraction, int gb_width_cells, int precip_cluster_cells,
                                             int dim, const uint8_t* node_type, int seed, int* grain_id,
                                             uint8_t* is_gb, uint8_t* is_precip, int* n_grains);

extern "C" int pdhost_generate_grains(const PdConfig* cfg, double grain_size_mean, double precip_fraction,
                                      int gb_width_cells, int precip_cluster_cells, int dim,
                                      const uint8_t* node_type, int seed, int* grain_id, uint8_t* is_gb,
                                      uint8_t* is_precip, int* n_grains);

// the same generator on the AMR point cloud (arrays of pdamr_get; 2D)
extern "C" int pdhost_generate_grains_cloud(const PdConfig* cfg, double grain_size_mean, double precip_fraction,
                                            int gb_width_cells, int precip_cluster_cells, int N, const double* pos,
                                            const uint8_t* type, const int* nbr_off, const int* nbr_idx,
                                            const double* nbr_dist, int seed, int* grain_id, uint8_t* is_gb,
                                            uint8_t* is_precip, int* n_grains);
