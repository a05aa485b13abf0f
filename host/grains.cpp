// host/grains.cpp -- see grains.h.  Compiled with -ffp-contract=off; the places where the
// reference's Release build fuses a multiply-add (node positions, squared distances) are
// written with std::fma so that nearest-seed ties resolve exactly as in the reference.
#include "grains.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <random>

namespace {
constexpr double PI = 3.14159265358979323846;
constexpr uint8_t SOLID = PDGPU_SOLID_MG, OUTSIDE = PDGPU_OUTSIDE;

struct Lattice {
    int dim, Nx, Ny, Nz;
    double dx, o[3];
    void coords(int n, int* i, int* j, int* k) const {
        *k = n / (Nx * Ny);
        int rem = n % (Nx * Ny);
        *j = rem / Nx;
        *i = rem % Nx;
    }
    void pos(int n, double p[3]) const {   // src/grid.cpp:88-92 (one fma each in the Release build)
        int i, j, k;
        coords(n, &i, &j, &k);
        p[0] = std::fma((double)i, dx, o[0]);
        p[1] = std::fma((double)j, dx, o[1]);
        p[2] = dim == 3 ? std::fma((double)k, dx, o[2]) : 0.0;
    }
    // norm(a - b) of src/utils.h:16-24: s = fma-chain over the components, then sqrt
    double dist(const double a[3], const double b[3]) const {
        double s = 0.0;
        for (int d = 0; d < dim; ++d) {
            double t = a[d] - b[d];
            s = std::fma(t, t, s);
        }
        return std::sqrt(s);
    }
    // immediate neighbours (|d|_inf <= 1) that exist in the reference's CSR: in the box and not OUTSIDE
    template <class F>
    void for_immediate(int n, const uint8_t* type, F&& f) const {
        int i, j, k;
        coords(n, &i, &j, &k);
        int klo = dim == 3 ? -1 : 0, khi = dim == 3 ? 1 : 0;
        for (int dk = klo; dk <= khi; ++dk)
            for (int dj = -1; dj <= 1; ++dj)
                for (int di = -1; di <= 1; ++di) {
                    if (!di && !dj && !dk) continue;
                    int ni = i + di, nj = j + dj, nk = k + dk;
                    if (ni < 0 || ni >= Nx || nj < 0 || nj >= Ny || nk < 0 || nk >= Nz) continue;
                    int nn = (nk * Ny + nj) * Nx + ni;
                    if (type[nn] == OUTSIDE) continue;
                    if (f(nn)) return;
                }
    }
};
}  // namespace

void GrainStructure::generate(const PdConfig& cfg, const GrainParams& gp, int dim, const uint8_t* type, int seed) {
    generate_impl(cfg, gp, dim, type, seed, nullptr);
}
// same draws, lattice passes on the device (pdgpu_grains_voronoi / pdgpu_grains_grow_precip)
int GrainStructure::generate_device(pdgpu_ctx* ctx, const PdConfig& cfg, const GrainParams& gp, int dim,
                                    const uint8_t* type, int seed) {
    return generate_impl(cfg, gp, dim, type, seed, ctx);
}

int GrainStructure::generate_impl(const PdConfig& cfg, const GrainParams& gp, int dim, const uint8_t* type, int seed,
                                  pdgpu_ctx* ctx) {
    Lattice L;
    L.dim = dim;
    L.dx = cfg.dx;
    pdgpu_grid_extents(&cfg, dim, &L.Nx, &L.Ny, &L.Nz, L.o);
    const int N = L.Nx * L.Ny * L.Nz;
    grain_id.assign(N, -1);
    is_grain_boundary.assign(N, 0);
    is_precipitate.assign(N, 0);
    n_grains = 0;

    std::vector<int> solid;
    for (int n = 0; n < N; ++n)
        if (type[n] == SOLID) solid.push_back(n);
    if (solid.empty()) return 0;                                             // :25-29

    double cell = std::pow(cfg.dx, dim);
    double grain_vol = dim == 2 ? PI / 4.0 * gp.grain_size_mean * gp.grain_size_mean
                                : PI / 6.0 * gp.grain_size_mean * gp.grain_size_mean * gp.grain_size_mean;
    n_grains = std::max(1, (int)std::round(solid.size() * cell / grain_vol));    // :31-40

    std::mt19937 rng(seed);                                                  // :46-53
    std::uniform_int_distribution<int> pick(0, (int)solid.size() - 1);
    std::vector<double> seeds(3 * (size_t)n_grains);
    for (int g = 0; g < n_grains; ++g) L.pos(solid[pick(rng)], &seeds[3 * (size_t)g]);

    if (ctx) {   // Voronoi, boundaries and dilation on the device
        if (pdgpu_grains_voronoi(ctx, seeds.data(), n_grains, gp.gb_width_cells, grain_id.data(),
                                 is_grain_boundary.data()) != 0)
            return 1;
    } else {
    for (int n : solid) {                                                    // Voronoi :56-70
        double p[3];
        L.pos(n, p);
        double best = std::numeric_limits<double>::max();
        int best_g = 0;
        for (int g = 0; g < n_grains; ++g) {
            double d = L.dist(p, &seeds[3 * (size_t)g]);
            if (d < best) { best = d; best_g = g; }
        }
        grain_id[n] = best_g;
    }
    for (int n : solid) {                                                    // boundaries :75-89
        int gi = grain_id[n];
        L.for_immediate(n, type, [&](int nn) {
            if (type[nn] == SOLID && grain_id[nn] != gi) { is_grain_boundary[n] = 1; return true; }
            return false;
        });
    }
    for (int pass = 0; pass < gp.gb_width_cells; ++pass) {                   // dilation :92-107
        std::vector<uint8_t> next = is_grain_boundary;
        for (int n : solid) {
            if (is_grain_boundary[n]) continue;
            L.for_immediate(n, type, [&](int nn) {
                if (is_grain_boundary[nn]) { next[n] = 1; return true; }
                return false;
            });
        }
        is_grain_boundary.swap(next);
    }
    }
    if (gp.precip_fraction > 0.0) {                                          // precipitates :119-175
        std::vector<int> interior;
        for (int n : solid)
            if (!is_grain_boundary[n]) interior.push_back(n);
        double per_cluster = 1.0;
        if (gp.precip_cluster_cells > 0) {
            double r = gp.precip_cluster_cells;
            per_cluster = dim == 2 ? PI * r * r : (4.0 / 3.0) * PI * r * r * r;
        }
        int n_seeds = (int)(interior.size() * gp.precip_fraction / per_cluster);
        n_seeds = std::max(1, n_seeds);
        std::shuffle(interior.begin(), interior.end(), rng);
        n_seeds = std::min(n_seeds, (int)interior.size());
        for (int s = 0; s < n_seeds; ++s) is_precipitate[interior[s]] = 1;
        if (gp.precip_cluster_cells > 0 && ctx) {
            std::vector<uint8_t> seeds_flag = is_precipitate;
            if (pdgpu_grains_grow_precip(ctx, is_grain_boundary.data(), seeds_flag.data(), gp.precip_cluster_cells,
                                         is_precipitate.data()) != 0)
                return 1;
        } else if (gp.precip_cluster_cells > 0) {
            double cluster_r = gp.precip_cluster_cells * cfg.dx;
            std::vector<uint8_t> grown = is_precipitate;
            for (int n : solid) {
                if (is_grain_boundary[n] || is_precipitate[n]) continue;
                double p[3], q[3];
                L.pos(n, p);
                for (int s = 0; s < n_seeds; ++s) {
                    L.pos(interior[s], q);
                    if (L.dist(p, q) <= cluster_r) { grown[n] = 1; break; }
                }
            }
            is_precipitate.swap(grown);
        }
    }
    return 0;
}

extern "C" int pdhost_generate_grains(const PdConfig* cfg, double grain_size_mean, double precip_fraction,
                                      int gb_width_cells, int precip_cluster_cells, int dim,
                                      const uint8_t* node_type, int seed, int* grain_id, uint8_t* is_gb,
                                      uint8_t* is_precip, int* n_grains) {
    if (!cfg || !node_type || (dim != 2 && dim != 3)) return 1;
    GrainParams gp;
    gp.grain_size_mean = grain_size_mean; gp.precip_fraction = precip_fraction;
    gp.gb_width_cells = gb_width_cells; gp.precip_cluster_cells = precip_cluster_cells;
    GrainStructure gs;
    gs.generate(*cfg, gp, dim, node_type, seed);
    return pdhost_copy_out(gs, grain_id, is_gb, is_precip, n_grains);
}

extern "C" int pdhost_generate_grains_device(pdgpu_ctx* ctx, const PdConfig* cfg, double grain_size_mean,
                                             double precip_fraction, int gb_width_cells, int precip_cluster_cells,
                                             int dim, const uint8_t* node_type, int seed, int* grain_id,
                                             uint8_t* is_gb, uint8_t* is_precip, int* n_grains) {
    if (!ctx || !cfg || !node_type || (dim != 2 && dim != 3)) return 1;
    GrainParams gp;
    gp.grain_size_mean = grain_size_mean; gp.precip_fraction = precip_fraction;
    gp.gb_width_cells = gb_width_cells; gp.precip_cluster_cells = precip_cluster_cells;
    GrainStructure gs;
    if (gs.generate_device(ctx, *cfg, gp, dim, node_type, seed) != 0) return 2;
    return pdhost_copy_out(gs, grain_id, is_gb, is_precip, n_grains);
}

int pdhost_copy_out(const GrainStructure& gs, int* grain_id, uint8_t* is_gb, uint8_t* is_precip, int* n_grains) {
    size_t N = gs.grain_id.size();
    for (size_t n = 0; n < N; ++n) {
        if (grain_id) grain_id[n] = gs.grain_id[n];
        if (is_gb) is_gb[n] = gs.is_grain_boundary[n];
        if (is_precip) is_precip[n] = gs.is_precipitate[n];
    }
    if (n_grains) *n_grains = gs.n_grains;
    return 0;
}
