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

// ---- point cloud (two-level AMR grid, csrc/amr.cu) --------------------------------------------------------
// The reference's generator is written against grid.pos and the CSR (src/grains.cpp:16-179), so on the AMR
// cloud it runs unchanged; here the same passes take the arrays pdamr_get returns (2D): positions, node types,
// CSR offsets / indices / distances.  Same libstdc++ RNG calls in the same order.
extern "C" int pdhost_generate_grains_cloud(const PdConfig* cfg, double grain_size_mean, double precip_fraction,
                                            int gb_width_cells, int precip_cluster_cells, int N, const double* pos,
                                            const uint8_t* type, const int* nbr_off, const int* nbr_idx,
                                            const double* nbr_dist, int seed, int* grain_id, uint8_t* is_gb,
                                            uint8_t* is_precip, int* n_grains_out) {
    if (!cfg || !pos || !type || !nbr_off || !nbr_idx || !nbr_dist || !grain_id || !is_gb || !is_precip) return 1;
    for (int n = 0; n < N; ++n) { grain_id[n] = -1; is_gb[n] = 0; is_precip[n] = 0; }
    if (n_grains_out) *n_grains_out = 0;
    std::vector<int> solid;
    for (int n = 0; n < N; ++n)
        if (type[n] == SOLID) solid.push_back(n);
    if (solid.empty()) return 0;
    auto dist = [&](int a, const double* b) {           // norm(p - q), src/utils.h:16-24 (fma chain, sqrt)
        double s = 0.0;
        for (int d = 0; d < 2; ++d) {
            const double t = pos[2 * a + d] - b[d];
            s = std::fma(t, t, s);
        }
        return std::sqrt(s);
    };
    const double cell = std::pow(cfg->dx, 2);
    const double grain_area = PI / 4.0 * grain_size_mean * grain_size_mean;
    const int n_grains = std::max(1, (int)std::round(solid.size() * cell / grain_area));
    if (n_grains_out) *n_grains_out = n_grains;
    std::mt19937 rng(seed);
    std::uniform_int_distribution<int> pick(0, (int)solid.size() - 1);
    std::vector<double> seeds(2 * (size_t)n_grains);
    for (int g = 0; g < n_grains; ++g) {
        const int si = solid[pick(rng)];
        seeds[2 * g] = pos[2 * si]; seeds[2 * g + 1] = pos[2 * si + 1];
    }
    for (int ni : solid) {                               // Voronoi, first strictly smaller distance wins
        double best = std::numeric_limits<double>::max();
        int bg = 0;
        for (int g = 0; g < n_grains; ++g) {
            const double d = dist(ni, &seeds[2 * g]);
            if (d < best) { best = d; bg = g; }
        }
        grain_id[ni] = bg;
    }
    const double gb_cutoff = std::sqrt(2.0) * cfg->dx * 1.01;
    for (int ni : solid)
        for (int q = nbr_off[ni]; q < nbr_off[ni + 1]; ++q) {
            const int nj = nbr_idx[q];
            if (nbr_dist[q] > gb_cutoff) continue;
            if (type[nj] == SOLID && grain_id[nj] != grain_id[ni]) { is_gb[ni] = 1; break; }
        }
    for (int pass = 0; pass < gb_width_cells; ++pass) {
        std::vector<uint8_t> nw(is_gb, is_gb + N);
        for (int ni : solid) {
            if (is_gb[ni]) continue;
            for (int q = nbr_off[ni]; q < nbr_off[ni + 1]; ++q) {
                if (nbr_dist[q] > gb_cutoff) continue;
                if (is_gb[nbr_idx[q]]) { nw[ni] = 1; break; }
            }
        }
        std::copy(nw.begin(), nw.end(), is_gb);
    }
    if (precip_fraction > 0.0) {
        std::vector<int> interior;
        for (int ni : solid)
            if (!is_gb[ni]) interior.push_back(ni);
        double cells_per_cluster = 1.0;
        if (precip_cluster_cells > 0) {
            const double r = precip_cluster_cells;
            cells_per_cluster = PI * r * r;
        }
        int n_seeds = (int)(interior.size() * precip_fraction / cells_per_cluster);
        n_seeds = std::max(1, n_seeds);
        std::shuffle(interior.begin(), interior.end(), rng);
        n_seeds = std::min(n_seeds, (int)interior.size());
        for (int p = 0; p < n_seeds; ++p) is_precip[interior[p]] = 1;
        if (precip_cluster_cells > 0) {
            const double cluster_r = precip_cluster_cells * cfg->dx;
            std::vector<uint8_t> seed_mark(is_precip, is_precip + N);
            for (int ni : solid) {
                if (is_gb[ni] || seed_mark[ni]) continue;
                for (int p = 0; p < n_seeds; ++p) {
                    const int si = interior[p];
                    const double q[2] = {pos[2 * si], pos[2 * si + 1]};
                    if (dist(ni, q) <= cluster_r) { is_precip[ni] = 1; break; }
                }
            }
        }
    }
    return 0;
}
