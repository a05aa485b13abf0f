// amr.cu -- two-level static refinement: the reference's AMR grid, its cell-list neighbour
// search and the explicit flow / transport operators on that point cloud.  2D only, like the reference
// (its fine-zone test, auxiliary bands and cell hash are 2D: src/grid.cpp:343-350, 466-479).
//
//   pdamr_build             Grid::build_amr (src/grid.cpp:352-655): fine nodes at dx inside the fine zone around
//                           the wire, coarse nodes at amr_ratio dx elsewhere, FICTITIOUS nodes of either spacing in
//                           the bands across the zone boundary, each with inverse-distance (1/d^4) weights over the
//                           REAL nodes of the other level within that level's horizon.
//   pdamr_build_neighbors   Grid::build_neighbors_celllist (:660-796): same-level neighbours within
//                           delta_i + dx_j / 2, partial-volume correction, CSR in cell-traversal order.
//   both are host code, as in the reference (one-off set-up of a 10^4..10^5 node cloud); node order, source
//   order and CSR order reproduce the reference's loops, the floating point its Release build (explicit fma()
//   where g++ -O3 -march=native contracts, see geom.cuh), so every array is bit-identical
//   (tests/test_amr.py, CPU).
//
// Device side (generic one-thread-per-row CSR kernels; the cloud breaks the uniform-stencil fast paths):
//   update_fictitious (:802-842), the BCs over CSR rows (src/boundary.cpp:31-131, 143-321, 381-390; the AMR wall
//   mirror searches the wall node's own row for the node nearest to the mirror point, :185-201), PD_NS_Solver::step
//   with per-node V_H / beta (src/pd_ns.cpp:18-33, 78-180), solve_steady with the IDW update after every swap
//   (:328), PD_ARD_Solver::step with FICTITIOUS neighbours counted as fluid (src/pd_ard.cpp:17-31, 55-191),
//   apply_phase_change (:193-212).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace {
constexpr uint8_t T_FLUID = 0, T_SOLID = 1, T_WALL = 2, T_INLET = 3, T_OUTLET = 4, T_OUTSIDE = 5, T_FICT = 6;
constexpr double kPi = 3.14159265358979323846;
}   // namespace

struct AmrImplicit;                     // implicit branch on the cloud: amr_implicit.cuh, included at the end of this file
struct pdamr_ctx {
    AmrImplicit* imp = nullptr;
    PdConfig cfg;
    int ratio = 3;
    double buffer = 0.0, dx_c = 0.0, delta_c = 0.0;
    int N = 0, n_fine = 0, n_coarse = 0, n_fict = 0;
    double origin[2] = {0.0, 0.0};
    std::vector<double> pos;            // [N][2]
    std::vector<uint8_t> type;
    std::vector<double> dxl, deltal;
    std::vector<int> level;
    std::vector<int> fict_off, fict_src;
    std::vector<double> fict_w;
    std::vector<int> nbr_off, nbr_idx;
    std::vector<double> nbr_dist, nbr_evec, nbr_vol;    // evec [nnz][2]
    std::vector<int> mirror;            // per node: WALL -> mirror node or -1; others -2
    std::vector<int> out_nodes, out_level_off;          // OUTLET nodes by Gauss-Seidel level
    std::vector<int> out_list, out_ord_level, out_eoff, out_eidx;   // ascending OUTLET nodes, their levels, earlier OUTLET neighbours (ordinals)
    bool built = false, nbrs = false;

    // device
    int device = -1;
    cudaStream_t stream = nullptr;
    uint8_t *d_type = nullptr, *d_phase = nullptr, *d_gb = nullptr, *d_precip = nullptr, *d_salt = nullptr;
    int *d_off = nullptr, *d_idx = nullptr, *d_foff = nullptr, *d_fsrc = nullptr, *d_mirror = nullptr;
    int *d_out_nodes = nullptr, *d_out_level_off = nullptr, *d_int = nullptr;
    int *d_out_list = nullptr, *d_out_ord_level = nullptr, *d_out_eoff = nullptr, *d_out_eidx = nullptr, *d_out_cnt = nullptr;
    double *d_out_bv = nullptr, *d_out_bc = nullptr;
    double *d_dist = nullptr, *d_evec = nullptr, *d_vol = nullptr, *d_fw = nullptr, *d_pos = nullptr, *d_delta = nullptr;
    double *rho[2] = {nullptr, nullptr}, *vel[2] = {nullptr, nullptr}, *C[2] = {nullptr, nullptr}, *p = nullptr;
    double* d_red = nullptr;
    double* h_red = nullptr;
    int cur = 0, curC = 0;
    double volume_loss = 0.0;
    bool dev = false;
};

namespace {

inline bool in_zone(double x, double y, double r, double z0, double z1) { return std::fabs(x) <= r && y >= z0 && y <= z1; }

// classify_node (src/grid.cpp:303-340), 2D
uint8_t classify2d(const PdConfig& c, double px, double py, double dx_local) {
    const double radial = std::fabs(px), axial = py;
    const double wall_lim = std::fma(0.5, dx_local, std::fma((double)c.m_ratio, dx_local, c.R_tube));
    const double zmin = -c.L_upstream, zmax = c.L_wire + c.L_downstream;
    if (axial < zmin) {
        if (radial <= c.R_tube) return T_INLET;
        return radial <= wall_lim ? T_WALL : T_OUTSIDE;
    }
    if (axial > zmax) {
        if (radial <= c.R_tube) return T_OUTLET;
        return radial <= wall_lim ? T_WALL : T_OUTSIDE;
    }
    if (radial <= c.R_tube) {
        const bool wire = (std::fabs(px) <= c.R_wire) && (py >= 0.0) && (py <= c.L_wire);
        return wire ? T_SOLID : T_FLUID;
    }
    return radial <= wall_lim ? T_WALL : T_OUTSIDE;
}

// uniform cell grid over the cloud; cells hold node indices in ascending order
struct Cells {
    double x0, y0, h;
    int ncx = 0, ncy = 0;
    std::vector<std::vector<int>> list;
    std::vector<int>* at(int ix, int iy) {
        if (ix < 0 || iy < 0 || ix >= ncx || iy >= ncy) return nullptr;
        return &list[(size_t)iy * ncx + ix];
    }
};

}   // namespace

extern "C" int pdamr_create(const PdConfig* cfg, int amr_ratio, double amr_buffer, pdamr_ctx** out) {
    if (!cfg || !out) PD_FAIL("pdamr_create: null argument");
    if (amr_ratio < 1) PD_FAIL("pdamr_create: amr_ratio must be >= 1");
    if (!(cfg->dx > 0.0) || cfg->m_ratio < 1) PD_FAIL("pdamr_create: dx / m_ratio");
    pdamr_ctx* c = new pdamr_ctx();
    c->cfg = *cfg;
    c->ratio = amr_ratio;
    c->buffer = amr_buffer;
    c->dx_c = amr_ratio * cfg->dx;                     // Config::compute_derived (src/config.cpp:98-101)
    c->delta_c = cfg->m_ratio * c->dx_c;
    *out = c;
    return 0;
}

// Grid::build_amr (src/grid.cpp:352-655)
extern "C" int pdamr_build(pdamr_ctx* c) {
    if (!c) PD_FAIL("pdamr_build: null context");
    const PdConfig& k = c->cfg;
    const double dx_f = k.dx, dx_c = c->dx_c, delta_f = k.delta, delta_c = c->delta_c;
    const double m = (double)k.m_ratio;
    const double fine_r = k.R_wire + c->buffer, fine_z0 = -c->buffer, fine_z1 = k.L_wire + c->buffer;
    const double z_min = -std::fma(m, dx_c, k.L_upstream), z_max = std::fma(m, dx_c, k.L_wire + k.L_downstream);
    const double r_min = -std::fma(m, dx_c, k.R_tube), r_max = std::fma(m, dx_c, k.R_tube);
    c->pos.clear(); c->type.clear(); c->dxl.clear(); c->deltal.clear(); c->level.clear();
    auto add = [&](double px, double py, uint8_t t, int lvl) {
        c->pos.push_back(px); c->pos.push_back(py);
        c->type.push_back(t);
        c->dxl.push_back(lvl == 0 ? dx_f : dx_c);
        c->deltal.push_back(lvl == 0 ? delta_f : delta_c);
        c->level.push_back(lvl);
    };
    const int nx_f = (int)std::round((r_max - r_min) / dx_f) + 1, ny_f = (int)std::round((z_max - z_min) / dx_f) + 1;
    const int nx_c = (int)std::round((r_max - r_min) / dx_c) + 1, ny_c = (int)std::round((z_max - z_min) / dx_c) + 1;
    // 1. fine nodes inside the fine zone, 2. coarse nodes outside it (:397-452)
    c->n_fine = 0;
    for (int jj = 0; jj < ny_f; ++jj) {
        const double py = std::fma((double)jj, dx_f, z_min);
        for (int ii = 0; ii < nx_f; ++ii) {
            const double px = std::fma((double)ii, dx_f, r_min);
            if (!in_zone(px, py, fine_r, fine_z0, fine_z1)) continue;
            const uint8_t t = classify2d(k, px, py, dx_f);
            if (t == T_OUTSIDE) continue;
            add(px, py, t, 0);
            c->n_fine++;
        }
    }
    c->n_coarse = 0;
    for (int jj = 0; jj < ny_c; ++jj) {
        const double py = std::fma((double)jj, dx_c, z_min);
        for (int ii = 0; ii < nx_c; ++ii) {
            const double px = std::fma((double)ii, dx_c, r_min);
            if (in_zone(px, py, fine_r, fine_z0, fine_z1)) continue;
            const uint8_t t = classify2d(k, px, py, dx_c);
            if (t == T_OUTSIDE) continue;
            add(px, py, t, 1);
            c->n_coarse++;
        }
    }
    const int n_real = c->n_fine + c->n_coarse;
    // 3. auxiliary nodes: IDW sources through a cell grid of size max(delta_f, delta_c) (:459-497)
    Cells cg;
    cg.h = std::max(delta_f, delta_c); cg.x0 = r_min; cg.y0 = z_min;
    int ix_max = 0, iy_max = 0;
    std::vector<int> cix(n_real), ciy(n_real);
    for (int i = 0; i < n_real; ++i) {
        cix[i] = (int)std::floor((c->pos[2 * i] - r_min) / cg.h);
        ciy[i] = (int)std::floor((c->pos[2 * i + 1] - z_min) / cg.h);
        ix_max = std::max(ix_max, cix[i]); iy_max = std::max(iy_max, ciy[i]);
    }
    cg.ncx = ix_max + 1; cg.ncy = iy_max + 1;
    cg.list.assign((size_t)cg.ncx * cg.ncy, {});
    for (int i = 0; i < n_real; ++i)
        if (cix[i] >= 0 && ciy[i] >= 0) cg.list[(size_t)ciy[i] * cg.ncx + cix[i]].push_back(i);
    auto in_radius = [&](double cx, double cy, double radius, int lvl, std::vector<int>& res) {
        res.clear();
        const int cr = (int)std::ceil(radius / cg.h) + 1;
        const int ax = (int)std::floor((cx - r_min) / cg.h), ay = (int)std::floor((cy - z_min) / cg.h);
        for (int dy = -cr; dy <= cr; ++dy)
            for (int ddx = -cr; ddx <= cr; ++ddx) {
                std::vector<int>* cell = cg.at(ax + ddx, ay + dy);
                if (!cell) continue;
                for (int idx : *cell) {
                    if (c->level[idx] != lvl) continue;
                    const double dr = c->pos[2 * idx] - cx, dz = c->pos[2 * idx + 1] - cy;
                    if (std::sqrt(std::fma(dr, dr, dz * dz)) <= radius) res.push_back(idx);
                }
            }
    };
    struct Fict { double x, y; int lvl; int first, count; };
    std::vector<Fict> fict;
    std::vector<int> fsrc, srcs;
    std::vector<double> fw;
    auto add_fict = [&](double px, double py, int lvl) {
        double W = 0.0;
        const int first = (int)fsrc.size();
        for (int s : srcs) {
            const double dr = c->pos[2 * s] - px, dz = c->pos[2 * s + 1] - py;
            double d2 = std::fma(dr, dr, dz * dz);
            if (d2 < 1e-30) d2 = 1e-30;
            const double w = 1.0 / (d2 * d2);
            fsrc.push_back(s); fw.push_back(w);
            W += w;
        }
        for (size_t q = first; q < fw.size(); ++q) fw[q] /= W;
        fict.push_back({px, py, lvl, first, (int)srcs.size()});
    };
    {   // auxiliary fine nodes outside the fine zone (:500-546)
        const double ar = fine_r + delta_f + dx_f, az0 = fine_z0 - delta_f - dx_f, az1 = fine_z1 + delta_f + dx_f;
        for (int jj = 0; jj < ny_f; ++jj) {
            const double py = std::fma((double)jj, dx_f, z_min);
            for (int ii = 0; ii < nx_f; ++ii) {
                const double px = std::fma((double)ii, dx_f, r_min);
                if (in_zone(px, py, fine_r, fine_z0, fine_z1)) continue;
                if (!in_zone(px, py, ar, az0, az1)) continue;
                if (classify2d(k, px, py, dx_f) == T_OUTSIDE) continue;
                in_radius(px, py, delta_c, 1, srcs);
                if (srcs.empty()) continue;
                add_fict(px, py, 0);
            }
        }
    }
    {   // auxiliary coarse nodes inside the fine zone, near its boundary (:550-596)
        const double ir = fine_r - delta_c - dx_c, iz0 = fine_z0 + delta_c + dx_c, iz1 = fine_z1 - delta_c - dx_c;
        for (int jj = 0; jj < ny_c; ++jj) {
            const double py = std::fma((double)jj, dx_c, z_min);
            for (int ii = 0; ii < nx_c; ++ii) {
                const double px = std::fma((double)ii, dx_c, r_min);
                if (!in_zone(px, py, fine_r, fine_z0, fine_z1)) continue;
                if (in_zone(px, py, ir, iz0, iz1)) continue;
                if (classify2d(k, px, py, dx_c) == T_OUTSIDE) continue;
                in_radius(px, py, delta_f, 0, srcs);
                if (srcs.empty()) continue;
                add_fict(px, py, 1);
            }
        }
    }
    c->n_fict = (int)fict.size();
    for (const Fict& f : fict) add(f.x, f.y, T_FICT, f.lvl);
    c->N = (int)c->type.size();
    c->fict_off.assign(c->N + 1, 0);
    for (int q = 0; q < c->n_fict; ++q) c->fict_off[n_real + q + 1] = fict[q].count;
    for (int i = 0; i < c->N; ++i) c->fict_off[i + 1] += c->fict_off[i];
    c->fict_src = fsrc;
    c->fict_w = fw;
    c->origin[0] = r_min; c->origin[1] = z_min;
    c->built = true;
    c->nbrs = false;
    return 0;
}

// wall-mirror table (src/boundary.cpp:143-264, AMR branch) and the Gauss-Seidel levels of the OUTLET nodes
static void amr_tables(pdamr_ctx* c) {
    const int N = c->N;
    c->mirror.assign(N, -2);
    for (int n = 0; n < N; ++n) {
        if (c->type[n] != T_WALL) continue;
        const double x = c->pos[2 * n], y = c->pos[2 * n + 1];
        int m = -1;
        bool geom = true;
        double xm = 0.0;
        if (x > c->cfg.R_tube) xm = 2.0 * c->cfg.R_tube - x;
        else if (x < -c->cfg.R_tube) xm = -2.0 * c->cfg.R_tube - x;
        else geom = false;
        if (geom) {
            double best = 1e30;
            for (int q = c->nbr_off[n]; q < c->nbr_off[n + 1]; ++q) {
                const int j = c->nbr_idx[q];
                const uint8_t t = c->type[j];
                if (t != T_FLUID && t != T_INLET && t != T_OUTLET && t != T_SOLID && t != T_FICT) continue;
                const double drx = c->pos[2 * j] - xm, dry = c->pos[2 * j + 1] - y;
                const double d2 = std::fma(drx, drx, dry * dry);
                if (d2 < best) { best = d2; m = j; }
            }
        }
        if (m < 0) {
            double best = 1e30;
            for (int q = c->nbr_off[n]; q < c->nbr_off[n + 1]; ++q) {
                const int j = c->nbr_idx[q];
                if (c->type[j] == T_FLUID && c->nbr_dist[q] < best) { best = c->nbr_dist[q]; m = j; }
            }
        }
        c->mirror[n] = m;
    }
    // apply_outlet_bc is an in-place sweep in index order (src/boundary.cpp:88-131): OUTLET node n sees the new
    // values of OUTLET neighbours with a smaller index; level(n) = 1 + max level of those
    std::vector<int> lvl(N, -1);
    int n_levels = 0;
    std::vector<int> outs;
    for (int n = 0; n < N; ++n) {
        if (c->type[n] != T_OUTLET) continue;
        int l = 0;
        for (int q = c->nbr_off[n]; q < c->nbr_off[n + 1]; ++q) {
            const int j = c->nbr_idx[q];
            if (j < n && c->type[j] == T_OUTLET) l = std::max(l, lvl[j] + 1);
        }
        lvl[n] = l;
        n_levels = std::max(n_levels, l + 1);
        outs.push_back(n);
    }
    c->out_level_off.assign(n_levels + 1, 0);
    for (int n : outs) c->out_level_off[lvl[n] + 1]++;
    for (int l = 0; l < n_levels; ++l) c->out_level_off[l + 1] += c->out_level_off[l];
    c->out_nodes.assign(outs.size(), 0);
    std::vector<int> fill(c->out_level_off.begin(), c->out_level_off.end() - 1);
    for (int n : outs) c->out_nodes[fill[lvl[n]]++] = n;
    // fast sweep: ordinals of the OUTLET nodes, per node the ordinals of its EARLIER outlet neighbours
    std::vector<int> ord(N, -1);
    for (size_t t = 0; t < outs.size(); ++t) ord[outs[t]] = (int)t;
    c->out_list = outs;
    c->out_ord_level.resize(outs.size());          // level of each OUTLET node, by ordinal
    for (size_t t = 0; t < outs.size(); ++t) c->out_ord_level[t] = lvl[outs[t]];
    c->out_eoff.assign(outs.size() + 1, 0);
    c->out_eidx.clear();
    for (size_t t = 0; t < outs.size(); ++t) {
        const int n = outs[t];
        for (int q = c->nbr_off[n]; q < c->nbr_off[n + 1]; ++q) {
            const int j = c->nbr_idx[q];
            if (j < n && c->type[j] == T_OUTLET) c->out_eidx.push_back(ord[j]);
        }
        c->out_eoff[t + 1] = (int)c->out_eidx.size();
    }
}

// Grid::build_neighbors_celllist (src/grid.cpp:660-796)
extern "C" int pdamr_build_neighbors(pdamr_ctx* c) {
    if (!c || !c->built) PD_FAIL("pdamr_build_neighbors: build the grid first");
    const int N = c->N;
    const double h = std::min(c->cfg.delta, c->delta_c) / 2.0;
    double xmin = 1e30, xmax = -1e30, ymin = 1e30, ymax = -1e30;
    for (int i = 0; i < N; ++i) {
        xmin = std::min(xmin, c->pos[2 * i]); xmax = std::max(xmax, c->pos[2 * i]);
        ymin = std::min(ymin, c->pos[2 * i + 1]); ymax = std::max(ymax, c->pos[2 * i + 1]);
    }
    Cells cg;
    cg.h = h; cg.x0 = xmin; cg.y0 = ymin;
    cg.ncx = (int)std::ceil((xmax - xmin) / h) + 1;
    cg.ncy = (int)std::ceil((ymax - ymin) / h) + 1;
    cg.list.assign((size_t)cg.ncx * cg.ncy, {});
    for (int i = 0; i < N; ++i) {
        if (c->type[i] == T_OUTSIDE) continue;
        int ix = (int)std::floor((c->pos[2 * i] - xmin) / h), iy = (int)std::floor((c->pos[2 * i + 1] - ymin) / h);
        ix = std::max(0, std::min(ix, cg.ncx - 1));
        iy = std::max(0, std::min(iy, cg.ncy - 1));
        cg.list[(size_t)iy * cg.ncx + ix].push_back(i);
    }
    c->nbr_off.assign(N + 1, 0);
    c->nbr_idx.clear(); c->nbr_dist.clear(); c->nbr_evec.clear(); c->nbr_vol.clear();
    for (int i = 0; i < N; ++i) {
        c->nbr_off[i] = (int)c->nbr_idx.size();
        if (c->type[i] == T_OUTSIDE) continue;
        const double px = c->pos[2 * i], py = c->pos[2 * i + 1], di = c->deltal[i];
        const int sr = (int)std::ceil(di / h) + 1;
        const int ax = (int)std::floor((px - xmin) / h), ay = (int)std::floor((py - ymin) / h);
        for (int dy = -sr; dy <= sr; ++dy)
            for (int ddx = -sr; ddx <= sr; ++ddx) {
                std::vector<int>* cell = cg.at(ax + ddx, ay + dy);
                if (!cell) continue;
                for (int j : *cell) {
                    if (j == i || c->level[j] != c->level[i]) continue;
                    const double dr = c->pos[2 * j] - px, dz = c->pos[2 * j + 1] - py;
                    const double r = std::sqrt(std::fma(dr, dr, dz * dz));
                    if (r < 1e-14) continue;
                    const double dxj = c->dxl[j];
                    const double hi = std::fma(0.5, dxj, di);
                    if (r > hi) continue;
                    const double beta = (r <= std::fma(-0.5, dxj, di)) ? 1.0 : (hi - r) / dxj;
                    const double inv_r = 1.0 / r;
                    c->nbr_idx.push_back(j);
                    c->nbr_dist.push_back(r);
                    c->nbr_evec.push_back(dr * inv_r);
                    c->nbr_evec.push_back(dz * inv_r);
                    c->nbr_vol.push_back(beta * dxj * dxj);
                }
            }
    }
    c->nbr_off[N] = (int)c->nbr_idx.size();
    c->nbrs = true;
    amr_tables(c);
    return 0;
}

extern "C" int pdamr_info(pdamr_ctx* c, PdAmrInfo* o) {
    if (!c || !o) PD_FAIL("pdamr_info: null argument");
    o->N_total = c->N; o->n_fine = c->n_fine; o->n_coarse = c->n_coarse; o->n_fict = c->n_fict;
    o->nnz = c->nbrs ? (long long)c->nbr_idx.size() : -1;
    o->n_fict_entries = (long long)c->fict_src.size();
    for (int t = 0; t < 7; ++t) o->counts[t] = 0;
    for (uint8_t t : c->type) o->counts[t]++;
    o->origin[0] = c->origin[0]; o->origin[1] = c->origin[1];
    o->dx_coarse = c->dx_c; o->delta_coarse = c->delta_c;
    return 0;
}

// host copies of the geometry arrays
extern "C" int pdamr_get(pdamr_ctx* c, const char* name, void* out) {
    if (!c || !name || !out) PD_FAIL("pdamr_get: null argument");
    const std::string n(name);
    auto cp = [&](const void* src, size_t bytes) { std::memcpy(out, src, bytes); return 0; };
    if (n == "pos") return cp(c->pos.data(), sizeof(double) * c->pos.size());
    if (n == "node_type") return cp(c->type.data(), c->type.size());
    if (n == "dx_local") return cp(c->dxl.data(), sizeof(double) * c->dxl.size());
    if (n == "delta_local") return cp(c->deltal.data(), sizeof(double) * c->deltal.size());
    if (n == "grid_level") return cp(c->level.data(), sizeof(int) * c->level.size());
    if (n == "fict_offset") return cp(c->fict_off.data(), sizeof(int) * c->fict_off.size());
    if (n == "fict_source") return cp(c->fict_src.data(), sizeof(int) * c->fict_src.size());
    if (n == "fict_weight") return cp(c->fict_w.data(), sizeof(double) * c->fict_w.size());
    if (n == "nbr_offset") return cp(c->nbr_off.data(), sizeof(int) * c->nbr_off.size());
    if (n == "nbr_index") return cp(c->nbr_idx.data(), sizeof(int) * c->nbr_idx.size());
    if (n == "nbr_dist") return cp(c->nbr_dist.data(), sizeof(double) * c->nbr_dist.size());
    if (n == "nbr_evec") return cp(c->nbr_evec.data(), sizeof(double) * c->nbr_evec.size());
    if (n == "nbr_vol") return cp(c->nbr_vol.data(), sizeof(double) * c->nbr_vol.size());
    if (n == "wall_mirror") return cp(c->mirror.data(), sizeof(int) * c->mirror.size());
    PD_FAIL("pdamr_get: unknown array '%s'", name);
}

// ------------------------------------------------------------------ device side ----------
namespace {

struct AmrDev {
    int N;
    const uint8_t* type;
    const int *off, *idx;
    const double *dist, *evec, *vol, *delta;
};
AmrDev dev_view(const pdamr_ctx* c) {
    AmrDev d;
    d.N = c->N; d.type = c->d_type; d.off = c->d_off; d.idx = c->d_idx;
    d.dist = c->d_dist; d.evec = c->d_evec; d.vol = c->d_vol; d.delta = c->d_delta;
    return d;
}

__device__ __forceinline__ double eos_p(double rho, double rho0, double gamma, double B) {
    return eos_pressure(rho, rho0, gamma, B);
}

// compute_pressure for all nodes (src/pd_ns.cpp:36-50)
__global__ void k_amr_pressure(int N, const double* __restrict__ rho, double* __restrict__ p, double rho0, double gamma,
                               double B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) p[i] = eos_p(rho[i], rho0, gamma, B);
}

// Grid::update_fictitious (src/grid.cpp:802-842)
__global__ void k_amr_fict(int N, const uint8_t* __restrict__ type, const int* __restrict__ foff,
                           const int* __restrict__ fsrc, const double* __restrict__ fw, double* C, double* rho,
                           double* p, double* vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || type[i] != T_FICT) return;
    double c = 0.0, r = 0.0, pv = 0.0, v0 = 0.0, v1 = 0.0;
    for (int q = foff[i]; q < foff[i + 1]; ++q) {
        const int j = fsrc[q];
        const double w = fw[q];
        c += w * C[j]; r += w * rho[j]; pv += w * p[j];
        v0 += w * vel[2 * j]; v1 += w * vel[2 * j + 1];
    }
    C[i] = c; rho[i] = r; p[i] = pv; vel[2 * i] = v0; vel[2 * i + 1] = v1;
}

// apply_inlet_bc (src/boundary.cpp:31-75)
__global__ void k_amr_inlet(AmrDev g, const double* __restrict__ pos, double* rho, double* vel, double* C, double R_tube,
                            double U_in, double rho_f, double C_in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.N || g.type[i] != T_INLET) return;
    const double px = pos[2 * i];
    double rr = (px * px) / (R_tube * R_tube);
    if (rr > 1.0) rr = 1.0;
    vel[2 * i] = 0.0;
    vel[2 * i + 1] = 1.5 * U_in * (1.0 - rr);
    double s = 0.0;
    int cnt = 0;
    for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
        const int j = g.idx[q];
        if (g.type[j] == T_FLUID) { s += rho[j]; ++cnt; }
    }
    rho[i] = cnt > 0 ? s / cnt : rho_f;
    C[i] = C_in;
}

// apply_outlet_bc (src/boundary.cpp:88-131): levels in order, nodes of a level in parallel (one CTA)
__global__ void __launch_bounds__(256) k_amr_outlet(AmrDev g, const int* __restrict__ nodes,
                                                    const int* __restrict__ level_off, int n_levels, double* rho,
                                                    double* vel, double* C, double rho_f, double U_in) {
    for (int l = 0; l < n_levels; ++l) {
        for (int t = level_off[l] + threadIdx.x; t < level_off[l + 1]; t += blockDim.x) {
            const int i = nodes[t];
            double sv = 0.0, sc = 0.0;
            int cnt = 0;
            for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
                const int j = g.idx[q];
                const uint8_t tj = g.type[j];
                if (tj == T_FLUID || tj == T_OUTLET) { sv += vel[2 * j + 1]; sc += C[j]; ++cnt; }
            }
            rho[i] = rho_f;
            vel[2 * i] = 0.0;
            if (cnt > 0) {
                const double inv_c = 1.0 / cnt;
                vel[2 * i + 1] = sv * inv_c;
                C[i] = sc / cnt;
            } else {
                vel[2 * i + 1] = U_in;
                C[i] = 0.0;
            }
        }
        __syncthreads();
    }
}

// The same sweep split like the lattice path (outlet.cu): a parallel pre-pass sums the FLUID neighbours and the
// LATER outlet neighbours (old values), the sequential part only adds the already swept EARLIER outlet
// neighbours, which live in shared memory.
__global__ void k_amr_outlet_prepass(AmrDev g, const int* __restrict__ list, int n_out, double* rho, double* vel,
                                     const double* C, double rho_f, double* __restrict__ bv, double* __restrict__ bc,
                                     int* __restrict__ cnt) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out) return;
    const int i = list[t];
    double sv = 0.0, sc = 0.0;
    int c = 0;
    for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
        const int j = g.idx[q];
        const uint8_t tj = g.type[j];
        if (tj == T_FLUID || (tj == T_OUTLET && j > i)) { sv += vel[2 * j + 1]; sc += C[j]; ++c; }
        else if (tj == T_OUTLET) ++c;
    }
    bv[t] = sv; bc[t] = sc; cnt[t] = c;
    rho[i] = rho_f;
    vel[2 * i] = 0.0;
}
// one thread per OUTLET node (n_out <= 1024): level, pre-pass sums and the list of earlier neighbours are loaded
// once; the level loop only touches shared memory
__global__ void __launch_bounds__(1024)
k_amr_outlet_sweep(const int* __restrict__ list, const int* __restrict__ level_of, int n_levels, int n_out,
                   const int* __restrict__ eoff, const int* __restrict__ eidx, int n_e, const double* __restrict__ bv,
                   const double* __restrict__ bc, const int* __restrict__ cnt, double* vel, double* C, double U_in) {
    extern __shared__ double sm[];
    double* nv = sm;
    double* ncc = sm + n_out;
    int* s_e = (int*)(sm + 2 * n_out);
    const int t = threadIdx.x;
    for (int e = t; e < n_e; e += blockDim.x) s_e[e] = eidx[e];
    int lvl = -1, e0 = 0, e1 = 0, n = 0, node = 0;
    double sv = 0.0, sc = 0.0;
    if (t < n_out) { lvl = level_of[t]; e0 = eoff[t]; e1 = eoff[t + 1]; n = cnt[t]; sv = bv[t]; sc = bc[t]; node = list[t]; }
    __syncthreads();
    for (int l = 0; l < n_levels; ++l) {
        if (lvl == l) {
            for (int e = e0; e < e1; ++e) { sv += nv[s_e[e]]; sc += ncc[s_e[e]]; }
            const double v = n > 0 ? sv * (1.0 / n) : U_in, cc = n > 0 ? sc / n : 0.0;
            nv[t] = v; ncc[t] = cc;
            vel[2 * node + 1] = v;
            C[node] = cc;
        }
        __syncthreads();
    }
}

// apply_wall_mirror_proper with the table (src/boundary.cpp:266-283)
__global__ void k_amr_wall(int N, const int* __restrict__ mirror, double* rho, double* vel, double rho_f) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int m = mirror[n];
    if (m == -2) return;
    if (m >= 0) {
        vel[2 * n] = -vel[2 * m]; vel[2 * n + 1] = -vel[2 * m + 1];
        rho[n] = rho[m];
    } else {
        vel[2 * n] = 0.0; vel[2 * n + 1] = 0.0;
        rho[n] = rho_f;
    }
}

// apply_wall_concentration_bc (src/boundary.cpp:302-321)
__global__ void k_amr_wall_conc(AmrDev g, double* C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.N || g.type[i] != T_WALL) return;
    double s = 0.0;
    int cnt = 0;
    for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
        const int j = g.idx[q];
        if (g.type[j] == T_FLUID) { s += C[j]; ++cnt; }
    }
    C[i] = cnt > 0 ? s / cnt : 0.0;
}

__global__ void k_amr_solid(int N, const uint8_t* __restrict__ type, double* vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && type[i] == T_SOLID) { vel[2 * i] = 0.0; vel[2 * i + 1] = 0.0; }
}

// PD_NS_Solver::step (src/pd_ns.cpp:86-179) with the per-node constants of :18-33
__global__ void __launch_bounds__(128)
k_amr_ns_step(AmrDev g, const double* __restrict__ rho, const double* __restrict__ p, const double* __restrict__ vel,
              double* __restrict__ rho_n, double* __restrict__ vel_n, double dt, double rho_f, double mu, double c0,
              double eta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.N) return;
    const double rho_i = rho[i], vi0 = vel[2 * i], vi1 = vel[2 * i + 1];
    if (g.type[i] != T_FLUID) {
        rho_n[i] = rho_i; vel_n[2 * i] = vi0; vel_n[2 * i + 1] = vi1;
        return;
    }
    const double d = g.delta[i];
    const double inv_VH = 1.0 / (kPi * d * d), beta_l = 4.0 / (kPi * d * d);
    const double dens_diff = beta_l * (eta * c0 * d);
    const double p_i = p[i];
    double mass_conv = 0.0, mass_diff = 0.0, mc0 = 0.0, mc1 = 0.0, mp0 = 0.0, mp1 = 0.0, mv0 = 0.0, mv1 = 0.0;
    for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
        const int j = g.idx[q];
        const double xi = g.dist[q], ex = g.evec[2 * q], ey = g.evec[2 * q + 1], Vj = g.vol[q];
        if (Vj < 1e-30) continue;
        const double inv_xi = 1.0 / xi, inv_xi2 = inv_xi * inv_xi;
        const double rho_j = rho[j], p_j = p[j], vj0 = vel[2 * j], vj1 = vel[2 * j + 1];
        const double dd = (rho_j * vj0 - rho_i * vi0) * ex + (rho_j * vj1 - rho_i * vi1) * ey;
        mass_conv += dd * inv_xi * Vj;
        mass_diff += (rho_j - rho_i) * inv_xi2 * Vj;
        const double c0v = (rho_j * vj0 * vj0 - rho_i * vi0 * vi0) * ex + (rho_j * vj0 * vj1 - rho_i * vi0 * vi1) * ey;
        const double c1v = (rho_j * vj1 * vj0 - rho_i * vi1 * vi0) * ex + (rho_j * vj1 * vj1 - rho_i * vi1 * vi1) * ey;
        mc0 += c0v * inv_xi * Vj; mc1 += c1v * inv_xi * Vj;
        const double dp = (p_j - p_i) * inv_xi * Vj;
        mp0 += dp * ex; mp1 += dp * ey;
        mv0 += (vj0 - vi0) * inv_xi2 * Vj; mv1 += (vj1 - vi1) * inv_xi2 * Vj;
    }
    const double c_div = 2.0 * inv_VH;   // alpha = DIM
    double rn = rho_i + dt * (-c_div * mass_conv + dens_diff * mass_diff);
    rn = fmin(fmax(rn, 0.5 * rho_f), 2.0 * rho_f);
    rho_n[i] = rn;
    const double s = dt / rho_i;
    vel_n[2 * i] = vi0 + s * (-c_div * mc0 - c_div * mp0 + mu * beta_l * mv0);
    vel_n[2 * i + 1] = vi1 + s * (-c_div * mc1 - c_div * mp1 + mu * beta_l * mv1);
}

// salt-layer pre-pass (src/pd_ard.cpp:61-73)
__global__ void k_amr_salt(AmrDev g, const double* __restrict__ C, double C_sat, uint8_t* __restrict__ salt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.N) return;
    uint8_t s = 0;
    if (g.type[i] == T_SOLID)
        for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
            const int j = g.idx[q];
            if (g.vol[q] < 1e-30) continue;
            if (g.type[j] == T_FLUID && C[j] >= C_sat) { s = 1; break; }
        }
    salt[i] = s;
}

struct ArdK { double D_liquid, D_grain, D_gb, D_precip, decay, alpha_art, dx; };

// PD_ARD_Solver::step (src/pd_ard.cpp:81-190)
__global__ void __launch_bounds__(128)
k_amr_ard_step(AmrDev g, ArdK k, const double* __restrict__ C, const double* __restrict__ vel,
               const uint8_t* __restrict__ is_gb, const uint8_t* __restrict__ is_precip, const uint8_t* __restrict__ salt,
               double* __restrict__ C_n, double dt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.N) return;
    const uint8_t ti = g.type[i];
    const double C_i = C[i];
    if (ti != T_FLUID && ti != T_SOLID) { C_n[i] = C_i; return; }
    const bool i_fl = ti == T_FLUID, i_so = ti == T_SOLID;
    const double d = g.delta[i];
    const double beta_i = 4.0 / (kPi * d * d), div_coeff = 2.0 / (kPi * d * d);
    const double vi0 = i_fl ? vel[2 * i] : 0.0, vi1 = i_fl ? vel[2 * i + 1] : 0.0;
    const double vi_mag = i_fl ? sqrt(vi0 * vi0 + vi1 * vi1) : 0.0;
    double diff = 0.0, adv = 0.0;
    for (int q = g.off[i]; q < g.off[i + 1]; ++q) {
        const int j = g.idx[q];
        const double xi = g.dist[q], Vj = g.vol[q];
        if (Vj < 1e-30) continue;
        const uint8_t tj = g.type[j];
        if (tj == T_WALL || tj == T_OUTSIDE) continue;
        const double C_j = C[j];
        const double inv_xi = 1.0 / xi, inv_xi2 = inv_xi * inv_xi;
        const bool j_fl = tj == T_FLUID || tj == T_INLET || tj == T_OUTLET || tj == T_FICT, j_so = tj == T_SOLID;
        if (i_so && j_so) continue;
        double D_avg = 0.0;
        if (i_fl && j_fl) D_avg = k.D_liquid;
        else {
            const int si = i_so ? i : j;
            if (!salt[si]) {
                double D_s = is_gb[si] ? k.D_gb : (is_precip[si] ? k.D_precip : k.D_grain);
                D_s *= k.decay;
                D_avg = 2.0 * k.D_liquid * D_s / (k.D_liquid + D_s + 1e-30);
            }
        }
        double D_art = 0.0;
        if (i_fl && j_fl) {
            const double vj0 = vel[2 * j], vj1 = vel[2 * j + 1];
            D_art = k.alpha_art * fmax(vi_mag, sqrt(vj0 * vj0 + vj1 * vj1)) * k.dx;
        }
        diff += beta_i * (D_avg + D_art) * (C_j - C_i) * inv_xi2 * Vj;
        if (i_fl && j_fl) adv += (C_j - C_i) * (vi0 * g.evec[2 * q] + vi1 * g.evec[2 * q + 1]) * inv_xi * Vj;
    }
    adv *= div_coeff;
    double cn = C_i + dt * (diff - adv);
    if (cn < 0.0) cn = 0.0;
    C_n[i] = cn;
}

// reductions: [0] max |v| over FLUID (bit pattern), convergence block of solve_steady
__global__ void k_amr_vmax(int N, const uint8_t* __restrict__ type, const double* __restrict__ vel,
                           unsigned long long* out) {
    double m = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        if (type[i] == T_FLUID) m = fmax(m, sqrt(vel[2 * i] * vel[2 * i] + vel[2 * i + 1] * vel[2 * i + 1]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// one CTA, fixed order: deterministic
__global__ void __launch_bounds__(1024) k_amr_residual(int N, const uint8_t* __restrict__ type,
                                                       const double* __restrict__ v, const double* __restrict__ vn,
                                                       const double* __restrict__ rn, double* out) {
    __shared__ double sh[6][32];
    double num = 0.0, den = 0.0, vmax = 0.0, rmin = 1e30, rmax = -1e30, nanf = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        if (type[i] != T_FLUID) continue;
        const double a0 = v[2 * i], a1 = v[2 * i + 1], b0 = vn[2 * i], b1 = vn[2 * i + 1], r = rn[i];
        if (isnan(b0) || isnan(r)) nanf = 1.0;
        num += (b0 - a0) * (b0 - a0) + (b1 - a1) * (b1 - a1);
        den += a0 * a0 + a1 * a1;
        vmax = fmax(vmax, sqrt(b0 * b0 + b1 * b1));
        rmin = fmin(rmin, r); rmax = fmax(rmax, r);
    }
    num = warp_sum(num); den = warp_sum(den); vmax = warp_max(vmax);
    rmin = warp_min(rmin); rmax = warp_max(rmax); nanf = warp_max(nanf);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[0][w] = num; sh[1][w] = den; sh[2][w] = vmax; sh[3][w] = rmin; sh[4][w] = rmax; sh[5][w] = nanf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 32; ++q) {
            num += sh[0][q]; den += sh[1][q]; vmax = fmax(vmax, sh[2][q]);
            rmin = fmin(rmin, sh[3][q]); rmax = fmax(rmax, sh[4][q]); nanf = fmax(nanf, sh[5][q]);
        }
        out[0] = num; out[1] = den; out[2] = vmax; out[3] = rmin; out[4] = rmax; out[5] = nanf;
    }
}

// apply_phase_change (src/pd_ard.cpp:193-212); the index order of the serial scan does not matter
__global__ void k_amr_phase(int N, uint8_t* type, uint8_t* phase, double* rho, double* vel, double* C, double C_thresh,
                            double rho_f, int* count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (phase[i] == 0 && type[i] == T_SOLID && C[i] < C_thresh) {
        phase[i] = 1; type[i] = T_FLUID;
        rho[i] = rho_f; vel[2 * i] = 0.0; vel[2 * i + 1] = 0.0; C[i] = C_thresh;
        atomicAdd(count, 1);
    }
}

template <typename T>
int up(T** d, const std::vector<T>& h) {
    if (*d) { CUDA_OK(cudaFree(*d)); *d = nullptr; }
    CUDA_OK(cudaMalloc(d, sizeof(T) * std::max<size_t>(h.size(), 1)));
    if (!h.empty()) CUDA_OK(cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return 0;
}

int upload_tables(pdamr_ctx* c) {
    PD_TRY(up(&c->d_type, c->type));
    PD_TRY(up(&c->d_mirror, c->mirror));
    PD_TRY(up(&c->d_out_nodes, c->out_nodes));
    PD_TRY(up(&c->d_out_level_off, c->out_level_off));
    PD_TRY(up(&c->d_out_list, c->out_list)); PD_TRY(up(&c->d_out_ord_level, c->out_ord_level));
    PD_TRY(up(&c->d_out_eoff, c->out_eoff)); PD_TRY(up(&c->d_out_eidx, c->out_eidx));
    const size_t no = std::max<size_t>(c->out_list.size(), 1);
    if (c->d_out_bv) { CUDA_OK(cudaFree(c->d_out_bv)); CUDA_OK(cudaFree(c->d_out_bc)); CUDA_OK(cudaFree(c->d_out_cnt)); }
    CUDA_OK(cudaMalloc(&c->d_out_bv, sizeof(double) * no));
    CUDA_OK(cudaMalloc(&c->d_out_bc, sizeof(double) * no));
    CUDA_OK(cudaMalloc(&c->d_out_cnt, sizeof(int) * no));
    return 0;
}

}   // namespace

#define AMR_DEV(c)                                                                              \
    do {                                                                                        \
        if (!(c) || !(c)->dev) PD_FAIL("AMR context has no device state (pdamr_device_init)");  \
        CUDA_OK(cudaSetDevice((c)->device));                                                    \
    } while (0)

static unsigned nb(int n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// uploads geometry and tables, allocates the double-buffered fields (Fields::allocate, src/fields.h:28-46)
extern "C" int pdamr_device_init(pdamr_ctx* c, int device) {
    if (!c || !c->nbrs) PD_FAIL("pdamr_device_init: build the grid and its neighbours first");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        PD_FAIL("pdamr_device_init: no CUDA device available (%s); libpdgpu has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) PD_FAIL("pdamr_device_init: device %d out of range", device);
    c->device = device;
    CUDA_OK(cudaSetDevice(device));
    if (!c->stream) CUDA_OK(cudaStreamCreate(&c->stream));
    PD_TRY(upload_tables(c));
    PD_TRY(up(&c->d_off, c->nbr_off)); PD_TRY(up(&c->d_idx, c->nbr_idx));
    PD_TRY(up(&c->d_dist, c->nbr_dist)); PD_TRY(up(&c->d_evec, c->nbr_evec)); PD_TRY(up(&c->d_vol, c->nbr_vol));
    PD_TRY(up(&c->d_foff, c->fict_off)); PD_TRY(up(&c->d_fsrc, c->fict_src)); PD_TRY(up(&c->d_fw, c->fict_w));
    PD_TRY(up(&c->d_pos, c->pos)); PD_TRY(up(&c->d_delta, c->deltal));
    const size_t N = (size_t)c->N;
    for (int b = 0; b < 2; ++b) {
        CUDA_OK(cudaMalloc(&c->rho[b], sizeof(double) * N)); CUDA_OK(cudaMemset(c->rho[b], 0, sizeof(double) * N));
        CUDA_OK(cudaMalloc(&c->vel[b], sizeof(double) * 2 * N)); CUDA_OK(cudaMemset(c->vel[b], 0, sizeof(double) * 2 * N));
        CUDA_OK(cudaMalloc(&c->C[b], sizeof(double) * N)); CUDA_OK(cudaMemset(c->C[b], 0, sizeof(double) * N));
    }
    CUDA_OK(cudaMalloc(&c->p, sizeof(double) * N)); CUDA_OK(cudaMemset(c->p, 0, sizeof(double) * N));
    for (uint8_t** q : {&c->d_phase, &c->d_gb, &c->d_precip, &c->d_salt}) {
        CUDA_OK(cudaMalloc(q, N)); CUDA_OK(cudaMemset(*q, 0, N));
    }
    CUDA_OK(cudaMalloc(&c->d_red, sizeof(double) * 16));
    CUDA_OK(cudaMalloc(&c->d_int, sizeof(int) * 4));
    CUDA_OK(cudaMallocHost(&c->h_red, sizeof(double) * 16));
    c->cur = 0; c->curC = 0;
    c->dev = true;
    return 0;
}

static void amr_implicit_free(pdamr_ctx* c);
static void amr_implicit_invalidate(pdamr_ctx* c);
static int amr_smooth_conc(pdamr_ctx* c);

extern "C" int pdamr_destroy(pdamr_ctx* c) {
    if (!c) return 0;
    if (c->dev) {
        cudaSetDevice(c->device);
        amr_implicit_free(c);
        for (void* q : {(void*)c->d_type, (void*)c->d_phase, (void*)c->d_gb, (void*)c->d_precip, (void*)c->d_salt,
                        (void*)c->d_off, (void*)c->d_idx, (void*)c->d_foff, (void*)c->d_fsrc, (void*)c->d_mirror,
                        (void*)c->d_out_nodes, (void*)c->d_out_level_off, (void*)c->d_int, (void*)c->d_out_list,
                        (void*)c->d_out_ord_level, (void*)c->d_out_eoff, (void*)c->d_out_eidx, (void*)c->d_out_cnt,
                        (void*)c->d_out_bv, (void*)c->d_out_bc, (void*)c->d_dist,
                        (void*)c->d_evec, (void*)c->d_vol, (void*)c->d_fw, (void*)c->d_pos, (void*)c->d_delta,
                        (void*)c->rho[0], (void*)c->rho[1], (void*)c->vel[0], (void*)c->vel[1], (void*)c->C[0],
                        (void*)c->C[1], (void*)c->p, (void*)c->d_red})
            cudaFree(q);
        cudaFreeHost(c->h_red);
        if (c->stream) cudaStreamDestroy(c->stream);
    }
    delete c;
    return 0;
}

// fields by the reference's names: rho, vel ([N][2]), pressure, C, rho_new, vel_new, C_new, phase, is_gb, is_precip
static int field_ptr(pdamr_ctx* c, const char* name, void** ptr, size_t* bytes) {
    const std::string n(name);
    const size_t N = (size_t)c->N;
    if (n == "rho") { *ptr = c->rho[c->cur]; *bytes = 8 * N; }
    else if (n == "rho_new") { *ptr = c->rho[1 - c->cur]; *bytes = 8 * N; }
    else if (n == "vel") { *ptr = c->vel[c->cur]; *bytes = 16 * N; }
    else if (n == "vel_new") { *ptr = c->vel[1 - c->cur]; *bytes = 16 * N; }
    else if (n == "C") { *ptr = c->C[c->curC]; *bytes = 8 * N; }
    else if (n == "C_new") { *ptr = c->C[1 - c->curC]; *bytes = 8 * N; }
    else if (n == "pressure") { *ptr = c->p; *bytes = 8 * N; }
    else if (n == "phase") { *ptr = c->d_phase; *bytes = N; }
    else if (n == "is_gb") { *ptr = c->d_gb; *bytes = N; }
    else if (n == "is_precip") { *ptr = c->d_precip; *bytes = N; }
    else if (n == "node_type") { *ptr = c->d_type; *bytes = N; }
    else PD_FAIL("pdamr: unknown field '%s'", name);
    return 0;
}
extern "C" int pdamr_field_set(pdamr_ctx* c, const char* name, const void* src) {
    AMR_DEV(c);
    void* p; size_t b;
    PD_TRY(field_ptr(c, name, &p, &b));
    if (std::string(name) == "node_type") PD_FAIL("pdamr_field_set: node types come from the grid build");
    CUDA_OK(cudaMemcpy(p, src, b, cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int pdamr_field_get(pdamr_ctx* c, const char* name, void* dst) {
    AMR_DEV(c);
    void* p; size_t b;
    PD_TRY(field_ptr(c, name, &p, &b));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaMemcpy(dst, p, b, cudaMemcpyDeviceToHost));
    return 0;
}

static int enqueue_fict(pdamr_ctx* c) {
    k_amr_fict<<<nb(c->N, 128), 128, 0, c->stream>>>(c->N, c->d_type, c->d_foff, c->d_fsrc, c->d_fw, c->C[c->curC],
                                                     c->rho[c->cur], c->p, c->vel[c->cur]);
    return 0;
}
extern "C" int pdamr_update_fictitious(pdamr_ctx* c) {
    AMR_DEV(c);
    PD_TRY(enqueue_fict(c));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

static int enqueue_bc(pdamr_ctx* c, int which, int buf) {   // 0 inlet, 1 outlet, 2 wall, 3 solid, 4 wall_conc
    AmrDev g = dev_view(c);
    const PdConfig& k = c->cfg;
    const int N = c->N;
    switch (which) {
        case 0: k_amr_inlet<<<nb(N, 128), 128, 0, c->stream>>>(g, c->d_pos, c->rho[buf], c->vel[buf], c->C[c->curC], k.R_tube,
                                                               k.U_in, k.rho_f, k.C_liquid_init); break;
        case 1: {
            const int n_out = (int)c->out_list.size(), n_levels = (int)c->out_level_off.size() - 1;
            if (!n_out) break;
            const int n_e = (int)c->out_eidx.size();
            const size_t smem = sizeof(double) * 2 * n_out + sizeof(int) * n_e;
            static const bool force_list = getenv("PDGPU_AMR_OUTLET_LIST") != nullptr;   // tests: the general kernel
            if (n_out <= 1024 && smem <= 48 * 1024 && !force_list) {
                k_amr_outlet_prepass<<<nb(n_out, 128), 128, 0, c->stream>>>(g, c->d_out_list, n_out, c->rho[buf], c->vel[buf],
                                                                            c->C[c->curC], k.rho_f, c->d_out_bv, c->d_out_bc,
                                                                            c->d_out_cnt);
                k_amr_outlet_sweep<<<1, 1024, smem, c->stream>>>(c->d_out_list, c->d_out_ord_level, n_levels, n_out,
                                                                 c->d_out_eoff, c->d_out_eidx, n_e, c->d_out_bv, c->d_out_bc,
                                                                 c->d_out_cnt, c->vel[buf], c->C[c->curC], k.U_in);
            } else {
                k_amr_outlet<<<1, 256, 0, c->stream>>>(g, c->d_out_nodes, c->d_out_level_off, n_levels, c->rho[buf],
                                                       c->vel[buf], c->C[c->curC], k.rho_f, k.U_in);
            }
            break;
        }
        case 2: k_amr_wall<<<nb(N, 128), 128, 0, c->stream>>>(N, c->d_mirror, c->rho[buf], c->vel[buf], k.rho_f); break;
        case 3: k_amr_solid<<<nb(N, 128), 128, 0, c->stream>>>(N, c->d_type, c->vel[buf]); break;
        default: k_amr_wall_conc<<<nb(N, 128), 128, 0, c->stream>>>(g, c->C[c->curC]); break;
    }
    return 0;
}
// which: 0 inlet, 1 outlet, 2 wall (current buffers), 3 solid surface, 4 wall concentration, 5 wall (new buffers),
// 6 smooth_boundary_concentration (implicit branch)
extern "C" int pdamr_bc(pdamr_ctx* c, int which) {
    AMR_DEV(c);
    if (which < 0 || which > 6) PD_FAIL("pdamr_bc: which must be 0..6");
    if (which == 6) PD_TRY(amr_smooth_conc(c));
    else PD_TRY(enqueue_bc(c, which == 5 ? 2 : which, which == 5 ? 1 - c->cur : c->cur));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int vmax_fluid(pdamr_ctx* c, double* v) {
    CUDA_OK(cudaMemsetAsync(c->d_red, 0, sizeof(double), c->stream));
    k_amr_vmax<<<64, 256, 0, c->stream>>>(c->N, c->d_type, c->vel[c->cur], (unsigned long long*)c->d_red);
    CUDA_OK(cudaMemcpyAsync(c->h_red, c->d_red, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    *v = c->h_red[0];
    return 0;
}

// PD_NS_Solver::compute_dt (src/pd_ns.cpp:52-76; the global dx and delta, as the reference)
extern "C" int pdamr_ns_compute_dt(pdamr_ctx* c, double* dt) {
    AMR_DEV(c);
    double v_max = 0.0;
    PD_TRY(vmax_fluid(c, &v_max));
    const PdConfig& k = c->cfg;
    const double dt_cfl = k.dx / (k.c0 + v_max + 1e-30);
    const double dt_visc = 0.25 * k.dx * k.dx / (k.mu_f / k.rho_f + 1e-30);
    const double dt_dens = 0.25 * k.dx * k.dx / (k.eta_density * k.c0 * k.delta + 1e-30);
    *dt = k.cfl_factor * std::min(dt_cfl, std::min(dt_visc, dt_dens));
    return 0;
}

static int enqueue_ns_step(pdamr_ctx* c, double dt) {
    const PdConfig& k = c->cfg;
    const double B = k.rho_f * k.c0 * k.c0 / k.gamma_eos;
    k_amr_pressure<<<nb(c->N, 256), 256, 0, c->stream>>>(c->N, c->rho[c->cur], c->p, k.rho_f, k.gamma_eos, B);
    k_amr_ns_step<<<nb(c->N, 128), 128, 0, c->stream>>>(dev_view(c), c->rho[c->cur], c->p, c->vel[c->cur],
                                                        c->rho[1 - c->cur], c->vel[1 - c->cur], dt, k.rho_f, k.mu_f, k.c0,
                                                        k.eta_density);
    return 0;
}
extern "C" int pdamr_ns_step(pdamr_ctx* c, double dt) {
    AMR_DEV(c);
    PD_TRY(enqueue_ns_step(c, dt));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

// one loop body of solve_steady (src/pd_ns.cpp:196-205) without swap
static int enqueue_ns_body(pdamr_ctx* c, double dt) {
    for (int w = 0; w < 4; ++w) PD_TRY(enqueue_bc(c, w, c->cur));
    PD_TRY(enqueue_ns_step(c, dt));
    PD_TRY(enqueue_bc(c, 2, 1 - c->cur));
    return 0;
}
// `iters` x { BCs, step, wall mirror of the new buffers, swap, update_fictitious } (:196-205, :325-328)
extern "C" int pdamr_ns_iterate(pdamr_ctx* c, int iters, double dt) {
    AMR_DEV(c);
    for (int it = 0; it < iters; ++it) {
        PD_TRY(enqueue_ns_body(c, dt));
        c->cur = 1 - c->cur;
        PD_TRY(enqueue_fict(c));
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

// PD_NS_Solver::solve_steady (src/pd_ns.cpp:182-372) on the AMR cloud
extern "C" int pdamr_ns_solve_steady(pdamr_ctx* c, PdSteadyResult* out, int verbose) {
    AMR_DEV(c);
    if (!out) PD_FAIL("pdamr_ns_solve_steady: null output");
    double dt = 0.0;
    PD_TRY(pdamr_ns_compute_dt(c, &dt));
    if (verbose) printf("\n--- Flow solver: solving to steady state ---\n  Initial dt = %.4e s\n", dt);
    double eps = 1.0, v_max = 0.0, rmin = 0.0, rmax = 0.0;
    int status = 1, iter;
    bool diverged = false;
    const int max_iters = c->cfg.flow_max_iters;
    for (iter = 1; iter <= max_iters; ++iter) {
        PD_TRY(enqueue_ns_body(c, dt));
        if (iter <= 10 || iter % 100 == 0) {
            k_amr_residual<<<1, 1024, 0, c->stream>>>(c->N, c->d_type, c->vel[c->cur], c->vel[1 - c->cur],
                                                      c->rho[1 - c->cur], c->d_red);
            CUDA_OK(cudaMemcpyAsync(c->h_red, c->d_red, sizeof(double) * 6, cudaMemcpyDeviceToHost, c->stream));
            CUDA_OK(cudaStreamSynchronize(c->stream));
            const double num = c->h_red[0], den = c->h_red[1];
            v_max = c->h_red[2]; rmin = c->h_red[3]; rmax = c->h_red[4];
            if (c->h_red[5] > 0.0) {
                if (verbose) printf("  Flow DIVERGED (NaN) at iter %d\n", iter);
                diverged = true; status = 2;
                break;
            }
            eps = den > 1e-30 ? std::sqrt(num / den) : std::sqrt(num);
            if (verbose && (iter <= 10 || iter % c->cfg.output_every_flow == 0))
                printf("  Flow iter %6d: eps=%.3e  v_max=%.4e  rho=[%.2f,%.2f]  dt=%.3e\n", iter, eps, v_max, rmin, rmax, dt);
            if (v_max > 100.0 * c->cfg.U_in) {
                if (verbose) printf("  Flow DIVERGED (v_max=%.2e >> U_in=%.2e) at iter %d\n", v_max, c->cfg.U_in, iter);
                diverged = true; status = 3;
                break;
            }
            if (eps < c->cfg.flow_conv_tol && iter > 100) {
                if (verbose) printf("  Flow converged at iter %d, eps=%.3e\n", iter, eps);
                status = 0;
                break;
            }
        }
        c->cur = 1 - c->cur;
        PD_TRY(enqueue_fict(c));
        if (iter % 200 == 0) PD_TRY(pdamr_ns_compute_dt(c, &dt));
    }
    if (!diverged && iter > max_iters && verbose) printf("  Flow did NOT converge after %d iters, eps=%.3e\n", max_iters, eps);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    out->iters = iter; out->status = status; out->eps = eps; out->dt = dt;
    out->v_max = v_max; out->rho_min = rmin; out->rho_max = rmax;
    out->poiseuille_l2 = -1.0; out->poiseuille_nodes = 0; out->pad = 0;
    return 0;
}

extern "C" int pdamr_ard_set_volume_loss(pdamr_ctx* c, double vl) {
    if (!c) PD_FAIL("null context");
    c->volume_loss = vl;
    return 0;
}

// PD_ARD_Solver::compute_dt (src/pd_ard.cpp:34-53)
extern "C" int pdamr_ard_compute_dt(pdamr_ctx* c, double* dt) {
    AMR_DEV(c);
    double v_max = 0.0;
    PD_TRY(vmax_fluid(c, &v_max));
    const PdConfig& k = c->cfg;
    const double D_max = std::max(k.D_liquid, std::max(k.D_grain, k.D_gb));
    const double dt_diff = 0.25 * k.dx * k.dx / (D_max + k.alpha_art_diff * v_max * k.dx + 1e-30);
    const double dt_adv = k.dx / (v_max + 1e-30);
    *dt = k.cfl_factor_corr * std::min(dt_diff, dt_adv);
    return 0;
}

static int enqueue_ard_step(pdamr_ctx* c, double dt) {
    const PdConfig& k = c->cfg;
    AmrDev g = dev_view(c);
    ArdK a;
    a.D_liquid = k.D_liquid; a.D_grain = k.D_grain; a.D_gb = k.D_gb; a.D_precip = k.D_precip;
    a.decay = k.corrosion_decay_l > 0.0 ? std::pow(10.0, -c->volume_loss / k.corrosion_decay_l) : 1.0;
    a.alpha_art = k.alpha_art_diff; a.dx = k.dx;
    k_amr_salt<<<nb(c->N, 128), 128, 0, c->stream>>>(g, c->C[c->curC], k.C_sat, c->d_salt);
    k_amr_ard_step<<<nb(c->N, 128), 128, 0, c->stream>>>(g, a, c->C[c->curC], c->vel[c->cur], c->d_gb, c->d_precip, c->d_salt,
                                                         c->C[1 - c->curC], dt);
    return 0;
}
extern "C" int pdamr_ard_step(pdamr_ctx* c, double dt) {
    AMR_DEV(c);
    PD_TRY(enqueue_ard_step(c, dt));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}
// `steps` x { inlet, outlet, wall_conc, step, swap C } (src/coupling.cpp:232-240)
extern "C" int pdamr_ard_iterate(pdamr_ctx* c, int steps, double dt) {
    AMR_DEV(c);
    for (int s = 0; s < steps; ++s) {
        PD_TRY(enqueue_bc(c, 0, c->cur));
        PD_TRY(enqueue_bc(c, 1, c->cur));
        PD_TRY(enqueue_bc(c, 4, c->cur));
        PD_TRY(enqueue_ard_step(c, dt));
        c->curC = 1 - c->curC;
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

// apply_phase_change + update_node_types_after_dissolution + table refresh (src/coupling.cpp:256-271; the
// cell-list CSR does not depend on FLUID / SOLID_MG, so the rebuild the reference runs there is a no-op)
extern "C" int pdamr_phase_change(pdamr_ctx* c, int* n_dissolved) {
    AMR_DEV(c);
    CUDA_OK(cudaMemsetAsync(c->d_int, 0, sizeof(int), c->stream));
    k_amr_phase<<<nb(c->N, 128), 128, 0, c->stream>>>(c->N, c->d_type, c->d_phase, c->rho[c->cur], c->vel[c->cur],
                                                      c->C[c->curC], c->cfg.C_thresh, c->cfg.rho_f, c->d_int);
    int n = 0;
    CUDA_OK(cudaMemcpyAsync(&n, c->d_int, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (n > 0) {
        CUDA_OK(cudaMemcpy(c->type.data(), c->d_type, c->type.size(), cudaMemcpyDeviceToHost));
        amr_tables(c);
        PD_TRY(upload_tables(c));
        amr_implicit_invalidate(c);
    }
    if (n_dissolved) *n_dissolved = n;
    return 0;
}

#include "amr_implicit.cuh"
