/* oracle/pd_oracle.c -- TEST INFRASTRUCTURE ONLY (see pd_oracle.h).
 *
 * Plain-C restatement of the reference's explicit PD hot path, neighbours enumerated
 * from the horizon-offset table in CSR order.  Every function cites the reference
 * lines it follows.  Built with -ffp-contract=off: the only places where the
 * reference's Release build depends on FMA contraction for a *discrete* result
 * (node classification) are written with explicit fma() in the pattern
 * the reference's object code uses (checked by disassembly of oracle/_ref and by
 * tests/test_oracle_vs_ref.py); field arithmetic is left un-fused and is compared with
 * the reference at 1e-13 relative.
 */
#include "pd_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PDO_PI 3.14159265358979323846

void pdo_set_threads(int n) { omp_set_num_threads(n); }

/* ------------------------------------------------------------------------- */
/* Grid::build  (src/grid.cpp:29-155)                                         */
/* ------------------------------------------------------------------------- */
static void grid_extents(const PdoConfig* c, int dim, PdoGrid* g) {
    double m = (double)c->m_ratio, dx = c->dx;
    /* src/grid.cpp:38-39; the Release build fuses m*dx into the add (vfnmsub/vfmadd). */
    double z_min = -fma(m, dx, c->L_upstream);
    double z_max = fma(m, dx, c->L_wire + c->L_downstream);
    double r_min = -fma(m, dx, c->R_tube);
    double r_max = fma(m, dx, c->R_tube);
    g->dim = dim;
    g->m = c->m_ratio;
    g->dx = dx;
    g->delta = c->delta;
    if (dim == 2) { /* :41-52 */
        g->Nx = (int)round((r_max - r_min) / dx) + 1;
        g->Ny = (int)round((z_max - z_min) / dx) + 1;
        g->Nz = 1;
        g->origin[0] = r_min; g->origin[1] = z_min; g->origin[2] = 0.0;
    } else { /* :53-65 */
        g->Nx = (int)round((r_max - r_min) / dx) + 1;
        g->Ny = g->Nx;
        g->Nz = (int)round((z_max - z_min) / dx) + 1;
        g->origin[0] = r_min; g->origin[1] = r_min; g->origin[2] = z_min;
    }
    g->N = (long long)g->Nx * g->Ny * g->Nz;
}

/* src/grid.cpp:88-147; pos = origin + idx*dx is a single fma in the reference build. */
static uint8_t classify(const PdoConfig* c, const PdoGrid* g, int i, int j, int k) {
    double dx = g->dx;
    double px = fma((double)i, dx, g->origin[0]);
    double py = fma((double)j, dx, g->origin[1]);
    double pz = (g->dim == 3) ? fma((double)k, dx, g->origin[2]) : 0.0;
    double axial = (g->dim == 2) ? py : pz;
    double r2 = 0.0, radial;
    if (g->dim == 2) radial = fabs(px);
    else { r2 = fma(px, px, py * py); radial = sqrt(r2); }
    double z_phys_min = -c->L_upstream;
    double z_phys_max = c->L_wire + c->L_downstream;
    double wall_lim = fma(0.5, dx, fma((double)g->m, dx, c->R_tube));
    if (axial < z_phys_min) {
        if (radial <= c->R_tube) return PDO_INLET;
        return radial <= wall_lim ? PDO_WALL : PDO_OUTSIDE;
    }
    if (axial > z_phys_max) {
        if (radial <= c->R_tube) return PDO_OUTLET;
        return radial <= wall_lim ? PDO_WALL : PDO_OUTSIDE;
    }
    if (radial <= c->R_tube) {
        int wire;
        if (g->dim == 2) wire = (fabs(px) <= c->R_wire) && (py >= 0.0) && (py <= c->L_wire);
        else wire = (r2 <= c->R_wire * c->R_wire) && (pz >= 0.0) && (pz <= c->L_wire);
        return wire ? PDO_SOLID : PDO_FLUID;
    }
    return radial <= wall_lim ? PDO_WALL : PDO_OUTSIDE;
}

/* Offset stencil of Grid::build_neighbors (src/grid.cpp:161-187) with beta (:274-288). */
static void build_stencil(PdoGrid* g) {
    int mext = g->m + 1, dim = g->dim, cap = 1;
    for (int d = 0; d < dim; ++d) cap *= (2 * mext + 1);
    g->off_d = (int*)malloc(sizeof(int) * 3 * cap);
    g->off_dist = (double*)malloc(sizeof(double) * cap);
    g->off_evec = (double*)malloc(sizeof(double) * dim * cap);
    g->off_vol = (double*)malloc(sizeof(double) * cap);
    double dx = g->dx, delta = g->delta;
    double dx_dim = 1.0;
    for (int d = 0; d < dim; ++d) dx_dim *= dx;
    int n = 0;
    int klo = dim == 3 ? -mext : 0, khi = dim == 3 ? mext : 0;
    for (int dk = klo; dk <= khi; ++dk)
        for (int dj = -mext; dj <= mext; ++dj)
            for (int di = -mext; di <= mext; ++di) {
                if (!di && !dj && !dk) continue;
                double r = sqrt((double)(di * di + dj * dj + dk * dk)) * dx;
                if (!(r <= delta + 0.5 * dx)) continue;
                g->off_d[3 * n] = di; g->off_d[3 * n + 1] = dj; g->off_d[3 * n + 2] = dk;
                g->off_dist[n] = r;
                g->off_evec[dim * n] = di * dx / r;
                g->off_evec[dim * n + 1] = dj * dx / r;
                if (dim == 3) g->off_evec[dim * n + 2] = dk * dx / r;
                double beta;
                if (r <= delta - 0.5 * dx) beta = 1.0;
                else if (r <= delta + 0.5 * dx) beta = (delta + 0.5 * dx - r) / dx;
                else beta = 0.0;
                g->off_vol[n] = beta * dx_dim;
                ++n;
            }
    g->n_off = n;
}

PdoGrid* pdo_grid_build(const PdoConfig* cfg, int dim) {
    PdoGrid* g = (PdoGrid*)calloc(1, sizeof(PdoGrid));
    grid_extents(cfg, dim, g);
    g->node_type = (uint8_t*)malloc((size_t)g->N);
    long long NxNy = (long long)g->Nx * g->Ny;
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < g->N; ++n) {
        int k = (int)(n / NxNy);
        int rem = (int)(n % NxNy);
        g->node_type[n] = classify(cfg, g, rem % g->Nx, rem / g->Nx, k);
    }
    build_stencil(g);
    g->wall_mirror = (int*)malloc(sizeof(int) * (size_t)g->N);
    pdo_wall_mirror_build(g, cfg);
    return g;
}

void pdo_grid_free(PdoGrid* g) {
    if (!g) return;
    free(g->node_type); free(g->off_d); free(g->off_dist); free(g->off_evec);
    free(g->off_vol); free(g->wall_mirror); free(g);
}

void pdo_grid_info(const PdoGrid* g, long long* out, double* origin) {
    out[0] = g->dim; out[1] = g->Nx; out[2] = g->Ny; out[3] = g->Nz;
    out[4] = g->m; out[5] = g->n_off; out[6] = g->N;
    memcpy(origin, g->origin, sizeof(double) * 3);
}
uint8_t* pdo_grid_types(PdoGrid* g) { return g->node_type; }
int* pdo_grid_off_d(PdoGrid* g) { return g->off_d; }
double* pdo_grid_off_dist(PdoGrid* g) { return g->off_dist; }
double* pdo_grid_off_evec(PdoGrid* g) { return g->off_evec; }
double* pdo_grid_off_vol(PdoGrid* g) { return g->off_vol; }
int* pdo_grid_wall_mirror(PdoGrid* g) { return g->wall_mirror; }

/* neighbour of node (i,j,k) through offset o: linear index, or -1 when the reference's
 * CSR has no such entry (outside the box or node_type == OUTSIDE, src/grid.cpp:218-224). */
static inline long long nbr(const PdoGrid* g, int i, int j, int k, int o) {
    int ni = i + g->off_d[3 * o], nj = j + g->off_d[3 * o + 1], nk = k + g->off_d[3 * o + 2];
    if (ni < 0 || ni >= g->Nx || nj < 0 || nj >= g->Ny || nk < 0 || nk >= g->Nz) return -1;
    long long nn = ((long long)nk * g->Ny + nj) * g->Nx + ni;
    return g->node_type[nn] == PDO_OUTSIDE ? -1 : nn;
}
static inline void ijk(const PdoGrid* g, long long n, int* i, int* j, int* k) {
    long long NxNy = (long long)g->Nx * g->Ny;
    *k = (int)(n / NxNy);
    int rem = (int)(n % NxNy);
    *j = rem / g->Nx;
    *i = rem % g->Nx;
}

/* ------------------------------------------------------------------------- */
/* CSR  (src/grid.cpp:190-291)                                                */
/* ------------------------------------------------------------------------- */
void pdo_csr_offsets(const PdoGrid* g, long long* offset) {
    offset[0] = 0;
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < g->N; ++n) {
        long long cnt = 0;
        if (g->node_type[n] != PDO_OUTSIDE) {
            int i, j, k;
            ijk(g, n, &i, &j, &k);
            for (int o = 0; o < g->n_off; ++o) cnt += nbr(g, i, j, k, o) >= 0;
        }
        offset[n + 1] = cnt;
    }
    for (long long n = 0; n < g->N; ++n) offset[n + 1] += offset[n]; /* :230-232 */
}

void pdo_csr_fill(const PdoGrid* g, const long long* offset, int* index, double* dist,
                  double* evec, double* vol) {
    int dim = g->dim;
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < g->N; ++n) {
        if (g->node_type[n] == PDO_OUTSIDE) continue;
        int i, j, k;
        ijk(g, n, &i, &j, &k);
        long long w = offset[n];
        for (int o = 0; o < g->n_off; ++o) {
            long long nn = nbr(g, i, j, k, o);
            if (nn < 0) continue;
            index[w] = (int)nn;
            dist[w] = g->off_dist[o];
            for (int d = 0; d < dim; ++d) evec[w * dim + d] = g->off_evec[o * dim + d];
            vol[w] = g->off_vol[o];
            ++w;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Wall mirror table  (src/boundary.cpp:143-264)                              */
/* ------------------------------------------------------------------------- */
static int mirror_ok(uint8_t t) {
    return t == PDO_FLUID || t == PDO_INLET || t == PDO_OUTLET || t == PDO_SOLID;
}

void pdo_wall_mirror_build(PdoGrid* g, const PdoConfig* cfg) {
    double R = cfg->R_tube, dx = g->dx;
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < g->N; ++n) {
        g->wall_mirror[n] = -1;
        if (g->node_type[n] != PDO_WALL) continue;
        int i, j, k;
        ijk(g, n, &i, &j, &k);
        double x = fma((double)i, dx, g->origin[0]);
        long long mirror = -1;
        if (g->dim == 2) { /* :158-185 */
            int have = 1;
            double xm = 0.0;
            if (x > R) xm = 2.0 * R - x;
            else if (x < -R) xm = -2.0 * R - x;
            else have = 0; /* goto fallback */
            if (have) {
                int im = (int)round((xm - g->origin[0]) / dx);
                if (im >= 0 && im < g->Nx) {
                    long long id = (long long)j * g->Nx + im;
                    if (mirror_ok(g->node_type[id])) mirror = id;
                }
            }
        } else { /* :203-229 */
            double y = fma((double)j, dx, g->origin[1]);
            double r = sqrt(fma(x, x, y * y));
            if (r > R && r > 1e-30) {
                double rm = 2.0 * R - r;
                double xm = x * rm / r, ym = y * rm / r;
                int im = (int)round((xm - g->origin[0]) / dx);
                int jm = (int)round((ym - g->origin[1]) / dx);
                if (im >= 0 && im < g->Nx && jm >= 0 && jm < g->Ny) {
                    long long id = ((long long)k * g->Ny + jm) * g->Nx + im;
                    if (mirror_ok(g->node_type[id])) mirror = id;
                }
            }
        }
        if (mirror < 0) { /* fallback :254-263: nearest FLUID neighbour, strict < */
            double best = 1e30;
            for (int o = 0; o < g->n_off; ++o) {
                long long nn = nbr(g, i, j, k, o);
                if (nn >= 0 && g->node_type[nn] == PDO_FLUID && g->off_dist[o] < best) {
                    best = g->off_dist[o];
                    mirror = nn;
                }
            }
        }
        g->wall_mirror[n] = (int)mirror;
    }
}

/* ------------------------------------------------------------------------- */
/* Boundary operators                                                         */
/* ------------------------------------------------------------------------- */
/* src/boundary.cpp:31-75 */
void pdo_inlet_bc(const PdoGrid* g, const PdoConfig* cfg, double* rho, double* vel, double* Cc) {
    int dim = g->dim;
    double R2 = cfg->R_tube * cfg->R_tube;
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < g->N; ++n) {
        if (g->node_type[n] != PDO_INLET) continue;
        int i, j, k;
        ijk(g, n, &i, &j, &k);
        double px = fma((double)i, g->dx, g->origin[0]);
        double v_axial;
        if (dim == 2) {
            double rr = (px * px) / R2;
            if (rr > 1.0) rr = 1.0;
            v_axial = 1.5 * cfg->U_in * (1.0 - rr);
        } else {
            double py = fma((double)j, g->dx, g->origin[1]);
            double rr = (px * px + py * py) / R2;
            if (rr > 1.0) rr = 1.0;
            v_axial = 2.0 * cfg->U_in * (1.0 - rr);
        }
        for (int d = 0; d < dim; ++d) vel[n * dim + d] = 0.0;
        vel[n * dim + (dim - 1)] = v_axial;
        double s = 0.0;
        int cnt = 0;
        for (int o = 0; o < g->n_off; ++o) {
            long long nn = nbr(g, i, j, k, o);
            if (nn >= 0 && g->node_type[nn] == PDO_FLUID) { s += rho[nn]; ++cnt; }
        }
        rho[n] = cnt > 0 ? s / cnt : cfg->rho_f;
        Cc[n] = cfg->C_liquid_init;
    }
}

/* src/boundary.cpp:88-131 -- in place, lexicographic (Gauss-Seidel): SERIAL on purpose. */
void pdo_outlet_bc(const PdoGrid* g, const PdoConfig* cfg, double* rho, double* vel, double* Cc) {
    int dim = g->dim, ax = dim - 1;
    double v[3];
    for (long long n = 0; n < g->N; ++n) {
        if (g->node_type[n] != PDO_OUTLET) continue;
        int i, j, k;
        ijk(g, n, &i, &j, &k);
        rho[n] = cfg->rho_f;
        v[0] = v[1] = v[2] = 0.0;
        double cs = 0.0;
        int cnt = 0;
        for (int o = 0; o < g->n_off; ++o) {
            long long nn = nbr(g, i, j, k, o);
            if (nn < 0) continue;
            uint8_t t = g->node_type[nn];
            if (t == PDO_FLUID || t == PDO_OUTLET) {
                for (int d = 0; d < dim; ++d) v[d] += vel[nn * dim + d];
                cs += Cc[nn];
                ++cnt;
            }
        }
        for (int d = 0; d < dim; ++d) vel[n * dim + d] = 0.0;
        if (cnt > 0) {
            double inv = 1.0 / cnt;
            vel[n * dim + ax] = v[ax] * inv;
            Cc[n] = cs / cnt;
        } else {
            vel[n * dim + ax] = cfg->U_in;
            Cc[n] = 0.0;
        }
    }
}

/* src/boundary.cpp:266-283 with the table from pdo_wall_mirror_build */
void pdo_wall_bc(const PdoGrid* g, const PdoConfig* cfg, double* rho, double* vel) {
    int dim = g->dim;
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < g->N; ++n) {
        if (g->node_type[n] != PDO_WALL) continue;
        int mi = g->wall_mirror[n];
        if (mi >= 0) {
            for (int d = 0; d < dim; ++d) vel[n * dim + d] = -vel[(long long)mi * dim + d];
            rho[n] = rho[mi];
        } else {
            for (int d = 0; d < dim; ++d) vel[n * dim + d] = 0.0;
            rho[n] = cfg->rho_f;
        }
    }
}

/* src/boundary.cpp:302-321 */
void pdo_wall_conc_bc(const PdoGrid* g, double* Cc) {
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < g->N; ++n) {
        if (g->node_type[n] != PDO_WALL) continue;
        int i, j, k;
        ijk(g, n, &i, &j, &k);
        double s = 0.0;
        int cnt = 0;
        for (int o = 0; o < g->n_off; ++o) {
            long long nn = nbr(g, i, j, k, o);
            if (nn >= 0 && g->node_type[nn] == PDO_FLUID) { s += Cc[nn]; ++cnt; }
        }
        Cc[n] = cnt > 0 ? s / cnt : 0.0;
    }
}

/* src/boundary.cpp:332-376: in place, sequential index order (what the reference produces when the
 * affected planes fall into one OpenMP chunk; the GPU kernel sweeps planes in the same order) */
void pdo_smooth_conc(const PdoGrid* g, const PdoConfig* cfg, double* Cc) {
    const int ax = g->dim == 2 ? 1 : 2;
    const double y_min = -cfg->L_upstream, y_max = cfg->L_wire + cfg->L_downstream, delta = cfg->delta;
    for (long long n = 0; n < g->N; ++n) {
        if (g->node_type[n] != PDO_FLUID) continue;
        int i, j, k;
        ijk(g, n, &i, &j, &k);
        const int a = ax == 1 ? j : k;
        const double y = fma((double)a, g->dx, g->origin[ax]);      /* src/grid.cpp:88-92, contracted */
        const int near_in = (y - y_min < delta), near_out = (y_max - y < delta);
        if (!near_in && !near_out) continue;
        double s = 0.0;
        int cnt = 0;
        for (int o = 0; o < g->n_off; ++o) {
            const int dax = g->off_d[3 * o + ax];
            if (!((near_out && dax < 0) || (near_in && dax > 0))) continue;   /* yj < y  <=>  axial offset < 0 */
            long long nn = nbr(g, i, j, k, o);
            if (nn >= 0 && g->node_type[nn] == PDO_FLUID) { s += Cc[nn]; ++cnt; }
        }
        if (cnt > 0) Cc[n] = s / cnt;
    }
}

/* src/boundary.cpp:381-390 */
void pdo_solid_bc(const PdoGrid* g, double* vel) {
    int dim = g->dim;
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < g->N; ++n)
        if (g->node_type[n] == PDO_SOLID)
            for (int d = 0; d < dim; ++d) vel[n * dim + d] = 0.0;
}

/* ------------------------------------------------------------------------- */
/* PD-NS                                                                      */
/* ------------------------------------------------------------------------- */
double pdo_max_fluid_speed(const PdoGrid* g, const double* vel) {
    int dim = g->dim;
    double vmax = 0.0;
#pragma omp parallel for reduction(max : vmax) schedule(static)
    for (long long n = 0; n < g->N; ++n) {
        if (g->node_type[n] != PDO_FLUID) continue;
        double s = 0.0;
        for (int d = 0; d < dim; ++d) s += vel[n * dim + d] * vel[n * dim + d];
        double v = sqrt(s);
        if (v > vmax) vmax = v;
    }
    return vmax;
}

/* src/pd_ns.cpp:52-76 */
double pdo_ns_compute_dt(const PdoGrid* g, const PdoConfig* cfg, const double* vel) {
    double v_max = pdo_max_fluid_speed(g, vel);
    double dx = cfg->dx;
    double dt_cfl = dx / (cfg->c0 + v_max + 1e-30);
    double nu = cfg->mu_f / cfg->rho_f;
    double dt_visc = 0.25 * dx * dx / (nu + 1e-30);
    double D_v = cfg->eta_density * cfg->c0 * cfg->delta;
    double dt_dens = 0.25 * dx * dx / (D_v + 1e-30);
    double mn = dt_cfl < dt_visc ? dt_cfl : dt_visc;
    if (dt_dens < mn) mn = dt_dens;
    return cfg->cfl_factor * mn;
}

/* src/pd_ns.cpp:36-50 (Tait EOS, every node) + :78-180 (bond sums + forward Euler) */
void pdo_ns_step(const PdoGrid* g, const PdoConfig* cfg, double dt, const double* rho,
                 const double* vel, double* pressure, double* rho_new, double* vel_new) {
    int dim = g->dim;
    double B = cfg->rho_f * cfg->c0 * cfg->c0 / cfg->gamma_eos;
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < g->N; ++n) {
        double ratio = rho[n] / cfg->rho_f;
        if (ratio < 0.5) ratio = 0.5;
        if (ratio > 2.0) ratio = 2.0;
        pressure[n] = B * (pow(ratio, cfg->gamma_eos) - 1.0);
    }
    double alpha = (double)dim, V_H, beta_lap; /* src/pd_ns.cpp:7-16 */
    if (dim == 2) {
        V_H = PDO_PI * cfg->delta * cfg->delta;
        beta_lap = 4.0 / (PDO_PI * cfg->delta * cfg->delta);
    } else {
        V_H = (4.0 / 3.0) * PDO_PI * cfg->delta * cfg->delta * cfg->delta;
        beta_lap = 12.0 / (PDO_PI * cfg->delta * cfg->delta);
    }
    double inv_VH = 1.0 / V_H;
    double D_v = cfg->eta_density * cfg->c0 * cfg->delta;
    double dens_diff_coeff = beta_lap * D_v;
    double mu = cfg->mu_f;

#pragma omp parallel for schedule(dynamic, 256)
    for (long long n = 0; n < g->N; ++n) {
        if (g->node_type[n] != PDO_FLUID) { /* :93-97 */
            rho_new[n] = rho[n];
            for (int d = 0; d < dim; ++d) vel_new[n * dim + d] = vel[n * dim + d];
            continue;
        }
        int i, j, k;
        ijk(g, n, &i, &j, &k);
        double rho_i = rho[n], p_i = pressure[n];
        double vi[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d) vi[d] = vel[n * dim + d];
        double mass_conv = 0.0, mass_diff = 0.0;
        double mom_conv[3] = {0, 0, 0}, mom_pres[3] = {0, 0, 0}, mom_visc[3] = {0, 0, 0};
        for (int o = 0; o < g->n_off; ++o) { /* :115-157 */
            long long nn = nbr(g, i, j, k, o);
            if (nn < 0) continue;
            double xi = g->off_dist[o], V_j = g->off_vol[o];
            const double* e = &g->off_evec[o * dim];
            if (V_j < 1e-30) continue;
            double rho_j = rho[nn], p_j = pressure[nn];
            double vj[3] = {0, 0, 0};
            for (int d = 0; d < dim; ++d) vj[d] = vel[nn * dim + d];
            double inv_xi = 1.0 / xi, inv_xi2 = inv_xi * inv_xi;
            double dd = 0.0;
            for (int d = 0; d < dim; ++d) dd += (rho_j * vj[d] - rho_i * vi[d]) * e[d];
            mass_conv += dd * inv_xi * V_j;
            mass_diff += dens_diff_coeff * (rho_j - rho_i) * inv_xi2 * V_j;
            for (int d = 0; d < dim; ++d) {
                double conv_d = 0.0;
                for (int dp = 0; dp < dim; ++dp)
                    conv_d += (rho_j * vj[d] * vj[dp] - rho_i * vi[d] * vi[dp]) * e[dp];
                mom_conv[d] += conv_d * inv_xi * V_j;
            }
            for (int d = 0; d < dim; ++d) mom_pres[d] += (p_j - p_i) * e[d] * inv_xi * V_j;
            for (int d = 0; d < dim; ++d) mom_visc[d] += (vj[d] - vi[d]) * inv_xi2 * V_j;
        }
        double rn = rho_i + dt * (-(alpha * inv_VH) * mass_conv + mass_diff); /* :160-168 */
        if (rn < 0.5 * cfg->rho_f) rn = 0.5 * cfg->rho_f;
        if (rn > 2.0 * cfg->rho_f) rn = 2.0 * cfg->rho_f;
        rho_new[n] = rn;
        double inv_rho = 1.0 / rho_i; /* :171-178 */
        for (int d = 0; d < dim; ++d)
            vel_new[n * dim + d] = vi[d] + dt * inv_rho * (-(alpha * inv_VH) * mom_conv[d]
                                                           - (alpha * inv_VH) * mom_pres[d]
                                                           + mu * beta_lap * mom_visc[d]);
    }
}

/* src/pd_ns.cpp:273-301 (serial, in index order) */
void pdo_ns_residual(const PdoGrid* g, const double* vel, const double* vel_new,
                     const double* rho_new, double* out) {
    int dim = g->dim;
    double num = 0.0, den = 0.0, vmax = 0.0, rmin = 1e30, rmax = -1e30, has_nan = 0.0;
    for (long long n = 0; n < g->N; ++n) {
        if (g->node_type[n] != PDO_FLUID) continue;
        if (isnan(vel_new[n * dim]) || isnan(rho_new[n])) { has_nan = 1.0; break; }
        double dv2 = 0.0, v2 = 0.0, vn2 = 0.0;
        for (int d = 0; d < dim; ++d) {
            double dv = vel_new[n * dim + d] - vel[n * dim + d];
            dv2 += dv * dv;
            v2 += vel[n * dim + d] * vel[n * dim + d];
            vn2 += vel_new[n * dim + d] * vel_new[n * dim + d];
        }
        num += dv2; den += v2;
        double vn = sqrt(vn2);
        if (vn > vmax) vmax = vn;
        if (rho_new[n] < rmin) rmin = rho_new[n];
        if (rho_new[n] > rmax) rmax = rho_new[n];
    }
    out[0] = num; out[1] = den; out[2] = vmax; out[3] = rmin; out[4] = rmax; out[5] = has_nan;
}

/* ------------------------------------------------------------------------- */
/* PD-ARD explicit                                                            */
/* ------------------------------------------------------------------------- */
/* src/pd_ard.cpp:34-53 */
double pdo_ard_compute_dt(const PdoGrid* g, const PdoConfig* cfg, const double* vel) {
    double D_max = cfg->D_liquid;
    if (cfg->D_grain > D_max) D_max = cfg->D_grain;
    if (cfg->D_gb > D_max) D_max = cfg->D_gb;
    double v_max = pdo_max_fluid_speed(g, vel);
    double D_eff = D_max + cfg->alpha_art_diff * v_max * cfg->dx;
    double dt_diff = 0.25 * cfg->dx * cfg->dx / (D_eff + 1e-30);
    double dt_adv = cfg->dx / (v_max + 1e-30);
    return cfg->cfl_factor_corr * (dt_diff < dt_adv ? dt_diff : dt_adv);
}

static inline double speed(const double* vel, long long n, int dim) {
    double s = 0.0;
    for (int d = 0; d < dim; ++d) s += vel[n * dim + d] * vel[n * dim + d];
    return sqrt(s);
}

/* src/pd_ard.cpp:55-191 */
void pdo_ard_step(const PdoGrid* g, const PdoConfig* cfg, double dt, double volume_loss,
                  const double* Cc, const double* vel, const uint8_t* is_gb,
                  const uint8_t* is_precip, double* C_new) {
    int dim = g->dim;
    uint8_t* salt = (uint8_t*)calloc((size_t)g->N, 1);
#pragma omp parallel for schedule(dynamic, 256)
    for (long long n = 0; n < g->N; ++n) { /* :61-73 */
        if (g->node_type[n] != PDO_SOLID) continue;
        int i, j, k;
        ijk(g, n, &i, &j, &k);
        for (int o = 0; o < g->n_off; ++o) {
            long long nn = nbr(g, i, j, k, o);
            if (nn < 0 || g->off_vol[o] < 1e-30) continue;
            if (g->node_type[nn] == PDO_FLUID && Cc[nn] >= cfg->C_sat) { salt[n] = 1; break; }
        }
    }
    double decay = 1.0; /* :75-79 */
    if (cfg->corrosion_decay_l > 0.0) decay = pow(10.0, -volume_loss / cfg->corrosion_decay_l);
    double alpha_p = (double)dim, V_H, beta; /* src/pd_ard.cpp:6-15 */
    if (dim == 2) {
        V_H = PDO_PI * cfg->delta * cfg->delta;
        beta = 4.0 / (PDO_PI * cfg->delta * cfg->delta);
    } else {
        V_H = (4.0 / 3.0) * PDO_PI * cfg->delta * cfg->delta * cfg->delta;
        beta = 12.0 / (PDO_PI * cfg->delta * cfg->delta);
    }
    double div_coeff = alpha_p / V_H;

#pragma omp parallel for schedule(dynamic, 256)
    for (long long n = 0; n < g->N; ++n) {
        uint8_t ti = g->node_type[n];
        if (ti != PDO_FLUID && ti != PDO_SOLID) { C_new[n] = Cc[n]; continue; } /* :86-89 */
        int i, j, k;
        ijk(g, n, &i, &j, &k);
        double C_i = Cc[n];
        int i_fluid = ti == PDO_FLUID, i_solid = ti == PDO_SOLID;
        double vi[3] = {0, 0, 0};
        if (i_fluid) for (int d = 0; d < dim; ++d) vi[d] = vel[n * dim + d];
        double vi_mag = i_fluid ? speed(vel, n, dim) : 0.0;
        double diff_sum = 0.0, adv_sum = 0.0;
        for (int o = 0; o < g->n_off; ++o) {
            long long nn = nbr(g, i, j, k, o);
            if (nn < 0) continue;
            double xi = g->off_dist[o], V_j = g->off_vol[o];
            const double* e = &g->off_evec[o * dim];
            if (V_j < 1e-30) continue;
            uint8_t tj = g->node_type[nn];
            if (tj == PDO_WALL || tj == PDO_OUTSIDE) continue; /* :120 */
            double C_j = Cc[nn];
            double inv_xi = 1.0 / xi, inv_xi2 = inv_xi * inv_xi;
            int j_fluid = (tj == PDO_FLUID || tj == PDO_INLET || tj == PDO_OUTLET);
            int j_solid = tj == PDO_SOLID;
            if (i_solid && j_solid) continue; /* :134 */
            double D_avg = 0.0;
            if (i_fluid && j_fluid) {
                D_avg = cfg->D_liquid;
            } else { /* :140-162 */
                long long s_idx = i_solid ? n : nn;
                if (salt[s_idx]) {
                    D_avg = 0.0;
                } else {
                    double D_s = is_gb[s_idx] ? cfg->D_gb
                                              : (is_precip[s_idx] ? cfg->D_precip : cfg->D_grain);
                    D_s *= decay;
                    D_avg = 2.0 * cfg->D_liquid * D_s / (cfg->D_liquid + D_s + 1e-30);
                }
            }
            double D_art = 0.0;
            if (i_fluid && j_fluid) { /* :166-170 */
                double vj_mag = speed(vel, nn, dim);
                D_art = cfg->alpha_art_diff * (vi_mag > vj_mag ? vi_mag : vj_mag) * cfg->dx;
            }
            diff_sum += beta * (D_avg + D_art) * (C_j - C_i) * inv_xi2 * V_j; /* :173 */
            if (i_fluid && j_fluid) { /* :178-181 */
                double vde = 0.0;
                for (int d = 0; d < dim; ++d) vde += vi[d] * e[d];
                adv_sum += (C_j - C_i) * vde * inv_xi * V_j;
            }
        }
        adv_sum *= div_coeff;
        double cn = C_i + dt * (diff_sum - adv_sum);
        C_new[n] = cn < 0.0 ? 0.0 : cn;
    }
    free(salt);
}

/* src/pd_ard.cpp:193-212 (+ the wall-mirror fallback may see new FLUID nodes) */
int pdo_phase_change(PdoGrid* g, const PdoConfig* cfg, uint8_t* phase, double* rho, double* vel,
                     double* Cc, int* dissolved) {
    int dim = g->dim, cnt = 0;
    for (long long n = 0; n < g->N; ++n) {
        if (phase[n] == 0 && g->node_type[n] == PDO_SOLID && Cc[n] < cfg->C_thresh) {
            phase[n] = 1;
            g->node_type[n] = PDO_FLUID;
            rho[n] = cfg->rho_f;
            for (int d = 0; d < dim; ++d) vel[n * dim + d] = 0.0;
            Cc[n] = cfg->C_thresh;
            if (dissolved) dissolved[cnt] = (int)n;
            ++cnt;
        }
    }
    if (cnt > 0) pdo_wall_mirror_build(g, cfg);
    return cnt;
}

/* ------------------------------------------------------------------------- */
/* Loop bodies                                                                */
/* ------------------------------------------------------------------------- */
/* src/pd_ns.cpp:196-205,325 */
double pdo_ns_iterate(PdoGrid* g, const PdoConfig* cfg, int iters, double dt, double* rho,
                      double* vel, double* pressure, double* Cc, double* rho_new, double* vel_new) {
    size_t nb = sizeof(double) * (size_t)g->N;
    double *r0 = rho, *v0 = vel, *r1 = rho_new, *v1 = vel_new;
    double t0 = omp_get_wtime();
    for (int it = 0; it < iters; ++it) {
        pdo_inlet_bc(g, cfg, r0, v0, Cc);
        pdo_outlet_bc(g, cfg, r0, v0, Cc);
        pdo_wall_bc(g, cfg, r0, v0);
        pdo_solid_bc(g, v0);
        pdo_ns_step(g, cfg, dt, r0, v0, pressure, r1, v1);
        pdo_wall_bc(g, cfg, r1, v1);
        double* t;
        t = r0; r0 = r1; r1 = t;
        t = v0; v0 = v1; v1 = t;
    }
    double el = omp_get_wtime() - t0;
    if (r0 != rho) { /* odd count: swap contents so that (rho, vel) hold the current state */
        double* tmp = (double*)malloc(nb * g->dim);
        memcpy(tmp, rho, nb); memcpy(rho, rho_new, nb); memcpy(rho_new, tmp, nb);
        memcpy(tmp, vel, nb * g->dim); memcpy(vel, vel_new, nb * g->dim);
        memcpy(vel_new, tmp, nb * g->dim);
        free(tmp);
    }
    return el;
}

/* src/coupling.cpp:232-240 */
double pdo_ard_iterate(PdoGrid* g, const PdoConfig* cfg, int steps, double dt, double volume_loss,
                       double* rho, double* vel, double* Cc, double* C_new, const uint8_t* is_gb,
                       const uint8_t* is_precip) {
    size_t nb = sizeof(double) * (size_t)g->N;
    double *c0 = Cc, *c1 = C_new;
    double t0 = omp_get_wtime();
    for (int it = 0; it < steps; ++it) {
        pdo_inlet_bc(g, cfg, rho, vel, c0);
        pdo_outlet_bc(g, cfg, rho, vel, c0);
        pdo_wall_conc_bc(g, c0);
        pdo_ard_step(g, cfg, dt, volume_loss, c0, vel, is_gb, is_precip, c1);
        double* t = c0; c0 = c1; c1 = t;
    }
    double el = omp_get_wtime() - t0;
    if (c0 != Cc) {
        double* tmp = (double*)malloc(nb);
        memcpy(tmp, Cc, nb); memcpy(Cc, C_new, nb); memcpy(C_new, tmp, nb);
        free(tmp);
    }
    return el;
}

/* ---- VTKWriter::write (src/vtk_writer.cpp:16-146) ------------------------------------------ */
static double vti_safe(double v) {                     /* src/vtk_writer.cpp:8-14 */
    if (isnan(v) || isinf(v)) return 0.0;
    if (v != 0.0 && fabs(v) < 1e-300) return 0.0;
    return v;
}

int pdo_write_vti(const PdoGrid* g, const char* path, const double* rho, const double* vel,
                  const double* pressure, const double* Cc, const uint8_t* phase, const int* grain_id,
                  const double* D_map, const uint8_t* is_gb, const uint8_t* is_precip) {
    FILE* f = fopen(path, "w");
    if (!f) return 1;
    const int dim = g->dim;
    const int nx = g->Nx, ny = g->Ny, nz = dim == 3 ? g->Nz : 1;
    const long long N = g->N;
    fprintf(f, "<?xml version=\"1.0\"?>\n");
    fprintf(f, "<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\">\n");
    fprintf(f, "  <ImageData WholeExtent=\"0 %d 0 %d 0 %d\" Origin=\"%g %g %g\" Spacing=\"%g %g %g\">\n",
            nx - 1, ny - 1, nz - 1, g->origin[0], g->origin[1], dim == 3 ? g->origin[2] : 0.0, g->dx, g->dx, g->dx);
    fprintf(f, "    <Piece Extent=\"0 %d 0 %d 0 %d\">\n", nx - 1, ny - 1, nz - 1);
    fprintf(f, "      <PointData Scalars=\"phase\" Vectors=\"velocity\">\n");
    fprintf(f, "        <DataArray type=\"Float64\" Name=\"velocity\" NumberOfComponents=\"3\" format=\"ascii\">\n");
    for (long long i = 0; i < N; ++i) {
        int fict = (g->node_type[i] == PDO_WALL || g->node_type[i] == PDO_OUTSIDE);
        double a = fict ? 0.0 : vti_safe(vel[i * dim]), b = fict ? 0.0 : vti_safe(vel[i * dim + 1]);
        if (dim == 2) fprintf(f, "          %g %g 0\n", a, b);
        else fprintf(f, "          %g %g %g\n", a, b, fict ? 0.0 : vti_safe(vel[i * dim + 2]));
    }
    fprintf(f, "        </DataArray>\n");
#define PDO_VTI_F64(name, arr)                                                                  \
    fprintf(f, "        <DataArray type=\"Float64\" Name=\"" name "\" format=\"ascii\">\n");    \
    for (long long i = 0; i < N; ++i) fprintf(f, "          %g\n", vti_safe((arr) ? (arr)[i] : 0.0)); \
    fprintf(f, "        </DataArray>\n")
#define PDO_VTI_INT(type, name, expr)                                                           \
    fprintf(f, "        <DataArray type=\"" type "\" Name=\"" name "\" format=\"ascii\">\n");   \
    for (long long i = 0; i < N; ++i) fprintf(f, "          %d\n", (int)(expr));                \
    fprintf(f, "        </DataArray>\n")
    PDO_VTI_F64("pressure", pressure);
    PDO_VTI_F64("density", rho);
    PDO_VTI_F64("concentration", Cc);
    PDO_VTI_INT("UInt8", "phase", phase[i]);
    PDO_VTI_INT("UInt8", "node_type", g->node_type[i]);
    PDO_VTI_INT("Int32", "grain_id", grain_id ? grain_id[i] : -1);
    PDO_VTI_F64("D_map", D_map);
    PDO_VTI_INT("UInt8", "is_grain_boundary", is_gb[i]);
    PDO_VTI_INT("UInt8", "is_precipitate", is_precip[i]);
#undef PDO_VTI_F64
#undef PDO_VTI_INT
    fprintf(f, "      </PointData>\n    </Piece>\n  </ImageData>\n</VTKFile>\n");
    fclose(f);
    return 0;
}
