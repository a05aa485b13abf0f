// geom.cuh -- lattice geometry with the reference Release build's exact floating-point
// pattern.  The reference compiles `origin + idx*dx`, `R_tube + m*dx` and `px*px + py*py`
// to single FMAs (g++ -O3 -march=native, CMakeLists.txt:89-91), and node classification
// flips whole planes when that changes.  Everything discrete is therefore
// written with explicit fma(): std::fma on the host and the device's fma() are both
// correctly rounded, so host tables and device kernels agree with the reference bit for
// bit.  The pattern was read off the reference's object code.
#pragma once
#include <cmath>

#include "../../include/pdgpu.h"

#ifdef __CUDACC__
#define PD_HD __host__ __device__ __forceinline__
#else
#define PD_HD inline
#endif

struct GeomParams {
    double dx, ox, oy, oz;     // spacing, origin
    double R_tube, R_wire, L_wire;
    double z_phys_min, z_phys_max;   // -L_upstream, L_wire + L_downstream
    double wall_lim;           // R_tube + m*dx + 0.5*dx  (two fused adds)
    double R_wire2;            // R_wire*R_wire
    int dim, m;
};

// Grid::build extents (src/grid.cpp:38-67)
inline void geom_extents(const PdConfig& c, int dim, int* Nx, int* Ny, int* Nz, double origin[3]) {
    double m = (double)c.m_ratio, dx = c.dx;
    double z_min = -std::fma(m, dx, c.L_upstream);
    double z_max = std::fma(m, dx, c.L_wire + c.L_downstream);
    double r_min = -std::fma(m, dx, c.R_tube);
    double r_max = std::fma(m, dx, c.R_tube);
    int nr = (int)std::round((r_max - r_min) / dx) + 1;
    int nz = (int)std::round((z_max - z_min) / dx) + 1;
    if (dim == 2) {
        *Nx = nr; *Ny = nz; *Nz = 1;
        origin[0] = r_min; origin[1] = z_min; origin[2] = 0.0;
    } else {
        *Nx = nr; *Ny = nr; *Nz = nz;
        origin[0] = r_min; origin[1] = r_min; origin[2] = z_min;
    }
}

inline GeomParams geom_params(const PdConfig& c, int dim, const double origin[3]) {
    GeomParams g;
    g.dx = c.dx; g.ox = origin[0]; g.oy = origin[1]; g.oz = origin[2];
    g.R_tube = c.R_tube; g.R_wire = c.R_wire; g.L_wire = c.L_wire;
    g.z_phys_min = -c.L_upstream;
    g.z_phys_max = c.L_wire + c.L_downstream;
    g.wall_lim = std::fma(0.5, c.dx, std::fma((double)c.m_ratio, c.dx, c.R_tube));
    g.R_wire2 = c.R_wire * c.R_wire;
    g.dim = dim; g.m = c.m_ratio;
    return g;
}

PD_HD double geom_coord(double o, int idx, double dx) { return fma((double)idx, dx, o); }

// Node classification (src/grid.cpp:94-147). (i,j,k) are GLOBAL lattice indices; in 2D
// j is the axial index and k is ignored.
PD_HD unsigned char geom_classify(const GeomParams& g, int i, int j, int k) {
    double px = geom_coord(g.ox, i, g.dx);
    double py = geom_coord(g.oy, j, g.dx);
    double axial, radial, r2 = 0.0;
    if (g.dim == 2) {
        axial = py;
        radial = fabs(px);
    } else {
        axial = geom_coord(g.oz, k, g.dx);
        r2 = fma(px, px, py * py);
        radial = sqrt(r2);
    }
    if (axial < g.z_phys_min) {
        if (radial <= g.R_tube) return PDGPU_INLET;
        return radial <= g.wall_lim ? PDGPU_WALL : PDGPU_OUTSIDE;
    }
    if (axial > g.z_phys_max) {
        if (radial <= g.R_tube) return PDGPU_OUTLET;
        return radial <= g.wall_lim ? PDGPU_WALL : PDGPU_OUTSIDE;
    }
    if (radial <= g.R_tube) {
        bool wire;
        if (g.dim == 2) wire = (fabs(px) <= g.R_wire) && (py >= 0.0) && (py <= g.L_wire);
        else wire = (r2 <= g.R_wire2) && (axial >= 0.0) && (axial <= g.L_wire);
        return wire ? PDGPU_SOLID_MG : PDGPU_FLUID;
    }
    return radial <= g.wall_lim ? PDGPU_WALL : PDGPU_OUTSIDE;
}

// Geometric mirror of a WALL node across the tube wall (src/boundary.cpp:158-229).
// Returns false when the reference takes the `goto fallback` / r <= R_tube path;
// otherwise (*im, *jm) are the rounded in-plane lattice indices of the mirror point
// (not yet range- or type-checked). In 2D *jm is unused.
PD_HD bool geom_wall_mirror(const GeomParams& g, int i, int j, int* im, int* jm) {
    double x = geom_coord(g.ox, i, g.dx);
    if (g.dim == 2) {
        double xm;
        if (x > g.R_tube) xm = 2.0 * g.R_tube - x;
        else if (x < -g.R_tube) xm = -2.0 * g.R_tube - x;
        else return false;
        *im = (int)round((xm - g.ox) / g.dx);
        *jm = 0;
        return true;
    }
    double y = geom_coord(g.oy, j, g.dx);
    double r = sqrt(fma(x, x, y * y));
    if (!(r > g.R_tube && r > 1e-30)) return false;
    double rm = 2.0 * g.R_tube - r;
    double xm = x * rm / r, ym = y * rm / r;
    *im = (int)round((xm - g.ox) / g.dx);
    *jm = (int)round((ym - g.oy) / g.dx);
    return true;
}

// Prescribed inlet velocity (src/boundary.cpp:38-52)
PD_HD double geom_inlet_velocity(const GeomParams& g, double U_in, int i, int j) {
    double px = geom_coord(g.ox, i, g.dx);
    double R2 = g.R_tube * g.R_tube;
    if (g.dim == 2) {
        double rr = (px * px) / R2;
        if (rr > 1.0) rr = 1.0;
        return 1.5 * U_in * (1.0 - rr);
    }
    double py = geom_coord(g.oy, j, g.dx);
    double rr = fma(px, px, py * py) / R2;
    if (rr > 1.0) rr = 1.0;
    return 2.0 * U_in * (1.0 - rr);
}
