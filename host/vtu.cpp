// host/vtu.cpp -- see vtu.h.  The reference streams every value through std::ofstream operator<<; the default
// stream format of a double is printf's "%g" with precision 6, of an int "%d", so the text is built with snprintf
// into one buffer and written with a single fwrite.
#include "vtu.h"

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

namespace {
// non-finite and sub-1e-300 values are printed as 0 (src/vtk_writer.cpp:7-14)
inline double safe_val(double v) {
    if (std::isnan(v) || std::isinf(v)) return 0.0;
    if (v != 0.0 && std::fabs(v) < 1e-300) return 0.0;
    return v;
}
struct Text {
    std::string s;
    char tmp[64];
    void lit(const char* t) { s += t; }
    void num(double v) { s.append(tmp, (size_t)std::snprintf(tmp, sizeof tmp, "%g", v)); }
    void num(int v) { s.append(tmp, (size_t)std::snprintf(tmp, sizeof tmp, "%d", v)); }
};
const char* IND = "          ";
}  // namespace

extern "C" int pdhost_write_vtu(const char* path, int N, const double* pos, const uint8_t* node_type, const double* vel,
                                const double* pressure, const double* C, const uint8_t* phase, const int* grid_level,
                                const double* dx_local, const int* grain_id, const double* D_map, const uint8_t* is_gb,
                                const uint8_t* is_precip) {
    std::vector<int> out;                       // OUTSIDE nodes are not written (:209-216)
    out.reserve((size_t)N);
    for (int i = 0; i < N; ++i)
        if (node_type[i] != 5) out.push_back(i);
    const int n = (int)out.size();
    Text t;
    t.s.reserve((size_t)n * 160 + 4096);
    t.lit("<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\">\n"
          "  <UnstructuredGrid>\n    <Piece NumberOfPoints=\"");
    t.num(n); t.lit("\" NumberOfCells=\""); t.num(n); t.lit("\">\n");
    t.lit("      <Points>\n        <DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n");
    for (int i : out) { t.lit(IND); t.num(pos[2 * i]); t.lit(" "); t.num(pos[2 * i + 1]); t.lit(" 0\n"); }
    t.lit("        </DataArray>\n      </Points>\n      <Cells>\n"
          "        <DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n");
    for (int k = 0; k < n; ++k) { t.lit(IND); t.num(k); t.lit("\n"); }
    t.lit("        </DataArray>\n        <DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n");
    for (int k = 0; k < n; ++k) { t.lit(IND); t.num(k + 1); t.lit("\n"); }
    t.lit("        </DataArray>\n        <DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n");
    for (int k = 0; k < n; ++k) t.lit("          1\n");
    t.lit("        </DataArray>\n      </Cells>\n      <PointData Scalars=\"phase\" Vectors=\"velocity\">\n"
          "        <DataArray type=\"Float64\" Name=\"velocity\" NumberOfComponents=\"3\" format=\"ascii\">\n");
    for (int i : out) {                         // WALL nodes carry the mirror velocity: written as 0 (:256-262)
        const bool wall = node_type[i] == 2;
        t.lit(IND); t.num(wall ? 0.0 : safe_val(vel[2 * i])); t.lit(" "); t.num(wall ? 0.0 : safe_val(vel[2 * i + 1]));
        t.lit(" 0\n");
    }
    auto scalars_d = [&](const char* name, const double* a, bool safe) {
        t.lit("        <DataArray type=\"Float64\" Name=\""); t.lit(name); t.lit("\" format=\"ascii\">\n");
        for (int i : out) { t.lit(IND); t.num(safe ? safe_val(a[i]) : a[i]); t.lit("\n"); }
        t.lit("        </DataArray>\n");
    };
    auto scalars_u8 = [&](const char* name, const uint8_t* a) {
        t.lit("        <DataArray type=\"UInt8\" Name=\""); t.lit(name); t.lit("\" format=\"ascii\">\n");
        for (int i : out) { t.lit(IND); t.num((int)a[i]); t.lit("\n"); }
        t.lit("        </DataArray>\n");
    };
    auto scalars_i = [&](const char* name, const int* a) {
        t.lit("        <DataArray type=\"Int32\" Name=\""); t.lit(name); t.lit("\" format=\"ascii\">\n");
        for (int i : out) { t.lit(IND); t.num(a[i]); t.lit("\n"); }
        t.lit("        </DataArray>\n");
    };
    t.lit("        </DataArray>\n");
    scalars_d("pressure", pressure, true);
    scalars_d("concentration", C, true);
    scalars_u8("phase", phase);
    scalars_u8("node_type", node_type);
    if (grid_level) scalars_i("grid_level", grid_level);
    if (dx_local) scalars_d("dx_local", dx_local, false);
    scalars_i("grain_id", grain_id);
    scalars_d("D_map", D_map, true);
    scalars_u8("is_grain_boundary", is_gb);
    if (is_precip) scalars_u8("is_precipitate", is_precip);
    t.lit("      </PointData>\n    </Piece>\n  </UnstructuredGrid>\n</VTKFile>\n");
    std::FILE* f = std::fopen(path, "wb");
    if (!f) {
        std::fprintf(stderr, "Error: Cannot open VTU file '%s'\n", path);
        return 1;
    }
    const bool ok = std::fwrite(t.s.data(), 1, t.s.size(), f) == t.s.size();
    return (std::fclose(f) == 0 && ok) ? 0 : 1;
}

extern "C" void pdhost_init_dmap(int N, const uint8_t* node_type, const uint8_t* is_gb, const uint8_t* is_precip,
                                 double D_liquid, double D_grain, double D_gb, double D_precip, double* D_map) {
    for (int i = 0; i < N; ++i) {
        switch (node_type[i]) {
            case 1: D_map[i] = is_gb[i] ? D_gb : (is_precip && is_precip[i]) ? D_precip : D_grain; break;   // SOLID_MG
            case 2: case 5: D_map[i] = 0.0; break;                                                         // WALL, OUTSIDE
            default: D_map[i] = D_liquid; break;                        // FLUID, INLET, OUTLET, FICTITIOUS
        }
    }
}
