// host/vtu.h -- VTU (UnstructuredGrid, one VTK_VERTEX cell per node) snapshot of a two-level AMR cloud:
// byte-identical to the reference's VTKWriter::write_vtu (src/vtk_writer.cpp:199-346; 2D clouds).
#pragma once
#include <cstdint>

extern "C" {
// All arrays have N entries (pos and vel: [N][2]).  grid_level, dx_local and is_precip may be null: the reference
// omits those arrays when the corresponding vectors are empty (:298-315, :338-344).  Returns 0, or 1 if the file
// cannot be written.
int pdhost_write_vtu(const char* path, int N, const double* pos, const uint8_t* node_type, const double* vel,
                     const double* pressure, const double* C, const uint8_t* phase, const int* grid_level,
                     const double* dx_local, const int* grain_id, const double* D_map, const uint8_t* is_gb,
                     const uint8_t* is_precip);
// D_map of initialize_fields (src/main.cpp:19-112) from the node types and grain flags
void pdhost_init_dmap(int N, const uint8_t* node_type, const uint8_t* is_gb, const uint8_t* is_precip, double D_liquid,
                      double D_grain, double D_gb, double D_precip, double* D_map);
}
