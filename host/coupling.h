// host/coupling.h -- host driver of the explicit coupling loop over libpdgpu.so: same phases,
// triggers and CSV output as the reference's CoupledSolver::run (src/coupling.cpp:82-302,
// explicit branch :217-253), with the solvers living on the device.
#pragma once
#include <iosfwd>
#include <string>
#include <utility>
#include <vector>

#include "config.h"
#include "grains.h"

struct pdgpu_ctx;

// Host-side Grid/Fields of the driver: what the reference keeps in std::vectors and the GPU
// path still needs on the host (types for the solid list, D_map / grain_id for output).
struct HostState {
    int dim = 2, Nx = 0, Ny = 0, Nz = 0;
    long long N = 0;
    std::vector<uint8_t> node_type;
    std::vector<double> D_map;
    std::vector<int> grain_id;
};

// PVD collection of the snapshots written so far, rewritten after every entry like the
// reference's VTKWriter::add_timestep / write_pvd (src/vtk_writer.cpp:150-186).
class PvdSeries {
public:
    void set_path(const std::string& p) { path_ = p; }
    void add_timestep(double time, const std::string& file);
    void save(std::ostream& out) const;   // driver checkpoint
    void load(std::istream& in);
private:
    std::string path_;
    std::vector<std::pair<double, std::string>> entries_;
};

class CoupledSolver {
public:
    // returns the final corrosion time
    double run(pdgpu_ctx* ctx, HostState& st, const HostConfig& cfg, bool verbose = true);
    // Restart (new; the reference cannot resume): after every `checkpoint_every` coupling cycles the device
    // state (pdgpu_checkpoint_save) and the loop state of this driver are written to
    // <checkpoint_prefix>_c<cycle>.pdck / .drv; `resume_prefix` continues such a pair. A resumed run appends to
    // diagnostics.csv / mass_loss.csv and produces the rows the uninterrupted run produces.
    std::string checkpoint_prefix, resume_prefix;
    int checkpoint_every = 0;
    int rank = 0, nranks = 1;   // z-slab runs: every rank runs the loop on its slab, rank 0 writes the files
    bool write_vti = true;   // state_/flow_/corr_/final_ snapshots + simulation.pvd / flow.pvd (src/coupling.cpp:117-147,242-246,292-296)

private:
    std::vector<int> initial_solid_indices_;
    int total_dissolved_ = 0, dissolved_since_flow_ = 0, total_implicit_steps_ = 0;
    double solid_C_sum(pdgpu_ctx* ctx);
    void write_diagnostics(pdgpu_ctx* ctx, double t_corr, const HostConfig& cfg);
    // VTKWriter::write through libpdgpu.so (text formatted on the device) + PVD bookkeeping
    void snapshot(pdgpu_ctx* ctx, const HostState& st, const HostConfig& cfg, const char* prefix, double t,
                  PvdSeries& series, bool count_frame);
    PvdSeries writer_, flow_writer_;
    int frame_count_ = 0;
    void save_driver_state(const std::string& path, const HostState& st, double t_corr, int cycle, bool need_flow) const;
    bool load_driver_state(const std::string& path, HostState& st, double* t_corr, int* cycle, bool* need_flow);
};
