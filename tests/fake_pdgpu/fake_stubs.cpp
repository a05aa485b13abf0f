// tests/fake_pdgpu/fake_stubs.cpp -- TEST INFRASTRUCTURE ONLY: the lattice entry points host/pd_corrosion_gpu is linked
// against (BIND_NOW), present so that the binary loads next to the fake pdamr_* functions of fake_pdgpu.cpp; every one
// of them reports failure when called (C linkage: the argument lists do not matter to the loader).
extern "C" {
int pdgpu_ard_compute_dt() { return 1; }
int pdgpu_ard_iterate() { return 1; }
int pdgpu_ard_set_volume_loss() { return 1; }
int pdgpu_bc_inlet() { return 1; }
int pdgpu_bc_outlet() { return 1; }
int pdgpu_bc_smooth_conc() { return 1; }
int pdgpu_bc_wall_conc() { return 1; }
int pdgpu_checkpoint_load() { return 1; }
int pdgpu_checkpoint_save() { return 1; }
int pdgpu_comm_allreduce() { return 1; }
int pdgpu_comm_get_uid() { return 1; }
int pdgpu_comm_init() { return 1; }
int pdgpu_comm_uid_bytes() { return 1; }
int pdgpu_create_slab() { return 1; }
int pdgpu_destroy() { return 1; }
int pdgpu_diag() { return 1; }
int pdgpu_fields_download_all() { return 1; }
int pdgpu_fields_init() { return 1; }
int pdgpu_gather() { return 1; }
int pdgpu_grains_grow_precip() { return 1; }
int pdgpu_grains_voronoi() { return 1; }
int pdgpu_grid_build() { return 1; }
int pdgpu_grid_extents() { return 1; }
int pdgpu_grid_info() { return 1; }
int pdgpu_implicit_assemble() { return 1; }
int pdgpu_implicit_compute_dt() { return 1; }
int pdgpu_implicit_step() { return 1; }
int pdgpu_ns_solve_steady() { return 1; }
int pdgpu_phase_change() { return 1; }
int pdgpu_solid_below_thresh() { return 1; }
int pdgpu_vti_write() { return 1; }
}
