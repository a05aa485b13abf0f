// tests/fake_pdgpu/fake_stubs.cpp -- TEST INFRASTRUCTURE ONLY: the lattice entry points host/pd_corrosion_gpu is linked
// against (BIND_NOW), present so that the binary loads next to the fake pdamr_* functions of fake_pdgpu.cpp; every one
// of them reports failure when called (C linkage: the argument lists do not matter to the loader).
extern "C" {
int pdgpu_checkpoint_load() { return 1; }
int pdgpu_checkpoint_save() { return 1; }
int pdgpu_comm_get_uid() { return 1; }
int pdgpu_comm_init() { return 1; }
int pdgpu_comm_uid_bytes() { return 1; }
int pdgpu_grains_grow_precip() { return 1; }
int pdgpu_grains_voronoi() { return 1; }
}
