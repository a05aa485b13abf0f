// ns_tile.cu -- tiled fast path of the PD-NS bond kernel (placeholder until the first
// GPU parity run of the generic kernel is green; see DESIGN.md "kernel plan").
#include "common.cuh"

int pd_enqueue_ns_step_fast(pdgpu_ctx* c, int src, const double* d_dt) {
    (void)c; (void)src; (void)d_dt;
    return -1;   // not applicable -> generic kernel
}
