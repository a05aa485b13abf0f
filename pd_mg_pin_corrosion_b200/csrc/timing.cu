// timing.cu -- kernel-only timing of the dominant bond kernels (bench.py's roofline leg).
#include "common.cuh"

extern "C" int pdgpu_time_kernel(pdgpu_ctx* c, int which, int reps, float* ms_avg) {
    NEED_FIELDS(c);
    if (!ms_avg || reps < 1) PD_FAIL("pdgpu_time_kernel: bad arguments");
    long long before = c->launches;
    // one untimed launch (instruction cache, clocks)
    if (which == 0) PD_TRY(pd_enqueue_ns_step(c, c->cur, c->d_dt));
    else PD_TRY(pd_enqueue_ard_step(c, c->cur, c->curC, c->d_dt + 1));
    CUDA_OK(cudaEventRecord(c->ev_t0, c->stream));
    for (int r = 0; r < reps; ++r) {
        if (which == 0) PD_TRY(pd_enqueue_ns_step(c, c->cur, c->d_dt));
        else PD_TRY(pd_enqueue_ard_step(c, c->cur, c->curC, c->d_dt + 1));
    }
    CUDA_OK(cudaEventRecord(c->ev_t1, c->stream));
    CUDA_OK(cudaEventSynchronize(c->ev_t1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
    *ms_avg = ms / reps;
    (void)before;
    return 0;
}
