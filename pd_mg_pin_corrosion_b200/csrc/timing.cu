// timing.cu -- kernel-only timing of the dominant bond kernels (bench.py's roofline leg).
#include "common.cuh"

extern "C" int pdgpu_time_kernel(pdgpu_ctx* c, int which, int reps, float* ms_avg) {
    NEED_FIELDS(c);
    if (!ms_avg || reps < 1) PD_FAIL("pdgpu_time_kernel: bad arguments");
    // which = 1 times the ARD bond kernel(s) alone (FLUID tiles + SOLID_MG rows): the |v| / packed-weight
    // pass, the salt pre-pass and -- for slab contexts -- the exchange of the ghost solids' weights run
    // once, untimed, so that no NCCL call and no pre-pass sits inside the timed region
    if (which != 0) {
        PD_TRY(pd_ensure_vmag(c, c->cur));
        PD_TRY(pd_enqueue_ard_prepass_solids(c, c->curC));
        if (c->nranks > 1 && c->comm) PD_TRY(pd_enqueue_halo(c, 3, c->cur, c->curC));
    }
    // one untimed launch (instruction cache, clocks)
    if (which == 0) PD_TRY(pd_enqueue_ns_step(c, c->cur, c->d_dt));
    else PD_TRY(pd_enqueue_ard_main(c, c->cur, c->curC, c->d_dt + 1, -1, -1, true));
    CUDA_OK(cudaEventRecord(c->ev_t0, c->stream));
    for (int r = 0; r < reps; ++r) {
        if (which == 0) PD_TRY(pd_enqueue_ns_step(c, c->cur, c->d_dt));
        else PD_TRY(pd_enqueue_ard_main(c, c->cur, c->curC, c->d_dt + 1, -1, -1, true));
    }
    CUDA_OK(cudaEventRecord(c->ev_t1, c->stream));
    CUDA_OK(cudaEventSynchronize(c->ev_t1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
    *ms_avg = ms / reps;
    return 0;
}

// FP64 FMA peak of this device, measured (MEASURED_PEAKS.json has no FP64 figure): 8 independent
// DFMA chains per thread, 148*8 CTAs of 256 threads. Reported as the denominator of the
// FP64-pipe view of the tiled bond kernels, next to the CSR-equivalent HBM view.
__global__ void __launch_bounds__(256)
k_dfma_peak(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;   // keep the chains alive
}

extern "C" int pdgpu_fp64_peak(pdgpu_ctx* c, double* tflops) {
    CHECK_CTX(c);
    if (!tflops) PD_FAIL("pdgpu_fp64_peak: null output");
    const int iters = 1 << 14, blocks = 148 * 8, threads = 256;
    k_dfma_peak<<<blocks, threads, 0, c->stream>>>(c->d_red, 64, 0.999999, 1e-9);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_OK(cudaEventRecord(c->ev_t0, c->stream));
        k_dfma_peak<<<blocks, threads, 0, c->stream>>>(c->d_red, iters, 0.999999, 1e-9);
        CUDA_OK(cudaEventRecord(c->ev_t1, c->stream));
        CUDA_OK(cudaEventSynchronize(c->ev_t1));
        float ms = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
        if (ms < best) best = ms;
    }
    double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    return 0;
}

// Same measurement with three distinct register-pair sources per DFMA (x = fma(x, y, z) with
// per-chain y, z): the register-file operand bandwidth, not the FP64 pipe, bounds this form.
__global__ void __launch_bounds__(256)
k_dfma_peak3(double* out, int iters) {
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = threadIdx.x + i;
        y[i] = 1.0 - 1e-9 * (threadIdx.x + i);
        z[i] = 1e-9 * (blockIdx.x + i);
    }
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], y[i], z[i]);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

extern "C" int pdgpu_fp64_peak3(pdgpu_ctx* c, double* tflops) {
    CHECK_CTX(c);
    if (!tflops) PD_FAIL("pdgpu_fp64_peak3: null output");
    const int iters = 1 << 14, blocks = 148 * 8, threads = 256;
    k_dfma_peak3<<<blocks, threads, 0, c->stream>>>(c->d_red, 64);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_OK(cudaEventRecord(c->ev_t0, c->stream));
        k_dfma_peak3<<<blocks, threads, 0, c->stream>>>(c->d_red, iters);
        CUDA_OK(cudaEventRecord(c->ev_t1, c->stream));
        CUDA_OK(cudaEventSynchronize(c->ev_t1));
        float ms = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
        if (ms < best) best = ms;
    }
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return 0;
}
