// sdfc_microbench.cu — roofline denominators the driver does not measure.
//
// MEASURED_PEAKS.json holds the HBM copy bandwidth and the bf16 tensor throughput of this pool's B200s.
// The contact path is bounded elsewhere: the per-contact force / Jacobian algebra by the FP64 pipe, the
// cache-resident detection of the planner batches (C2: two grids + two point sets shared by 1024
// rollouts) by L2 / L1 delivery. These two kernels measure those ceilings on the device the bench runs
// on, so that every roofline fraction in bench.py divides by a number measured in the same process.
#include "sdfc_host.h"

namespace sdfc {

// 8 independent DFMA chains per thread: 2 flops x 8 x iters per thread.
__global__ void __launch_bounds__(256) k_peak_fp64(int iters, double seed, double* out) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
        a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) out[0] = s;   // never true: keeps the chains alive
}

// Every CTA streams the whole buffer (which fits L2) with 128-bit loads: after the first pass all
// traffic is L2 -> SM.
__global__ void __launch_bounds__(256) k_peak_l2(const uint4* __restrict__ buf, size_t n_vec, int passes, unsigned* out) {
    unsigned acc = 0;
    const size_t stride = (size_t)blockDim.x;
    for (int p = 0; p < passes; ++p) {
        // start offset differs per CTA so that the SMs do not hit the same slices in lock step
        const size_t start = ((size_t)blockIdx.x * 7919u + (size_t)p * 104729u) % n_vec;
        for (size_t i = threadIdx.x; i < n_vec; i += stride) {
            size_t j = start + i;
            if (j >= n_vec) j -= n_vec;
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + j));
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x12345u) out[0] = acc;
}

}  // namespace sdfc

using namespace sdfc;

extern "C" sdfc_status sdfc_measure_peaks(sdfc_context* ctx, double out[4]) {
    SDFC_REQUIRE(ctx && out, "sdfc_measure_peaks: NULL argument");
    SDFC_CUDA(cudaSetDevice(ctx->device));
    out[0] = out[1] = out[2] = out[3] = 0.0;
    cudaStream_t st = ctx->stream;
    cudaEvent_t a, b;
    SDFC_CUDA(cudaEventCreate(&a));
    SDFC_CUDA(cudaEventCreate(&b));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    double* d_out = nullptr;
    SDFC_CUDA(cudaMalloc((void**)&d_out, 64));
    float best;
    {   // FP64 pipe: DFMA throughput
        const int iters = 1 << 14, blocks = sms * 8;
        best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(a, st);
            k_peak_fp64<<<blocks, 256, 0, st>>>(iters, 1.0 + rep, d_out);
            cudaEventRecord(b, st);
            SDFC_CUDA(cudaEventSynchronize(b));
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            if (rep) best = ms < best ? ms : best;
        }
        ctx->launches += 5;
        out[0] = 2.0 * 8.0 * iters * 256.0 * blocks / (best * 1e-3) / 1e9;   // GFLOP/s
    }
    {   // L2 -> SM read bandwidth over a 48 MiB buffer (L2: 126 MB)
        const size_t bytes = (size_t)48 << 20, n_vec = bytes / sizeof(uint4);
        uint4* buf = nullptr;
        SDFC_CUDA(cudaMalloc((void**)&buf, bytes));
        SDFC_CUDA(cudaMemsetAsync(buf, 1, bytes, st));
        const int blocks = sms * 4, passes = 2;
        best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a, st);
            k_peak_l2<<<blocks, 256, 0, st>>>(buf, n_vec, passes, reinterpret_cast<unsigned*>(d_out));
            cudaEventRecord(b, st);
            SDFC_CUDA(cudaEventSynchronize(b));
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            if (rep) best = ms < best ? ms : best;
        }
        ctx->launches += 4;
        out[1] = (double)bytes * passes * blocks / (best * 1e-3) / 1e9;   // GB/s
        cudaFree(buf);
    }
    cudaFree(d_out);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    SDFC_CUDA(cudaGetLastError());
    return SDFC_OK;
}
