// DMMA m8n8k4 throughput vs. independent accumulator chains per warp and warps per SM sub-partition.
// Decides how many warps per scheduler the MMA filter kernel needs (latency- or issue-limited?).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <int NT>
__global__ void __launch_bounds__(128) k_dmma(double* out, int iters, double a, double b) {
    double c0[NT], c1[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

template <int NT>
static int run(int sms, double* d, int wps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 2048;
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        k_dmma<NT><<<sms, 32 * wps>>>(d, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r && ms < best) best = ms;
    }
    printf(" \"dmma_chains%d_warps_per_sm%d\": %.2f,\n", NT, wps, 2.0 * 256 * NT * iters * (double)wps * sms / best * 1e-9);
    return 0;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    double* d; CK(cudaMalloc(&d, 8));
    printf("{\n");
    for (int wps = 1; wps <= 4; ++wps) {
        int w = wps == 3 ? 4 : (wps == 4 ? 8 : wps);   // 1, 2, 4, 8 warps per SM (1 CTA per SM; 4 = one per scheduler)
        if (run<1>(prop.multiProcessorCount, d, w)) return 1;
        if (run<2>(prop.multiProcessorCount, d, w)) return 1;
        if (run<4>(prop.multiProcessorCount, d, w)) return 1;
        if (run<9>(prop.multiProcessorCount, d, w)) return 1;
        if (run<16>(prop.multiProcessorCount, d, w)) return 1;
        if (run<49>(prop.multiProcessorCount, d, w)) return 1;
    }
    printf(" \"unit\": \"TFLOP/s\"\n}\n");
    return 0;
}
