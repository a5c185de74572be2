// FP64 peak micro-benchmark for B200 (sm_100a).
//
// MEASURED_PEAKS.json carries HBM and bf16 peaks only; the Kalman-filter hot path is bound by the
// FP64 pipes, so the roofline denominator for it is measured here:
//   (1) DFMA peak      : independent fma.rn.f64 chains in registers
//   (2) DMMA peak      : mma.sync.aligned.m8n8k4.row.col.f64 chains in registers
//   (3) DFMA+DMMA mix  : do the two pipes overlap?
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NT>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
    double c0[NT], c1[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

template <int NT, int NACC>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b) {
    double c0[NT], c1[NT], acc[NACC];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) dmma884(c0[i], c1[i], a, b);
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// DFMA with one shared-memory operand per FMA pair (models the register-tiled filter inner loop:
// TS*TS FMAs fed by 2*TS doubles read with LDS.128/LDS.64)
template <int TS>
__global__ void __launch_bounds__(256) k_dfma_lds(double* out, int iters, int n) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 2 * n * 64; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
    __syncthreads();
    double acc[TS][TS];
#pragma unroll
    for (int i = 0; i < TS; ++i)
#pragma unroll
        for (int j = 0; j < TS; ++j) acc[i][j] = 0;
    const int a = (threadIdx.x >> 2) & 3, b = threadIdx.x & 3;
    const int bs = TS + (TS & 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int k = 0; k < n; ++k) {
            const double* pa = sm + k * 64 + a * bs;
            const double* pb = sm + (n + k) * 64 + b * bs;
            double x[TS], y[TS];
#pragma unroll
            for (int i = 0; i + 1 < TS; i += 2) {
                double2 v = *reinterpret_cast<const double2*>(pa + i); x[i] = v.x; x[i + 1] = v.y;
                double2 u = *reinterpret_cast<const double2*>(pb + i); y[i] = u.x; y[i + 1] = u.y;
            }
            if (TS & 1) { x[TS - 1] = pa[TS - 1]; y[TS - 1] = pb[TS - 1]; }
#pragma unroll
            for (int i = 0; i < TS; ++i)
#pragma unroll
                for (int j = 0; j < TS; ++j) acc[i][j] = fma(x[i], y[j], acc[i][j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TS; ++i)
#pragma unroll
        for (int j = 0; j < TS; ++j) s += acc[i][j];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

// Roofline denominators for bench.py / tools/sweep.py (MEASURED_PEAKS.json has no FP64 entry).  Built as its own
// small library, tools/libfp64peak.so (bild_b200/build.py), NOT part of the product ABI:
//   fp64_peak_measure(device, &dfma_tflops, &dmma_tflops) -> 0 on success, else the CUDA error code
extern "C" int fp64_peak_measure(int device, double* dfma_tflops, double* dmma_tflops) {
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return static_cast<int>(e);
    int sms = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return static_cast<int>(e);
    double* d = nullptr;
    if ((e = cudaMalloc(&d, 8)) != cudaSuccess) return static_cast<int>(e);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, blocks = sms * 4;   // 32 warps per SM
    double best[2] = {1e30, 1e30};
    for (int which = 0; which < 2; ++which)
        for (int rep = 0; rep < 7; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) k_dfma<16><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9);
            else k_dmma<8><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2 && ms < best[which]) best[which] = ms;
        }
    e = cudaGetLastError();
    *dfma_tflops = 2.0 * 16 * iters * 256.0 * blocks / best[0] * 1e-9;
    *dmma_tflops = 2.0 * 256 * 8 * iters * 8.0 * blocks / best[1] * 1e-9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return static_cast<int>(e);
}

#ifndef FP64_PEAK_NO_MAIN
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 4096;
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);

    for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM
        int blocks = sms * wps / 8;
        double ms = time_ms([&] { k_dfma<16><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 16 * iters * 256.0 * blocks;
        printf(",\n \"dfma_tflops_w%d\": %.3f", wps, fl / ms * 1e-9);
    }
    for (int wps = 4; wps <= 32; wps *= 2) {
        int blocks = sms * wps / 8;
        double ms = time_ms([&] { k_dmma<8><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 256 * 8 * iters * 8.0 * blocks;   // 8x8x4 MACs per warp-instr, 8 warps/block
        printf(",\n \"dmma_tflops_w%d\": %.3f", wps, fl / ms * 1e-9);
    }
    {
        int blocks = sms * 2;
        double ms = time_ms([&] { k_mix<4, 32><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = (2.0 * 256 * 4 * 8.0 + 2.0 * 32 * 256.0) * iters * blocks;
        printf(",\n \"mix_dmma4_dfma32_tflops\": %.3f", fl / ms * 1e-9);
        ms = time_ms([&] { k_mix<8, 16><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        fl = (2.0 * 256 * 8 * 8.0 + 2.0 * 16 * 256.0) * iters * blocks;
        printf(",\n \"mix_dmma8_dfma16_tflops\": %.3f", fl / ms * 1e-9);
    }
    {
        const int n = 20, it2 = 2000;
        size_t smem = 2 * n * 64 * sizeof(double);
        for (int wps = 4; wps <= 16; wps *= 2) {
            int blocks = sms * wps / 8;
            double ms = time_ms([&] { k_dfma_lds<5><<<blocks, 256, smem>>>(out, it2, n); }, 5);
            double fl = 2.0 * 25 * n * it2 * 256.0 * blocks;
            printf(",\n \"dfma_lds_ts5_tflops_w%d\": %.3f", wps, fl / ms * 1e-9);
            ms = time_ms([&] { k_dfma_lds<4><<<blocks, 256, smem>>>(out, it2, n); }, 5);
            fl = 2.0 * 16 * n * it2 * 256.0 * blocks;
            printf(",\n \"dfma_lds_ts4_tflops_w%d\": %.3f", wps, fl / ms * 1e-9);
            ms = time_ms([&] { k_dfma_lds<6><<<blocks, 256, smem>>>(out, it2, n); }, 5);
            fl = 2.0 * 36 * n * it2 * 256.0 * blocks;
            printf(",\n \"dfma_lds_ts6_tflops_w%d\": %.3f", wps, fl / ms * 1e-9);
            ms = time_ms([&] { k_dfma_lds<8><<<blocks, 256, smem>>>(out, it2, n); }, 5);
            fl = 2.0 * 64 * n * it2 * 256.0 * blocks;
            printf(",\n \"dfma_lds_ts8_tflops_w%d\": %.3f", wps, fl / ms * 1e-9);
        }
    }
    printf("\n}\n");
    return 0;
}
#endif
