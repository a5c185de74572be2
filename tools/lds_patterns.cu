// Shared-memory wavefront cost of broadcast-heavy LDS.64 / LDS.128 access patterns on sm_100a.
// The register-tiled filter kernel reads "row k, my column block" operands that many lanes share;
// this probe measures how many LSU cycles each such warp-wide load costs, per lane->address pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/lds_patterns tools/lds_patterns.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ int pattern_offset(int pat, int lane) {   // byte offset
    switch (pat) {
        case 0: return 0;                                              // full broadcast
        case 1: return lane * 16;                                      // all distinct, contiguous (512 B)
        case 2: return ((lane >> 2) & 3) * 48 + (lane >> 4) * 3840;    // tile-kernel B operand: 4 per filter, 2 filters (same banks)
        case 3: return (lane & 3) * 48 + (lane >> 4) * 5792;           // tile-kernel C operand
        case 4: return (lane >> 4) * 64;                               // 2 distinct (half warps)
        case 5: return (lane >> 3) * 16;                               // 4 distinct, one per quarter warp
        case 6: return (lane & 7) * 16;                                // 8 distinct, same set in every quarter
        case 7: return (lane >> 2) * 16;                               // 8 distinct, 2 per quarter
        case 8: return (lane & 3) * 16;                                // 4 distinct, same set in every quarter
        case 9: return (lane & 1) * 16;                                // 2 distinct interleaved
        case 10: return ((lane >> 2) & 3) * 48;                        // B operand, both filters same state
        case 11: return (lane & 3) * 48 + (lane >> 4) * (5792 + 64);   // C operand, filter stride shifted by 64 B
        case 12: return (lane & 15) * 16;                              // 16 distinct (256 B), same in both halves
        case 13: return (lane >> 1) * 16;                              // 16 distinct, pairs
        case 14: return lane * 8;                                      // 32 distinct doubles, contiguous (256 B)
        case 15: return (lane >> 2) * 192 + (lane & 3) * 8;            // MMA A-frag 8x4 from row-major, row stride 24 doubles
        case 16: return (lane & 3) * 192 + (lane >> 2) * 8;            // MMA B-frag 4x8 from row-major, row stride 24 doubles
        case 17: return (lane >> 2) * 208 + (lane & 3) * 8;            // A-frag, row stride 26 doubles
        case 18: return (lane & 3) * 208 + (lane >> 2) * 8;            // B-frag, row stride 26 doubles
        case 19: return (lane >> 2) * 224 + (lane & 3) * 8;            // A-frag, row stride 28 doubles
        case 20: return (lane & 3) * 224 + (lane >> 2) * 8;            // B-frag, row stride 28 doubles
        case 21: return (lane >> 2) * 32 + (lane & 3) * 8;             // A-frag from a packed 8x4 block (== contiguous)
        case 22: return (lane >> 2) * 160 + (lane & 3) * 8;            // A-frag, row stride 20 doubles
        case 23: return (lane & 3) * 160 + (lane >> 2) * 8;            // B-frag, row stride 20 doubles
        case 24: return (lane >> 2) * 448 + (lane & 3) * 8;            // A-frag, row stride 56 doubles
        case 25: return (lane & 3) * 448 + (lane >> 2) * 8;            // B-frag, row stride 56 doubles
        case 26: return (lane >> 2) * 416 + (lane & 3) * 8;            // A-frag, row stride 52 doubles
        case 27: return (lane & 3) * 416 + (lane >> 2) * 8;            // B-frag, row stride 52 doubles
        default: return 0;
    }
}

template <int WIDTH>   // 8 or 16 bytes
__global__ void k_lds(int pat, int iters, long long* cycles, double* sink) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 16384 / 8; i += blockDim.x) reinterpret_cast<double*>(sm)[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned base = static_cast<unsigned>(__cvta_generic_to_shared(sm)) + pattern_offset(pat, lane);
    // loads are issued back to back and never consumed inside the loop, so only the LSU / shared-memory
    // pipe limits the rate (asm volatile keeps them)
    unsigned r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    unsigned x0, x1, x2, x3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (WIDTH == 16) {
                asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(base + (u & 1) * 32));
            } else {
                asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x0), "=r"(x1) : "r"(base + (u & 1) * 32));
                x2 = x3 = 0;
            }
            if (u == 15) { r0 ^= x0; r1 ^= x1; r2 ^= x2; r3 ^= x3; }
        }
    }
    long long t1 = clock64();
    double acc = (double)(r0 ^ r1 ^ r2 ^ r3);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 123.456) sink[0] = acc;
}

int main() {
    long long* d_cyc; double* d_sink;
    CK(cudaMalloc(&d_cyc, 8 * 1024)); CK(cudaMalloc(&d_sink, 8));
    const int iters = 2000;
    const char* names[] = {"broadcast_all", "distinct_contig", "tileB_2states", "tileC", "2_halves", "4_quarters", "8_same_each_quarter",
                           "8_two_per_quarter", "4_same_each_quarter", "2_interleaved", "tileB_1state", "tileC_shift64", "16_same_halves", "16_pairs",
                           "contig256", "Afrag_ld24", "Bfrag_ld24", "Afrag_ld26", "Bfrag_ld26", "Afrag_ld28", "Bfrag_ld28", "Afrag_packed", "Afrag_ld20", "Bfrag_ld20", "Afrag_ld56", "Bfrag_ld56", "Afrag_ld52", "Bfrag_ld52"};
    printf("{\n");
    for (int width = 8; width <= 16; width += 8)
        for (int warps = 1; warps <= 8; warps *= 8)
            for (int pat = 0; pat < 28; ++pat) {
                for (int rep = 0; rep < 2; ++rep) {
                    if (width == 16) k_lds<16><<<1, 32 * warps, 16384>>>(pat, iters, d_cyc, d_sink);
                    else k_lds<8><<<1, 32 * warps, 16384>>>(pat, iters, d_cyc, d_sink);
                    CK(cudaDeviceSynchronize());
                }
                long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
                // cycles per warp-wide load instruction, per SM (all warps issue concurrently)
                printf(" \"lds%d_w%d_%s\": %.2f,\n", width * 8, warps, names[pat], (double)c / (iters * 16.0 * warps));
            }
    printf(" \"note\": \"SM cycles per warp-wide LDS (1 CTA; w8 = 8 warps sharing the LSU)\"\n}\n");
    return 0;
}
