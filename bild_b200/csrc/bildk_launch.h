// Launchers of the register-chained kernel families, one translation unit per family (bildk_tu_*.cu), used by bildk.cu.
#pragma once
#include <cuda_runtime.h>

#include "bildk_mmar2.cuh"
#include "bildk_mmarb.cuh"

cudaError_t ensure_dyn_smem(const void* kernel, size_t smem);   // bildk.cu: per-device cache of the dynamic shared-memory attribute

bool mmar_has(int GT, int NB, bool MX);
cudaError_t mmar_launch_for(int GT, int NB, bool MX, const bildk::RParams& rp, dim3 grid, int threads, size_t smem, cudaStream_t st);
bool mmarb_has(int GT, int NB, int RB);
cudaError_t mmarb_launch_for(int GT, int NB, int RB, const bildk::RParams& rp, dim3 grid, int threads, size_t smem, cudaStream_t st);
constexpr int mmar2_nw(int GT) { return GT == 9 ? 5 : GT >= 8 ? 4 : 2; }   // warps per filter
bool mmar2_has(int GT, int MAXF, bool MX);
cudaError_t mmar2_launch_for(int GT, int MAXF, bool MX, const bildk::R2Params& rp, dim3 grid, int threads, size_t smem, cudaStream_t st);
// k_mmar8: one translation unit per instantiation (bildk_tu_mmar8.cu with -DBILDK_MMAR8_GT / -DBILDK_MMAR8_MX)
#define BILDK_MMAR8_DECL(gt, mx) cudaError_t mmar8_launch_##gt##_##mx(const bildk::R2Params& rp, dim3 grid, size_t smem, cudaStream_t st);
BILDK_MMAR8_DECL(10, 0) BILDK_MMAR8_DECL(10, 1) BILDK_MMAR8_DECL(11, 0) BILDK_MMAR8_DECL(11, 1)
BILDK_MMAR8_DECL(12, 0) BILDK_MMAR8_DECL(12, 1) BILDK_MMAR8_DECL(13, 0) BILDK_MMAR8_DECL(13, 1)
#undef BILDK_MMAR8_DECL
inline cudaError_t mmar8_launch_for(int GT, bool MX, const bildk::R2Params& rp, dim3 grid, size_t smem, cudaStream_t st) {
    switch (GT * 2 + (MX ? 1 : 0)) {
        case 20: return mmar8_launch_10_0(rp, grid, smem, st);
        case 21: return mmar8_launch_10_1(rp, grid, smem, st);
        case 22: return mmar8_launch_11_0(rp, grid, smem, st);
        case 23: return mmar8_launch_11_1(rp, grid, smem, st);
        case 24: return mmar8_launch_12_0(rp, grid, smem, st);
        case 25: return mmar8_launch_12_1(rp, grid, smem, st);
        case 26: return mmar8_launch_13_0(rp, grid, smem, st);
        case 27: return mmar8_launch_13_1(rp, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}
