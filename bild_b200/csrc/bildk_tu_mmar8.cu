// Translation unit of libbild_b200.so: the launcher of ONE instantiation of k_mmar8, selected by -DBILDK_MMAR8_GT=<10..13>
// -DBILDK_MMAR8_MX=<0|1> (bild_b200/build.py compiles this file once per instantiation, all in parallel: NVVM needs about a
// minute for each of them).  See bildk_launch.h.
#define BILDK_SATELLITE_TU 1
#ifndef BILDK_MMAR8_GT   // a plain `nvcc -c` of this file compiles the BASELINE configs[2] instantiation
#define BILDK_MMAR8_GT 13
#endif
#ifndef BILDK_MMAR8_MX
#define BILDK_MMAR8_MX 0
#endif
#include "bildk_launch.h"

using namespace bildk;

#define BILDK_CAT3(a, b, c) a##b##_##c
#define BILDK_MMAR8_NAME(gt, mx) BILDK_CAT3(mmar8_launch_, gt, mx)

cudaError_t BILDK_MMAR8_NAME(BILDK_MMAR8_GT, BILDK_MMAR8_MX)(const R2Params& rp, dim3 grid, size_t smem, cudaStream_t st) {
    constexpr int GT = BILDK_MMAR8_GT;
    constexpr bool MX = BILDK_MMAR8_MX != 0;
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&k_mmar8<GT, MX>), smem);
        if (e != cudaSuccess) return e;
    }
    k_mmar8<GT, MX><<<grid, 256, smem, st>>>(rp);
    return cudaGetLastError();
}
