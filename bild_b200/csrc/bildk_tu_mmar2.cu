// Translation unit of libbild_b200.so: launchers of k_mmar2 (two, four or five warps per filter; GT 5..9) (see bildk_launch.h).
#define BILDK_SATELLITE_TU 1
#include "bildk_launch.h"

using namespace bildk;

template <int GT, int MAXF, bool MX>
static cudaError_t mmar2_launch(const R2Params& rp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&k_mmar2<GT, MAXF, MX, mmar2_nw(GT)>), smem);
        if (e != cudaSuccess) return e;
    }
    k_mmar2<GT, MAXF, MX, mmar2_nw(GT)><<<grid, threads, smem, st>>>(rp);
    return cudaGetLastError();
}
// MAXF = filters per CTA the kernel is compiled for (registers per thread = 65536 / (64 MAXF)): GT = 7 is spill-free only at
// 4 (8 warps, 255 registers; 5 or 6 filters spill, see bildk_mmar2.cuh); the smaller tile grids leave room for more warps.
#define MMAR2_VARIANTS(X) X(5, 4, false) X(6, 4, false) X(7, 4, false) X(5, 4, true) X(6, 4, true) X(7, 4, true) \
                          X(5, 6, false) X(5, 6, true) X(8, 2, false) X(8, 2, true) X(9, 2, false) X(9, 2, true)
bool mmar2_has(int GT, int MAXF, bool MX) {
#define X(G_, F_, M_) if (GT == G_ && MAXF == F_ && MX == M_) return true;
    MMAR2_VARIANTS(X)
#undef X
    return false;
}
cudaError_t mmar2_launch_for(int GT, int MAXF, bool MX, const R2Params& rp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
#define X(G_, F_, M_) if (GT == G_ && MAXF == F_ && MX == M_) return mmar2_launch<G_, F_, M_>(rp, grid, threads, smem, st);
    MMAR2_VARIANTS(X)
#undef X
    return cudaErrorInvalidValue;
}

