// Translation unit of libbild_b200.so: launchers of k_mmar (one warp per filter) and k_mmarb (see bildk_launch.h).
#define BILDK_SATELLITE_TU 1
#include "bildk_launch.h"

using namespace bildk;

template <int GT, int NB, bool MX>
static cudaError_t mmar_launch(const RParams& rp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&k_mmar<GT, NB, MX>), smem);
        if (e != cudaSuccess) return e;
    }
    k_mmar<GT, NB, MX><<<grid, threads, smem, st>>>(rp);
    return cudaGetLastError();
}
// compiled register budgets (resident 4-warp CTAs per SM = warps per scheduler); MX = mean in an extra row block
#define MMAR_VARIANTS(X) X(1, 4, false) X(1, 7, false) X(2, 4, false) X(2, 5, false) X(2, 7, false) X(3, 4, false) X(3, 5, false) X(3, 7, false) \
                         X(4, 3, false) X(4, 4, false) X(4, 5, false) \
                         X(1, 4, true) X(2, 4, true) X(3, 4, true) X(3, 3, true) X(4, 3, true)
bool mmar_has(int GT, int NB, bool MX) {
#define X(G_, N_, M_) if (GT == G_ && NB == N_ && MX == M_) return true;
    MMAR_VARIANTS(X)
#undef X
    return false;
}
cudaError_t mmar_launch_for(int GT, int NB, bool MX, const RParams& rp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
#define X(G_, N_, M_) if (GT == G_ && NB == N_ && MX == M_) return mmar_launch<G_, N_, M_>(rp, grid, threads, smem, st);
    MMAR_VARIANTS(X)
#undef X
    return cudaErrorInvalidValue;
}

template <int GT, int NB, int RB>
static cudaError_t mmarb_launch(const RParams& rp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&k_mmarb<GT, NB, RB>), smem);
        if (e != cudaSuccess) return e;
    }
    k_mmarb<GT, NB, RB><<<grid, threads, smem, st>>>(rp);
    return cudaGetLastError();
}
#define MMARB_VARIANTS(X) X(2, 4, 1) X(2, 4, 2) X(3, 4, 1) X(3, 4, 2) X(4, 3, 1) X(4, 3, 2)
bool mmarb_has(int GT, int NB, int RB) {
#define X(G_, N_, R_) if (GT == G_ && NB == N_ && RB == R_) return true;
    MMARB_VARIANTS(X)
#undef X
    return false;
}
cudaError_t mmarb_launch_for(int GT, int NB, int RB, const RParams& rp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
#define X(G_, N_, R_) if (GT == G_ && NB == N_ && RB == R_) return mmarb_launch<G_, N_, R_>(rp, grid, threads, smem, st);
    MMARB_VARIANTS(X)
#undef X
    return cudaErrorInvalidValue;
}

