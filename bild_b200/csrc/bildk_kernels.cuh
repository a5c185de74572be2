// Device code of the BILD likelihood engine for sm_100a.
//
// One (profile, trajectory, d*-index) "filter" is a Kalman filter over frames
// (/root/reference/bild/src/MSRouse_logL.pyx:203-248):
//     M <- B_s M + G_s            pyx:206-216
//     C <- B_s C B_s + Sig_s      pyx:220-241      (4 N^3 flop: all of the time)
//     valid frame: rank-1 update  pyx:19-90
// Layout of the work:
//   * a filter is owned by a G x G grid of threads (G = ceil(N/TS)); thread (a,b) keeps the TS x TS
//     register tile C[a*TS.., b*TS..] as FP64 accumulators, so the covariance lives in registers
//     while it is being produced and in shared memory while it is an operand;
//   * B_s (all states) is staged into shared memory once per CTA with TMA bulk copies
//     (cp.async.bulk + mbarrier) and shared by every filter of the CTA;
//   * both products are rank-1-update loops over k with 128-bit shared-memory operand loads:
//         P1:  T[i][j]  = sum_k B[k][i] C[k][j]      (B symmetric), written back TRANSPOSED
//         P2:  C'[i][j] = sum_k Tt[k][i] B[k][j] + Sig[i][j]
//     so every operand access is "row k, my column block" - contiguous and conflict-free;
//   * the mean columns ride along in P1 (thread (a,b<d) also accumulates M'[a*TS.., b]);
//   * the measurement update needs C'w: for the sparse measurement vectors BILD uses
//     (end-to-end: w = e_{N-1} - e_0) the owners of the non-zero columns publish them to a small
//     shared buffer, everybody forms Cw, S, K locally and applies C -= K Cw^T to its own tile;
//   * filters narrower than a warp (G*G <= 32) synchronise with __syncwarp only, several per warp.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bildk {

constexpr int NZMAX = 4;     // sparse measurement vectors up to this many non-zeros take the fast path
constexpr int DMAX = 4;      // max spatial dimension
constexpr int MSTRIDE = 4;   // row stride of the mean buffers in shared memory (doubles)
constexpr double LOG_2PI = 1.8378770664093453;   // np.log(2*np.pi), pyx:14

struct KParams {
    // ---- model (padded tile layout: element (i,j) at i*LD + (j/TS)*BS + j%TS, BS = TS + (TS&1))
    int N, D, S, G, LD, NP;
    const double* Bpad;      // [S][NP][LD]
    const double* Sigpad;    // [S][NP][LD]
    const double* C0pad;     // [S][NP][LD]
    const double* Gm;        // [S][N][D]
    const double* M0;        // [S][N][D]
    const double* w;         // [N]
    int hasG;                // any non-zero entry in G
    int nnz;                 // non-zeros of w (fast path when <= NZMAX)
    int wz_idx[NZMAX];
    double wz_val[NZMAX];
    // ---- trajectories: filter p belongs to trajectory traj_of(p); one launch covers `dstar` sub-filters
    //      per profile (blockIdx.y), each with its own localisation error and mean columns
    int n_traj;
    const double* const* x;        // [n_traj] -> (T,D)
    const uint8_t* const* valid;   // [n_traj] -> (T)
    const int* T;                  // [n_traj]
    const int* traj_first;         // [n_traj+1] first profile of every trajectory
    const int* cta_traj;           // [gridDim.x] trajectory of every CTA (nullptr: single trajectory)
    const int* cta_first;          // [gridDim.x] first profile of every CTA   (nullptr: blockIdx.x*FPC)
    int dstar;
    double s2[DMAX];
    int ncols[DMAX];
    int cols[DMAX][DMAX];
    // ---- batch
    int P, K1;
    const int32_t* run_starts;     // [P][K1]
    const uint8_t* run_states;     // [P][K1]
    double* out;                   // [dstar][P]
    // ---- geometry
    int FPC;        // filters per CTA
    int TPFS;       // thread stride between filters of a CTA (>= G*G)
    int b_all;      // 1: all S propagators resident in smem; 0: only the current one (FPC == 1)
    int fstride;    // doubles of shared memory per filter (== 8 mod 16: neighbouring filters use disjoint banks)
    int bstride;    // doubles between the propagators of consecutive states in shared memory (== 8 mod 16)
    const uint16_t* lane_ab;   // [G*G] thread-in-filter -> (a << 8 | b); aligned lane quads own 2x2 tile blocks
};

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (global -> shared)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// stage `bytes` (multiple of 16) with one elected thread, in chunks, onto one mbarrier phase
__device__ __forceinline__ void tma_stage(void* dst, const void* src, size_t bytes, uint64_t* bar) {
    constexpr uint32_t CH = 32768;
    mbar_expect_tx(bar, static_cast<uint32_t>(bytes));
    for (size_t off = 0; off < bytes; off += CH) {
        uint32_t n = static_cast<uint32_t>(bytes - off < CH ? bytes - off : CH);
        tma_load_1d(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, n, bar);
    }
}

// ------------------------------------------------------------------------------------------------
template <int TS>
__device__ __forceinline__ void ld_frag(double (&v)[TS], const double* __restrict__ p) {
#pragma unroll
    for (int i = 0; i + 1 < TS; i += 2) {
        double2 t = *reinterpret_cast<const double2*>(p + i);
        v[i] = t.x;
        v[i + 1] = t.y;
    }
    if (TS & 1) v[TS - 1] = p[TS - 1];
}
template <int TS>
__device__ __forceinline__ void ldg_frag(double (&v)[TS], const double* __restrict__ p) {
#pragma unroll
    for (int i = 0; i + 1 < TS; i += 2) {
        double2 t = __ldg(reinterpret_cast<const double2*>(p + i));
        v[i] = t.x;
        v[i + 1] = t.y;
    }
    if (TS & 1) v[TS - 1] = __ldg(p + TS - 1);
}
template <int TS>
__device__ __forceinline__ void st_frag(double* p, const double (&v)[TS]) {
#pragma unroll
    for (int i = 0; i + 1 < TS; i += 2) *reinterpret_cast<double2*>(p + i) = make_double2(v[i], v[i + 1]);
    if (TS & 1) p[TS - 1] = v[TS - 1];
}

template <bool WS>
__device__ __forceinline__ void fsync() {
    if (WS) __syncwarp(); else __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Tile kernel.  TS: register tile edge; WS: filter fits in a warp (warp-scope barriers);
// DENSEW: measurement vector has more than NZMAX non-zeros.  MAXT bounds the block size.
template <int TS, bool WS, bool DENSEW, int MAXT>
__global__ void __launch_bounds__(MAXT, (WS && TS <= 5) ? 4 : 1) k_tile(const __grid_constant__ KParams p) {
    constexpr int BS = TS + (TS & 1);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    double* Bsm = reinterpret_cast<double*>(smem_raw + 16);

    const int tid = threadIdx.x;
    const int e = blockIdx.y;
    const int N = p.N, D = p.D, G = p.G, LD = p.LD, NP = p.NP;
    const int TPF = G * G;
    const size_t matd = static_cast<size_t>(NP) * LD;   // doubles per padded matrix

    const int fl = tid / p.TPFS;
    const int l = tid - fl * p.TPFS;
    const int tj = p.cta_traj ? p.cta_traj[blockIdx.x] : 0;
    const int first = p.cta_first ? p.cta_first[blockIdx.x] : blockIdx.x * p.FPC;
    const int pend = p.traj_first[tj + 1];
    const int pidx = first + fl;
    const bool alive = (fl < p.FPC) && (pidx < pend) && (l < TPF);
    // Lane -> tile map: an aligned quad of lanes owns a 2x2 block of tiles, so every operand load sees at
    // most 2 distinct addresses per quad (measured with tools/lds_patterns.cu: 4 distinct per quad doubles
    // the shared-memory wavefronts of a warp-wide LDS.64/LDS.128)
    const int ab = alive ? p.lane_ab[l] : 0;
    const int a = ab >> 8;
    const int b = ab & 0xff;

    // ---- stage propagators with TMA
    int s_loaded = -1;
    uint32_t bphase = 0;
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (p.b_all) {
        if (tid == 0) {
            mbar_expect_tx(mbar, static_cast<uint32_t>(matd * p.S * sizeof(double)));
            for (int st = 0; st < p.S; ++st) {
                constexpr uint32_t CH = 32768;
                const size_t bytes = matd * sizeof(double);
                for (size_t off = 0; off < bytes; off += CH)
                    tma_load_1d(reinterpret_cast<char*>(Bsm + static_cast<size_t>(st) * p.bstride) + off,
                                reinterpret_cast<const char*>(p.Bpad + matd * st) + off,
                                static_cast<uint32_t>(bytes - off < CH ? bytes - off : CH), mbar);
            }
        }
    }

    // ---- per-filter shared buffers
    double* fbase = Bsm + (p.b_all ? static_cast<size_t>(p.bstride) * p.S : matd) + static_cast<size_t>(fl < p.FPC ? fl : 0) * p.fstride;
    double* Cb = fbase;                              // [NP][LD]   covariance / transposed intermediate
    double* Mb0 = Cb + matd;                         // [NP][MSTRIDE] mean used by the next propagation
    double* Mb1 = Mb0 + NP * MSTRIDE;                // [NP][MSTRIDE] prior mean of a valid frame
    double* colb = Mb1 + NP * MSTRIDE;               // sparse: [NZMAX][NP] columns of C'; dense: [NP] Cw | [NP] w
    double* llb = colb + (DENSEW ? 2 : NZMAX) * NP;  // [DMAX]

    const int T = p.T[tj];
    const double* __restrict__ xg = p.x[tj];
    const uint8_t* __restrict__ vg = p.valid[tj];

    const int ncols = p.ncols[e];
    const bool hasM = alive && (b < ncols);
    const int q = hasM ? b : 0;
    const int col = hasM ? p.cols[e][q] : 0;
    const double s2 = p.s2[e];
    double ll = 0.0;

    // run-length profile cursor (amis.py:685-693 semantics)
    // (idle threads shadow the CTA's first filter so that CTA-uniform decisions stay uniform)
    const int pcur = (fl < p.FPC && pidx < pend) ? pidx : first;
    const int32_t* rs = p.run_starts + static_cast<size_t>(pcur) * p.K1;
    const uint8_t* rt = p.run_states + static_cast<size_t>(pcur) * p.K1;
    int r_cur = 0;
    int s = rt[0];
    int next_sw = (p.K1 > 1) ? rs[1] : 0x7fffffff;

    if (DENSEW && alive) {
        for (int i = l; i < NP; i += TPF) {
            colb[i] = 0.0;
            colb[NP + i] = (i < N) ? p.w[i] : 0.0;
        }
    }

    double acc[TS][TS];
    double macc[TS];

    if (p.b_all) mbar_wait(mbar, bphase);

    for (int t = 0; t < T; ++t) {
        while (t >= next_sw) {   // enter the run that contains frame t (empty runs vanish)
            ++r_cur;
            s = rt[r_cur];
            next_sw = (r_cur + 1 < p.K1) ? rs[r_cur + 1] : 0x7fffffff;
        }
        const bool is_valid = vg[t] != 0;

        if (t == 0) {
            // steady state of the state profile[0] (pyx:160-163)
            if (alive) {
                const double* C0 = p.C0pad + matd * s + static_cast<size_t>(a * TS) * LD + b * BS;
#pragma unroll
                for (int r = 0; r < TS; ++r) ldg_frag<TS>(acc[r], C0 + r * LD);
#pragma unroll
                for (int r = 0; r < TS; ++r) {
                    const int row = a * TS + r;
                    macc[r] = (hasM && row < N) ? __ldg(p.M0 + (static_cast<size_t>(s) * N + row) * D + col) : 0.0;
                }
            }
        } else {
            if (!p.b_all && s != s_loaded) {   // single resident propagator (large N, FPC == 1): swap it
                if (tid == 0) tma_stage(Bsm, p.Bpad + matd * s, matd * sizeof(double), mbar);
                mbar_wait(mbar, bphase);
                bphase ^= 1;
                s_loaded = s;
            }
            const double* Bs = Bsm + (p.b_all ? static_cast<size_t>(p.bstride) * s : 0);
            // ---------------- P1: T = B C (and M' = B M), operands: row k of B / C / M
            if (alive) {
#pragma unroll
                for (int r = 0; r < TS; ++r) {
                    macc[r] = 0.0;
#pragma unroll
                    for (int c = 0; c < TS; ++c) acc[r][c] = 0.0;
                }
                const double* Bp = Bs + a * BS;
                const double* Cp = Cb + b * BS;
                const double* Mp = Mb0 + q;
#pragma unroll 2
                for (int k = 0; k < N; ++k) {
                    double bf[TS], cf[TS];
                    ld_frag<TS>(bf, Bp + k * LD);
                    ld_frag<TS>(cf, Cp + k * LD);
                    const double m = Mp[k * MSTRIDE];
#pragma unroll
                    for (int r = 0; r < TS; ++r) {
#pragma unroll
                        for (int c = 0; c < TS; ++c) acc[r][c] = fma(bf[r], cf[c], acc[r][c]);
                        macc[r] = fma(bf[r], m, macc[r]);
                    }
                }
            }
            fsync<WS>();   // A: everybody finished reading C and M
            if (alive) {
                // T written transposed: Tt[j][i] = T[i][j]
#pragma unroll
                for (int c = 0; c < TS; ++c) {
                    double v[TS];
#pragma unroll
                    for (int r = 0; r < TS; ++r) v[r] = acc[r][c];
                    st_frag<TS>(Cb + static_cast<size_t>(b * TS + c) * LD + a * BS, v);
                }
                if (p.hasG && hasM) {
#pragma unroll
                    for (int r = 0; r < TS; ++r) {
                        const int row = a * TS + r;
                        if (row < N) macc[r] += __ldg(p.Gm + (static_cast<size_t>(s) * N + row) * D + col);
                    }
                }
            }
        }
        // prior mean: straight to the propagation buffer if this frame has no data
        if (hasM) {
            double* Mdst = is_valid ? Mb1 : Mb0;
#pragma unroll
            for (int r = 0; r < TS; ++r) Mdst[(a * TS + r) * MSTRIDE + q] = macc[r];
        }
        if (t > 0) {
            fsync<WS>();   // B: Tt complete
            // ---------------- P2: C' = T B + Sig
            if (alive) {
                const double* Sg = p.Sigpad + matd * s + static_cast<size_t>(a * TS) * LD + b * BS;
#pragma unroll
                for (int r = 0; r < TS; ++r) ldg_frag<TS>(acc[r], Sg + r * LD);
                const double* Bs = Bsm + (p.b_all ? static_cast<size_t>(p.bstride) * s : 0);
                const double* Tp = Cb + a * BS;
                const double* Bp = Bs + b * BS;
#pragma unroll 2
                for (int k = 0; k < N; ++k) {
                    double tf[TS], bf[TS];
                    ld_frag<TS>(tf, Tp + k * LD);
                    ld_frag<TS>(bf, Bp + k * LD);
#pragma unroll
                    for (int r = 0; r < TS; ++r)
#pragma unroll
                        for (int c = 0; c < TS; ++c) acc[r][c] = fma(tf[r], bf[c], acc[r][c]);
                }
            }
        }

        // ---------------- measurement update (pyx:19-90) on frames with data
        if (!DENSEW) {
            if (is_valid && alive) {
#pragma unroll
                for (int z = 0; z < NZMAX; ++z) {
                    if (z < p.nnz) {
                        const int jz = p.wz_idx[z];
                        if (jz / TS == b) {
                            const int cz = jz - b * TS;
#pragma unroll
                            for (int c = 0; c < TS; ++c)
                                if (c == cz) {
#pragma unroll
                                    for (int r = 0; r < TS; ++r) colb[z * NP + a * TS + r] = acc[r][c];
                                }
                        }
                    }
                }
            }
            fsync<WS>();   // C: Tt no longer needed; published columns visible
        } else {
            fsync<WS>();   // C
            if (is_valid) {
                if (alive) {
#pragma unroll
                    for (int r = 0; r < TS; ++r) st_frag<TS>(Cb + static_cast<size_t>(a * TS + r) * LD + b * BS, acc[r]);
                }
                fsync<WS>();
                if (alive) {
                    const double* wv = colb + NP;
                    for (int i = l; i < N; i += TPF) {   // Cw = C' w, one row per thread
                        const double* row = Cb + static_cast<size_t>(i) * LD;
                        double sum = 0.0;
                        for (int j = 0; j < N; ++j) sum = fma(row[(j / TS) * BS + (j % TS)], wv[j], sum);
                        colb[i] = sum;
                    }
                }
                fsync<WS>();
            }
        }
        if (is_valid && alive) {
            double cwr[TS], cwc[TS];
            double S = s2;
            if (!DENSEW) {
#pragma unroll
                for (int r = 0; r < TS; ++r) cwr[r] = 0.0;
#pragma unroll
                for (int c = 0; c < TS; ++c) cwc[c] = 0.0;
#pragma unroll
                for (int z = 0; z < NZMAX; ++z) {
                    if (z < p.nnz) {
                        const double wz = p.wz_val[z];
                        const double* cz = colb + z * NP;
#pragma unroll
                        for (int r = 0; r < TS; ++r) cwr[r] = fma(wz, cz[a * TS + r], cwr[r]);
#pragma unroll
                        for (int c = 0; c < TS; ++c) cwc[c] = fma(wz, cz[b * TS + c], cwc[c]);
                    }
                }
#pragma unroll
                for (int z = 0; z < NZMAX; ++z) {
                    if (z < p.nnz) {
                        double cw_j = 0.0;   // (C' w)[idx_z]
#pragma unroll
                        for (int y = 0; y < NZMAX; ++y)
                            if (y < p.nnz) cw_j = fma(p.wz_val[y], colb[y * NP + p.wz_idx[z]], cw_j);
                        S = fma(p.wz_val[z], cw_j, S);
                    }
                }
            } else {
                const double* wv = colb + NP;
#pragma unroll
                for (int r = 0; r < TS; ++r) cwr[r] = colb[a * TS + r];
#pragma unroll
                for (int c = 0; c < TS; ++c) cwc[c] = colb[b * TS + c];
                double dot = 0.0;
                for (int i = 0; i < N; ++i) dot = fma(colb[i], wv[i], dot);
                S += dot;
            }
            const double Sinv = 1.0 / S;   // pyx:63
            double kr[TS];
#pragma unroll
            for (int r = 0; r < TS; ++r) kr[r] = cwr[r] * Sinv;   // pyx:66-67
#pragma unroll
            for (int r = 0; r < TS; ++r)
#pragma unroll
                for (int c = 0; c < TS; ++c) acc[r][c] = fma(-kr[r], cwc[c], acc[r][c]);   // pyx:71-75
            if (hasM) {
                double wm = 0.0;   // w . M'[:, col]
                if (!DENSEW) {
#pragma unroll
                    for (int z = 0; z < NZMAX; ++z)
                        if (z < p.nnz) wm = fma(p.wz_val[z], Mb1[p.wz_idx[z] * MSTRIDE + q], wm);
                } else {
                    const double* wv = colb + NP;
                    for (int i = 0; i < N; ++i) wm = fma(wv[i], Mb1[i * MSTRIDE + q], wm);
                }
                const double xmm = __ldg(xg + static_cast<size_t>(t) * D + col) - wm;   // pyx:79
#pragma unroll
                for (int r = 0; r < TS; ++r) {
                    macc[r] = fma(kr[r], xmm, macc[r]);                                    // pyx:82-85
                    Mb0[(a * TS + r) * MSTRIDE + q] = macc[r];
                }
                if (a == 0) ll += -0.5 * (xmm * xmm * Sinv - log(Sinv) + LOG_2PI);         // pyx:88
            }
        }
        // posterior covariance becomes the operand of the next propagation
        if (alive && t + 1 < T) {
#pragma unroll
            for (int r = 0; r < TS; ++r) st_frag<TS>(Cb + static_cast<size_t>(a * TS + r) * LD + b * BS, acc[r]);
        }
        fsync<WS>();   // D
    }

    // ---- sum the per-dimension partial log-likelihoods in dimension order and write out
    if (hasM && a == 0) llb[q] = ll;
    fsync<WS>();
    if (alive && l == 0) {
        double tot = 0.0;
        for (int c = 0; c < ncols; ++c) tot += llb[c];
        p.out[static_cast<size_t>(e) * p.P + pidx] = tot;
    }
}

// ------------------------------------------------------------------------------------------------
// Catch-all kernel for shapes the tile kernel cannot hold on chip (N too large for shared memory /
// registers, or N < d): one CTA per filter, covariance and intermediate in a global workspace
// (L2-resident), plain unpadded layouts.  Correct for every N, d; not tuned.
struct GParams {
    int N, D, S;
    const double *B, *Sig, *C0, *Gm, *M0, *w;   // unpadded
    int n_traj;
    const double* const* x;
    const uint8_t* const* valid;
    const int* T;
    const int* traj_first;
    const int* prof_traj;      // [P] trajectory of every profile
    int dstar;
    double s2[DMAX];
    int ncols[DMAX];
    int cols[DMAX][DMAX];
    int P, K1;
    const int32_t* run_starts;
    const uint8_t* run_states;
    double* out;               // [dstar][P]
    double* work;              // [gridDim.x*gridDim.y][2*N*N + 3*N*D + 2*N]
};

// The non-template kernels below are compiled in the main translation unit only (bildk.cu); the launcher translation units
// (bildk_tu_*.cu, BILDK_SATELLITE_TU) include this header for the shared helpers and parameter structs.
#ifndef BILDK_SATELLITE_TU
__global__ void __launch_bounds__(256) k_generic(const __grid_constant__ GParams p) {
    const int N = p.N, D = p.D;
    const int e = blockIdx.y;
    const int nthr = blockDim.x, tid = threadIdx.x;
    const size_t wsz = 2 * static_cast<size_t>(N) * N + 3 * static_cast<size_t>(N) * D + 2 * N;
    double* W = p.work + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * wsz;
    double* C = W;
    double* Tm = C + static_cast<size_t>(N) * N;
    double* M = Tm + static_cast<size_t>(N) * N;
    double* Mn = M + N * D;
    double* Cw = Mn + N * D;
    double* Kv = Cw + N;
    __shared__ double sh_S, sh_ll;
    __shared__ double sh_xmm[DMAX];

    for (int pidx = blockIdx.x; pidx < p.P; pidx += gridDim.x) {
        const int tj = p.prof_traj ? p.prof_traj[pidx] : 0;
        const int T = p.T[tj];
        const double* xg = p.x[tj];
        const uint8_t* vg = p.valid[tj];
        const int32_t* rs = p.run_starts + static_cast<size_t>(pidx) * p.K1;
        const uint8_t* rt = p.run_states + static_cast<size_t>(pidx) * p.K1;
        int r_cur = 0, s = rt[0];
        int next_sw = (p.K1 > 1) ? rs[1] : 0x7fffffff;
        const int ncols = p.ncols[e];
        const double s2 = p.s2[e];
        if (tid == 0) sh_ll = 0.0;
        for (int t = 0; t < T; ++t) {
            while (t >= next_sw) {
                ++r_cur;
                s = rt[r_cur];
                next_sw = (r_cur + 1 < p.K1) ? rs[r_cur + 1] : 0x7fffffff;
            }
            const double* B = p.B + static_cast<size_t>(s) * N * N;
            if (t == 0) {
                for (int i = tid; i < N * N; i += nthr) C[i] = p.C0[static_cast<size_t>(s) * N * N + i];
                for (int i = tid; i < N * ncols; i += nthr) {
                    const int row = i / ncols, c = i % ncols;
                    M[row * D + c] = p.M0[(static_cast<size_t>(s) * N + row) * D + p.cols[e][c]];
                }
            } else {
                for (int o = tid; o < N * N; o += nthr) {   // T = B C
                    const int i = o / N, j = o % N;
                    double sum = 0.0;
                    for (int k = 0; k < N; ++k) sum = fma(B[static_cast<size_t>(i) * N + k], C[static_cast<size_t>(k) * N + j], sum);
                    Tm[o] = sum;
                }
                for (int o = tid; o < N * ncols; o += nthr) {
                    const int i = o / ncols, c = o % ncols;
                    double sum = 0.0;
                    for (int k = 0; k < N; ++k) sum = fma(B[static_cast<size_t>(i) * N + k], M[k * D + c], sum);
                    Mn[i * D + c] = sum + p.Gm[(static_cast<size_t>(s) * N + i) * D + p.cols[e][c]];
                }
                __syncthreads();
                for (int o = tid; o < N * N; o += nthr) {   // C = T B + Sig
                    const int i = o / N, j = o % N;
                    double sum = p.Sig[static_cast<size_t>(s) * N * N + o];
                    for (int k = 0; k < N; ++k) sum = fma(Tm[static_cast<size_t>(i) * N + k], B[static_cast<size_t>(k) * N + j], sum);
                    C[o] = sum;
                }
                for (int o = tid; o < N * ncols; o += nthr) M[(o / ncols) * D + o % ncols] = Mn[(o / ncols) * D + o % ncols];
            }
            __syncthreads();
            if (vg[t]) {
                for (int i = tid; i < N; i += nthr) {
                    double sum = 0.0;
                    for (int j = 0; j < N; ++j) sum = fma(C[static_cast<size_t>(i) * N + j], p.w[j], sum);
                    Cw[i] = sum;
                }
                __syncthreads();
                if (tid == 0) {
                    double dot = 0.0;
                    for (int i = 0; i < N; ++i) dot = fma(Cw[i], p.w[i], dot);
                    sh_S = s2 + dot;
                }
                if (tid < ncols) {
                    double wm = 0.0;
                    for (int i = 0; i < N; ++i) wm = fma(p.w[i], M[i * D + tid], wm);
                    sh_xmm[tid] = xg[static_cast<size_t>(t) * D + p.cols[e][tid]] - wm;
                }
                __syncthreads();
                const double Sinv = 1.0 / sh_S;
                for (int i = tid; i < N; i += nthr) Kv[i] = Cw[i] * Sinv;
                __syncthreads();
                for (int o = tid; o < N * N; o += nthr) C[o] = fma(-Kv[o / N], Cw[o % N], C[o]);
                for (int o = tid; o < N * ncols; o += nthr) M[(o / ncols) * D + o % ncols] += Kv[o / ncols] * sh_xmm[o % ncols];
                if (tid == 0) {
                    double acc = sh_ll;
                    for (int c = 0; c < ncols; ++c) acc += -0.5 * (sh_xmm[c] * sh_xmm[c] * Sinv - log(Sinv) + LOG_2PI);
                    sh_ll = acc;
                }
                __syncthreads();
            }
        }
        __syncthreads();
        if (tid == 0) p.out[static_cast<size_t>(e) * p.P + pidx] = sh_ll;
        __syncthreads();
    }
}

// out[p] = sum_e part[e][p]  (d* > 1: anisotropic localisation error)
__global__ void k_sum_parts(const double* __restrict__ part, double* __restrict__ out, int P, int dstar) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) {
        double s = 0.0;
        for (int e = 0; e < dstar; ++e) s += part[static_cast<size_t>(e) * P + i];
        out[i] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// AMIS weight normalisation (amis.py:843-845, 878-900): fixed-order, single-CTA, warp-shuffle tree.
// The reduction order depends only on n, never on the grid, so every rank gets identical bits.
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Multi-CTA version: ONE launch of a thread-block cluster of AW_CLUSTER CTAs.  Every CTA owns a contiguous chunk of the
// ensemble; the per-CTA partials of each pass are exchanged through distributed shared memory (cluster.map_shared_rank)
// and combined by every CTA in rank order, so the result is a function of n alone - identical bits on every rank of
// a multi-GPU run - while the three passes over the (gathered) ensemble are spread over AW_CLUSTER SMs.
// (Round 1 used one CTA: 10.7 us at n = 4096 but ~8x that for the 32768-element vector of an 8-GPU step.)
constexpr int AW_CLUSTER = 8;
__global__ void __launch_bounds__(1024) k_amis_weights(int n, const double* __restrict__ logL,
                                                       const double* __restrict__ logdelta,
                                                       const double* __restrict__ curlp, double log_nsteps,
                                                       double* __restrict__ log_w, double* __restrict__ stats) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = static_cast<int>(cluster.block_rank()), ncta = static_cast<int>(cluster.num_blocks());
    __shared__ double red[32][2];
    __shared__ double part[4];   // this CTA's partials (max, sum w, sum (w - mean)^2, nansum), read by the peers
    __shared__ double bc[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    const int chunk = (n + ncta - 1) / ncta;
    const int lo = crank * chunk, hi = min(n, lo + chunk);
    // pass 1: log weights and their maximum
    double mx = -INFINITY;
    for (int i = lo + tid; i < hi; i += blockDim.x) {
        const double lw = logL[i] - logdelta[i] + log_nsteps;
        if (log_w) log_w[i] = lw;
        mx = fmax(mx, lw);
    }
    mx = warp_max(mx);
    if (lane == 0) red[wid][0] = mx;
    __syncthreads();
    if (wid == 0) {
        mx = warp_max((lane < nw) ? red[lane][0] : -INFINITY);
        if (lane == 0) part[0] = mx;
    }
    cluster.sync();
    if (tid == 0) {
        double v = -INFINITY;
        for (int r = 0; r < ncta; ++r) v = fmax(v, *cluster.map_shared_rank(&part[0], r));
        bc[0] = v;
    }
    __syncthreads();
    mx = bc[0];
    // pass 2: sum of the shifted weights
    double s1 = 0.0;
    for (int i = lo + tid; i < hi; i += blockDim.x) s1 += exp(logL[i] - logdelta[i] + log_nsteps - mx);
    s1 = warp_sum(s1);
    if (lane == 0) red[wid][0] = s1;
    __syncthreads();
    if (wid == 0) {
        s1 = warp_sum((lane < nw) ? red[lane][0] : 0.0);
        if (lane == 0) part[1] = s1;
    }
    cluster.sync();
    if (tid == 0) {
        double v = 0.0;
        for (int r = 0; r < ncta; ++r) v += *cluster.map_shared_rank(&part[1], r);   // rank order: same bits in every CTA
        bc[1] = v;
    }
    __syncthreads();
    s1 = bc[1];
    const double mean = s1 / n;
    // pass 3: centred second moment (what scipy.stats.sem needs) and the KL numerator
    double ssd = 0.0, s3 = 0.0;
    for (int i = lo + tid; i < hi; i += blockDim.x) {
        const double wo = exp(logL[i] - logdelta[i] + log_nsteps - mx);
        const double dv = wo - mean;
        ssd = fma(dv, dv, ssd);
        const double term = wo * (logL[i] - curlp[i]);
        if (term == term) s3 += term;   // nansum (amis.py:885-895)
    }
    ssd = warp_sum(ssd);
    s3 = warp_sum(s3);
    if (lane == 0) { red[wid][0] = ssd; red[wid][1] = s3; }
    __syncthreads();
    if (wid == 0) {
        ssd = warp_sum((lane < nw) ? red[lane][0] : 0.0);
        s3 = warp_sum((lane < nw) ? red[lane][1] : 0.0);
        if (lane == 0) { part[2] = ssd; part[3] = s3; }
    }
    cluster.sync();
    if (crank == 0 && tid == 0) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < ncta; ++r) { a += *cluster.map_shared_rank(&part[2], r); b += *cluster.map_shared_rank(&part[3], r); }
        stats[0] = mx; stats[1] = s1; stats[2] = a; stats[3] = b;
    }
    cluster.sync();   // no CTA may exit while rank 0 still reads its shared memory
}

// ------------------------------------------------------------------------------------------------
// Marginal posterior of the state at every frame from a weighted ensemble of run-length profiles
// (FixedkSampler.log_marginal_posterior, /root/reference/bild/amis.py:942-972, which materialises an
// (n, S, T) boolean tensor on the host): out[s][t] = log sum_{i: state_i(t) = s} w_i, normalised over s.
// One CTA per frame; for every state a fixed-order (thread-strided, warp-shuffle, cross-warp) sum, so the
// result does not depend on the launch geometry of other frames and is reproducible.  The state of every sample at
// this frame is looked up ONCE (a scan of its run starts) and cached as one byte in dynamic shared memory
// (`cache_n` = n when the launch provided n bytes, else 0: look-up on the fly), not once per state and pass.
// The log-sum-exp of a state is shifted by the maximum over ITS OWN samples - what scipy >= 1.15 does for
// logsumexp(a, b=mask) (masked entries are set to -inf before the shift); older scipy releases, the ones the
// reference's `numpy < 2` pin allows, shift by the global maximum, so states more than ~745 below the best sample
// underflow to -inf there and stay finite here.
__global__ void __launch_bounds__(256) k_marginal_posterior(int n, int K1, int T, int S, const int32_t* __restrict__ starts,
                                                            const uint8_t* __restrict__ states, const double* __restrict__ log_w,
                                                            double* __restrict__ out, int cache_n) {
    extern __shared__ uint8_t st_cache[];
    __shared__ double red[8];
    __shared__ double bc;
    __shared__ double lse[256];
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    auto lookup = [&](int i) {
        const int32_t* rs = starts + static_cast<size_t>(i) * K1;
        int r = 0;
        for (int q = 1; q < K1; ++q) r = (rs[q] <= t) ? q : r;   // last run that has started (empty runs vanish)
        return static_cast<int>(states[static_cast<size_t>(i) * K1 + r]);
    };
    const bool cached = cache_n >= n;
    if (cached) {
        for (int i = tid; i < n; i += blockDim.x) st_cache[i] = static_cast<uint8_t>(lookup(i));
        __syncthreads();
    }
    auto state_at = [&](int i) { return cached ? static_cast<int>(st_cache[i]) : lookup(i); };
    for (int s = 0; s < S; ++s) {
        double mx = -INFINITY;
        for (int i = tid; i < n; i += blockDim.x)
            if (state_at(i) == s) mx = fmax(mx, log_w[i]);
        mx = warp_max(mx);
        if (lane == 0) red[wid] = mx;
        __syncthreads();
        if (tid == 0) {
            double v = -INFINITY;
            for (int k = 0; k < nw; ++k) v = fmax(v, red[k]);
            bc = v;
        }
        __syncthreads();
        mx = bc;
        double acc = 0.0;
        if (mx > -INFINITY && mx < INFINITY)
            for (int i = tid; i < n; i += blockDim.x)
                if (state_at(i) == s) acc += exp(log_w[i] - mx);
        acc = warp_sum(acc);
        __syncthreads();
        if (lane == 0) red[wid] = acc;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            for (int k = 0; k < nw; ++k) v += red[k];
            lse[s] = (mx > -INFINITY && mx < INFINITY) ? log(v) + mx : mx;
        }
        __syncthreads();
    }
    if (tid == 0) {   // normalise over the states (amis.py:972)
        double top = -INFINITY;
        for (int s = 0; s < S; ++s) top = fmax(top, lse[s]);
        double all = 0.0;
        if (top > -INFINITY && top < INFINITY)
            for (int s = 0; s < S; ++s) all += exp(lse[s] - top);
        const double norm = (top > -INFINITY && top < INFINITY) ? log(all) + top : top;
        for (int s = 0; s < S; ++s) out[static_cast<size_t>(s) * T + t] = lse[s] - norm;
    }
}

#endif  // BILDK_SATELLITE_TU

}  // namespace bildk
