// k_mmact - k_mmac (one CTA per filter, covariance in shared memory, 56 < N <= 112, mean in the padding columns) with
// the work of EVERY phase balanced over the warps (BASELINE configs[2], N = 100: GT = 13 tile columns, 16 warps).
//
// ncu of k_mmac<13> on configs[2] (profiles/r01_ncu_c3_mmac_v9.txt): the tensor pipe is 65 % active and 37 % of the
// stall samples are CTA barriers.  Per frame the phases are separated by barriers, so each phase must be balanced on
// its own, and with one warp per tile column it is not:
//   P1   every column costs GT tiles, but 13 column warps sit 4/3/3/3 on the four schedulers  -> for GT = 4k + 1 helper
//        warps (as in k_mmac): warps 13, 14, 15 take the first GT/4 tile rows of the columns of warps 0, 4, 8:
//        43/42/42/42 tiles (NH = 3; other GT run without helpers, NH = 0);
//   P2   column c has c + 1 upper tiles: the warp of the last column works 13 x longer than that of the first and
//        runs alone at the end of the phase; the rank-1 update and the write-back inherit the same imbalance while
//        the tensor pipe idles.  Every upper tile (ti, c) is independent given the published columns of C', so here
//        the 91 upper tiles are dealt out as SLOTS, 5 or 6 per warp (host table, 23/23/23/22 per scheduler): a warp
//        computes, publishes, updates and writes back its slots, whatever column they belong to.
// A warp's slots form at most two SEGMENTS (consecutive tile rows of one column) whose tiles share the B fragment in
// P2, as a column warp's do.  The prior mean rows that the innovation needs are read after the P1
// barrier and M+ is written after the publish barrier, so the warps that hold tiles of the mean-carrying last
// column need no extra synchronisation.  Arithmetic per tile is k_mmac's.
#pragma once
#include "bildk_mma.cuh"

namespace bildk {

constexpr int MMACT_MAXS = 7;   // most slots (upper tiles) one warp can hold

struct CTParams {
    MParams m;
    int b_all;                    // all S propagators resident
    unsigned char nslot[16];      // slots of every warp (<= MMACT_MAXS): segment A = slots [0, nsegA), segment B the rest
    unsigned char nsegA[16];      // ... each segment = consecutive tile rows of one tile column
    unsigned char slot_ti[16][8]; // tile row / column of every slot; slot 0 of warp 0 is tile (0, GT-1)
    unsigned char slot_c[16][8];
};

// P2 of a warp's slots: NA consecutive tile rows of one column (segment A) and NB of another (segment B); the tiles of
// a segment share their B fragment.  (One A and one B fragment per DMMA - independent slots - saturates the
// shared-memory pipe: 4 wavefronts per DMMA is exactly the tensor pipe's rate; measured 0.51 instead of 0.80.)
template <int NA, int NB, int LDA, int LDBB>
__device__ __forceinline__ void mmact_p2(double (&acc)[MMACT_MAXS][2], const double* __restrict__ ApA, const double* __restrict__ BpA,
                                         const double* __restrict__ ApB, const double* __restrict__ BpB, int NK) {
#pragma unroll 1
    for (int k0 = 0; k0 < NK; k0 += 4) {
        double a[NA + NB];
#pragma unroll
        for (int i = 0; i < NA; ++i) a[i] = ApA[8 * i * LDA + k0];
#pragma unroll
        for (int i = 0; i < NB; ++i) a[NA + i] = ApB[8 * i * LDA + k0];
        const double bA = BpA[k0 * LDBB];
        const double bB = NB > 0 ? BpB[k0 * LDBB] : 0.0;
#pragma unroll
        for (int i = 0; i < NA; ++i) dmma884(acc[i], a[i], bA);
#pragma unroll
        for (int i = 0; i < NB; ++i) dmma884(acc[NA + i], a[NA + i], bB);
    }
}

// NW warps in total: GT column warps for P1 (+ NH helpers); warps beyond GT + NH only join P2 / update / write-back
template <int GT, int NH, int NW>
__global__ void __maxnreg__((16384 / (32 * ((NW + 3) / 4))) / 8 * 8 > 160 ? 160 : (16384 / (32 * ((NW + 3) / 4))) / 8 * 8)
k_mmact(const __grid_constant__ CTParams cp) {
    static_assert(NW >= GT + NH && NW <= 16, "warp budget");
    static_assert(NH == 0 || (NH == 3 && GT % 4 == 1 && GT >= 9), "helper warps are the GT = 4k + 1 variant");
    constexpr int TJM = GT - 1;                   // the mean rides in the padding columns of the last tile column
    constexpr int NPm = 8 * GT, LDB = NPm + 4, LDC = 8 * GT + 4;
    constexpr int MATB = NPm * LDB, MATG = NPm * NPm;
    constexpr int HR = NH ? GT / 4 : 0;
    const MParams& mp = cp.m;
    const KParams& p = mp.k;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    double* Bsm = reinterpret_cast<double*>(smem_raw + 16);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, c4 = lane & 3;
    const int e_sub = blockIdx.y;
    const int N = p.N, D = p.D, NK = mp.NK;
    // P1 roles: warps 0..GT-1 own column wid; warps 0, 4, 8 are helped by warps GT, GT+1, GT+2
    const bool helper = NH > 0 && wid >= GT && wid < GT + NH;
    const bool p1idle = wid >= GT + NH;                      // no P1 work: barriers only
    const bool helped = NH > 0 && !helper && (wid & 3) == 0 && (wid >> 2) < NH;
    const int pc = helper ? 4 * (wid - GT) : wid;           // P1 column
    const int pair_id = 1 + (helper ? wid - GT : (wid >> 2));
    // P2 / update / write-back slots
    const int nt = cp.nslot[wid], nA = cp.nsegA[wid];
    int sti[MMACT_MAXS], sc[MMACT_MAXS];
    bool meanw = false;                                     // holds a tile of the last column: carries part of the mean
#pragma unroll
    for (int i = 0; i < MMACT_MAXS; ++i) {
        sti[i] = cp.slot_ti[wid][i];
        sc[i] = cp.slot_c[wid][i];
        meanw |= (i < nt) && sc[i] == GT - 1;
    }
    const bool mown = (wid == 0);                           // slot 0 of warp 0 is tile (0, GT-1): it also keeps the logL sums

    const int tj = p.cta_traj ? p.cta_traj[blockIdx.x] : 0;
    const int pidx = p.cta_first ? p.cta_first[blockIdx.x] : blockIdx.x;

    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    auto stage_B = [&](int st, int slot) {
        constexpr uint32_t CH = 32768;
        constexpr uint32_t bytes = MATB * sizeof(double);
        for (uint32_t off = 0; off < bytes; off += CH)
            tma_load_1d(reinterpret_cast<char*>(Bsm + slot * MATB) + off, reinterpret_cast<const char*>(mp.Bm + static_cast<size_t>(MATB) * st) + off,
                        bytes - off < CH ? bytes - off : CH, mbar);
    };
    uint32_t bphase = 0;
    int s_loaded = -1;
    if (cp.b_all) {
        if (tid == 0) {
            mbar_expect_tx(mbar, static_cast<uint32_t>(MATB * p.S * sizeof(double)));
            for (int st = 0; st < p.S; ++st) stage_B(st, st);
        }
    }

    double* Cb = Bsm + MATB * (cp.b_all ? p.S : 1);   // [NPm][LDC]
    double* colb = Cb + NPm * LDC;                    // [2][NPm]
    double* const lst = colb + 2 * NPm;               // [0] mantissa  [1] (int2) exponent sum, valid frames

    const int T = p.T[tj];
    const double* __restrict__ xg = p.x[tj];
    const uint32_t* __restrict__ vbits = reinterpret_cast<const uint32_t*>(p.valid[tj] + (T + 3) / 4 * 4);
    uint32_t vword = 0;
    const int ncols = p.ncols[e_sub];
    const double s2 = p.s2[e_sub];
    const int j0 = p.wz_idx[0], j1 = p.wz_idx[1];
    const double w0 = p.wz_val[0], w1 = p.wz_val[1];

    const int q0 = 2 * c4 - (mp.MC0 - 8 * TJM), q1 = q0 + 1;
    const bool qv0 = static_cast<unsigned>(q0) < static_cast<unsigned>(ncols);
    const bool qv1 = static_cast<unsigned>(q1) < static_cast<unsigned>(ncols);
    const int xc0 = p.cols[e_sub][qv0 ? q0 : 0], xc1 = p.cols[e_sub][qv1 ? q1 : 0];
    double quad = 0.0;
    if (tid == 0) { lst[0] = 1.0; reinterpret_cast<int*>(lst + 1)[0] = 0; reinterpret_cast<int*>(lst + 1)[1] = 0; }

    int r_cur = 0;
    int s = p.run_states[static_cast<size_t>(pidx) * p.K1];
    int next_sw = (p.K1 > 1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + 1] : 0x7fffffff;

    double* const myC = Cb + g * LDC + 2 * c4;   // accumulator pair of tile (ti, tj): myC + 8 ti LDC + 8 tj

    if (cp.b_all) mbar_wait(mbar, 0);

    for (int t = 0; t < T; ++t) {
        while (t >= next_sw) {
            ++r_cur;
            s = p.run_states[static_cast<size_t>(pidx) * p.K1 + r_cur];
            next_sw = (r_cur + 1 < p.K1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + r_cur + 1] : 0x7fffffff;
        }
        if ((t & 31) == 0) vword = __ldg(vbits + (t >> 5));
        const bool is_valid = (vword >> (t & 31)) & 1u;
        if (!cp.b_all && t > 0 && s != s_loaded) {   // all readers of the old propagator passed the last barrier
            if (tid == 0) {
                mbar_expect_tx(mbar, static_cast<uint32_t>(MATB * sizeof(double)));
                stage_B(s, 0);
            }
            mbar_wait(mbar, bphase);
            bphase ^= 1;
            s_loaded = s;
        }
        const double* Bs = Bsm + (cp.b_all ? s * MATB : 0);
        const double* Gsrc = ((t == 0) ? mp.C0m : mp.Sigm) + static_cast<size_t>(MATG) * s + g * NPm + 2 * c4;

        if (t > 0 && !p1idle) {
            // ---------------- P1: T[:, pc] = B_s Caug[:, pc], in place (MSRouse_logL.pyx:206-241, first product)
            double acc[GT][2];
            const double* Ap = Bs + g * LDB + c4;
            const double* Bp = Cb + c4 * LDC + 8 * pc + g;
            if (helper) {
                mmac_p1<GT, 0, (HR > 0 ? HR : 1), LDB, LDC>(acc, Ap, Bp, NK);
                asm volatile("bar.sync %0, 64;" ::"r"(pair_id) : "memory");   // owner and helper have read column pc
#pragma unroll
                for (int ti = 0; ti < HR; ++ti)
                    *reinterpret_cast<double2*>(myC + 8 * ti * LDC + 8 * pc) = make_double2(acc[ti][0], acc[ti][1]);
            } else if (helped) {
                mmac_p1<GT, HR, GT, LDB, LDC>(acc, Ap, Bp, NK);
                asm volatile("bar.sync %0, 64;" ::"r"(pair_id) : "memory");
#pragma unroll
                for (int ti = HR; ti < GT; ++ti)
                    *reinterpret_cast<double2*>(myC + 8 * ti * LDC + 8 * pc) = make_double2(acc[ti][0], acc[ti][1]);
            } else {
                mmac_p1<GT, 0, GT, LDB, LDC>(acc, Ap, Bp, NK);
                __syncwarp();   // this warp is the only reader of column pc
#pragma unroll
                for (int ti = 0; ti < GT; ++ti)
                    *reinterpret_cast<double2*>(myC + 8 * ti * LDC + 8 * pc) = make_double2(acc[ti][0], acc[ti][1]);
            }
        }
        // slots start at Sig (t > 0) / hold C0 (t = 0); the global loads overlap the barrier
        double acc[MMACT_MAXS][2];
#pragma unroll
        for (int i = 0; i < MMACT_MAXS; ++i)
            if (i < nt) {
                const double2 v = __ldg(reinterpret_cast<const double2*>(Gsrc + 8 * sti[i] * NPm + 8 * sc[i]));
                acc[i][0] = v.x;
                acc[i][1] = v.y;
            } else {
                acc[i][0] = acc[i][1] = 0.0;
            }
        double x0 = 0.0, x1 = 0.0;
        if (is_valid && meanw) {
            if (qv0) x0 = __ldg(xg + t * D + xc0);
            if (qv1) x1 = __ldg(xg + t * D + xc1);
        }
        if (t > 0) {
            __syncthreads();   // T (with M' in the padding columns of its last tile column) complete
            // ---------------- P2: C'[ti][c] = T[ti][:] B_s[:][c] + Sig for the warp's slots
            const int sB = nA < nt ? nA : 0;     // first slot of segment B (if any)
            const double* ApA = Cb + (8 * sti[0] + g) * LDC + c4;
            const double* BpA = Bs + c4 * LDB + 8 * sc[0] + g;
            const double* ApB = Cb + (8 * cp.slot_ti[wid][sB] + g) * LDC + c4;
            const double* BpB = Bs + c4 * LDB + 8 * cp.slot_c[wid][sB] + g;
            switch (nA * 8 + (nt - nA)) {   // instantiated per shape: no predicates in the hot loop
#define P2S(A_, B_) case A_ * 8 + B_: mmact_p2<A_, B_, LDC, LDB>(acc, ApA, BpA, ApB, BpB, NK); break;
                P2S(1, 0) P2S(1, 1) P2S(1, 2) P2S(1, 3) P2S(1, 4) P2S(1, 5) P2S(1, 6)
                P2S(2, 0) P2S(2, 1) P2S(2, 2) P2S(2, 3) P2S(2, 4) P2S(2, 5)
                P2S(3, 0) P2S(3, 1) P2S(3, 2) P2S(3, 3) P2S(3, 4)
                P2S(4, 0) P2S(4, 1) P2S(4, 2) P2S(4, 3)
                P2S(5, 0) P2S(5, 1) P2S(5, 2)
                P2S(6, 0) P2S(6, 1)
                P2S(7, 0)
#undef P2S
                default: break;
            }
        }
        // innovation x - w . M' (pyx:79) from the prior mean rows j0, j1: read BEFORE the publish barrier, M+ is
        // written after it, so the warps that share the last tile column never race
        double xm0 = 0.0, xm1 = 0.0;
        if (is_valid && meanw) {
            if (qv0) {
                double ma, mb;
                if (t == 0) { ma = __ldg(p.M0 + (s * N + j0) * D + xc0); mb = __ldg(p.M0 + (s * N + j1) * D + xc0); }
                else {
                    ma = Cb[j0 * LDC + mp.MC0 + q0]; mb = Cb[j1 * LDC + mp.MC0 + q0];
                    if (p.hasG) { ma += __ldg(p.Gm + (s * N + j0) * D + xc0); mb += __ldg(p.Gm + (s * N + j1) * D + xc0); }
                }
                xm0 = x0 - fma(w1, mb, w0 * ma);
            }
            if (qv1) {
                double ma, mb;
                if (t == 0) { ma = __ldg(p.M0 + (s * N + j0) * D + xc1); mb = __ldg(p.M0 + (s * N + j1) * D + xc1); }
                else {
                    ma = Cb[j0 * LDC + mp.MC0 + q1]; mb = Cb[j1 * LDC + mp.MC0 + q1];
                    if (p.hasG) { ma += __ldg(p.Gm + (s * N + j0) * D + xc1); mb += __ldg(p.Gm + (s * N + j1) * D + xc1); }
                }
                xm1 = x1 - fma(w1, mb, w0 * ma);
            }
        }

        auto mean_prior = [&](int ti, double& m0, double& m1) {
            const int row = 8 * ti + g;
            if (t == 0) {
                m0 = (qv0 && row < N) ? __ldg(p.M0 + (s * N + row) * D + xc0) : 0.0;
                m1 = (qv1 && row < N) ? __ldg(p.M0 + (s * N + row) * D + xc1) : 0.0;
            } else {
                const double2 v = *reinterpret_cast<const double2*>(myC + 8 * ti * LDC + 8 * TJM);
                m0 = qv0 ? v.x : 0.0;
                m1 = qv1 ? v.y : 0.0;
                if (p.hasG && row < N) {
                    if (qv0) m0 += __ldg(p.Gm + (s * N + row) * D + xc0);
                    if (qv1) m1 += __ldg(p.Gm + (s * N + row) * D + xc1);
                }
            }
        };
        // prior mean of the slots in the last column: read before the publish barrier as well (tile (ti, GT-1) of T)
        double mp0[MMACT_MAXS], mp1[MMACT_MAXS];
#pragma unroll
        for (int i = 0; i < MMACT_MAXS; ++i) {
            mp0[i] = mp1[i] = 0.0;
            if (i < nt && sc[i] == GT - 1) mean_prior(sti[i], mp0[i], mp1[i]);
        }

        if (is_valid) {
            // publish column j of C' (j = j0, j1; tile column tjz = j >> 3): tile (ti, tjz) holds its rows 8 ti + g;
            // tile (tjz, c) with c > tjz holds, by symmetry, the rows 8 c .. 8 c + 7 in its row j
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int jz = z ? j1 : j0;
                const int tjz = jz >> 3, cj = jz & 7;
#pragma unroll
                for (int i = 0; i < MMACT_MAXS; ++i)
                    if (i < nt) {
                        if (sc[i] == tjz && c4 == (cj >> 1)) colb[z * NPm + 8 * sti[i] + g] = (cj & 1) ? acc[i][1] : acc[i][0];
                        if (sc[i] > tjz && sti[i] == tjz && g == cj)
                            *reinterpret_cast<double2*>(colb + z * NPm + 8 * sc[i] + 2 * c4) = make_double2(acc[i][0], acc[i][1]);
                    }
            }
        }
        __syncthreads();   // T no longer needed; published columns visible
        double kr[MMACT_MAXS];
        if (is_valid) {
            const double cw_j0 = fma(w1, colb[NPm + j0], w0 * colb[j0]);
            const double cw_j1 = fma(w1, colb[NPm + j1], w0 * colb[j1]);
            const double S = fma(w1, cw_j1, fma(w0, cw_j0, s2));
            const double Sinv = __drcp_rn(S);                               // pyx:63
#pragma unroll
            for (int i = 0; i < MMACT_MAXS; ++i)
                if (i < nt) {
                    kr[i] = fma(w1, colb[NPm + 8 * sti[i] + g], w0 * colb[8 * sti[i] + g]) * Sinv;   // pyx:66-67
                    const double2 u = *reinterpret_cast<const double2*>(colb + 8 * sc[i] + 2 * c4);
                    const double2 v = *reinterpret_cast<const double2*>(colb + NPm + 8 * sc[i] + 2 * c4);
                    const double c0v = fma(w1, v.x, w0 * u.x), c1v = fma(w1, v.y, w0 * u.y);
                    acc[i][0] = fma(-kr[i], c0v, acc[i][0]);   // pyx:71-75
                    acc[i][1] = fma(-kr[i], c1v, acc[i][1]);
                }
            if (mown) {
                if (qv0 && g == 0) quad = fma(xm0 * xm0, Sinv, quad);
                if (qv1 && g == 0) quad = fma(xm1 * xm1, Sinv, quad);
                if (lane == 0) {
                    double lmant = lst[0] * Sinv;
                    const int ex = ((__double2hiint(lmant) >> 20) & 0x7ff) - 1023;
                    lmant = __hiloint2double(__double2hiint(lmant) - (ex << 20), __double2loint(lmant));
                    lst[0] = lmant;
                    reinterpret_cast<int*>(lst + 1)[0] += ex;
                    reinterpret_cast<int*>(lst + 1)[1] += 1;
                }
            }
        }
        // ---------------- C+ written back: upper pairs, mirrored below the diagonal; M+ in the padding of the last column
        if (t + 1 < T) {
#pragma unroll
            for (int i = 0; i < MMACT_MAXS; ++i)
                if (i < nt) {
                    const int ti = sti[i], c = sc[i];
                    double v0 = acc[i][0], v1 = acc[i][1];
                    if (ti < c) {   // mirror: C[8 c + 2 c4 + e][8 ti + g]
                        const int r0 = 8 * c + 2 * c4;
                        Cb[r0 * LDC + 8 * ti + g] = v0;
                        Cb[(r0 + 1) * LDC + 8 * ti + g] = v1;
                    }
                    if (c == GT - 1) {
                        double m0 = mp0[i], m1 = mp1[i];
                        if (is_valid) {
                            m0 = fma(kr[i], xm0, m0);   // pyx:82-85
                            m1 = fma(kr[i], xm1, m1);
                        }
                        if (qv0) v0 = m0;
                        if (qv1) v1 = m1;
                    }
                    *reinterpret_cast<double2*>(myC + 8 * ti * LDC + 8 * c) = make_double2(v0, v1);
                }
        }
        __syncthreads();   // C+ complete
    }

    if (mown) {
        quad += __shfl_xor_sync(0xffffffffu, quad, 1);
        quad += __shfl_xor_sync(0xffffffffu, quad, 2);
        if (lane == 0) {
            const int lexp = reinterpret_cast<const int*>(lst + 1)[0], nvalid = reinterpret_cast<const int*>(lst + 1)[1];
            const double logdet = log(lst[0]) + lexp * 0.6931471805599453;
            p.out[static_cast<size_t>(e_sub) * p.P + pidx] = -0.5 * (quad - ncols * logdet + static_cast<double>(nvalid) * ncols * LOG_2PI);
        }
    }
}

}  // namespace bildk
