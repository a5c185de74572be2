// k_mmag2 - the L2-workspace tensor-core kernel (k_mmag, 112 < N <= 256) with NC ADJACENT TILE COLUMNS PER WARP (NC = 2, 4).
//
// Why: k_mmag at N = 200 reaches 0.29 of the FP64 peak and is bound by L2 -> L1 fragment traffic, not by the tensor
// pipe: with one tile column per warp every warp streams ALL of B_s (320 KB at N = 200) through its A fragments in
// P1, 25 warps x 320 KB = 8 MB per filter-frame, plus 4 MB of T rows in P2 - about 4.3 TB/s over the whole chip.
// An A fragment (8 rows x 4 k of B_s, or of T) is independent of the output column, so a warp that owns the two
// adjacent columns (2j, 2j+1) feeds two DMMAs per A fragment: P1 and P2 fragment traffic halve, and the 13 warps
// (instead of 25) have twice the registers, so the row chunks grow to CH = 8 tiles (fewer passes over B).
//   P1   warp(j): T[:, 2j..2j+1] = B_s C[:, 2j..2j+1]        (chunks of CH tile rows, stored to the T buffer)
//   ---- CTA barrier
//   P2   warp(j): prior C'[ti <= c, c] = T B_s[:, c] + Sig for c = 2j, 2j+1 (the tile (2j+1, 2j) below the diagonal
//        is computed and dropped); published columns; stored to the C buffer
//   ---- CTA barrier
//   upd  read-modify-write of the warp's own tiles in the C buffer (rank-1 update), mirrored; mean by the owner of
//        the last tile column
//   ---- CTA barrier
// Same arithmetic per tile as k_mmag (identical results).  Column pairs are mapped to warps on the host so that the
// four schedulers carry equal DMMA counts (P2 work grows with the column index).
#pragma once
#include "bildk_mma.cuh"

namespace bildk {

// acc[col][i] += A-fragment(i) x B-fragment(col) over the contraction, for NCA active columns and CH tile rows of which
// the first `nrow` exist.  One A fragment feeds NCA DMMAs.  Rows beyond `nrow` re-read the last existing row (their
// accumulators are never stored), so the loop carries no predicates; the fragments of k-step k+1 are loaded before
// the DMMAs of k-step k are issued (the operands come from L2 / L1: several hundred cycles, and the 3-6 warps per
// scheduler cannot hide that on their own - measured 0.32 -> 0.5 of peak at N = 200).
template <int CH, int NC, int NCA>
__device__ __forceinline__ void mmag_chunk(double (&acc)[NC][CH][2], const double* __restrict__ Ap, int a_tile_stride, int nrow,
                                           const double* __restrict__ Bp, int b_k_stride, int NK) {
    const double* ap[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) ap[i] = Ap + static_cast<size_t>(i < nrow ? i : nrow - 1) * a_tile_stride;
    const size_t bstep = static_cast<size_t>(4) * b_k_stride;
    double a0[CH], b0[NCA], a1[CH], b1[NCA];
#pragma unroll
    for (int i = 0; i < CH; ++i) a0[i] = ap[i][0];
#pragma unroll
    for (int col = 0; col < NCA; ++col) b0[col] = Bp[8 * col];
    int k0 = 0;
#pragma unroll 1
    for (; k0 + 8 <= NK; k0 += 8) {
#pragma unroll
        for (int i = 0; i < CH; ++i) a1[i] = ap[i][k0 + 4];
#pragma unroll
        for (int col = 0; col < NCA; ++col) b1[col] = Bp[bstep + 8 * col];
#pragma unroll
        for (int i = 0; i < CH; ++i)
#pragma unroll
            for (int col = 0; col < NCA; ++col) dmma884(acc[col][i], a0[i], b0[col]);
        Bp += 2 * bstep;
        if (k0 + 8 < NK) {
#pragma unroll
            for (int i = 0; i < CH; ++i) a0[i] = ap[i][k0 + 8];
#pragma unroll
            for (int col = 0; col < NCA; ++col) b0[col] = Bp[8 * col];
        }
#pragma unroll
        for (int i = 0; i < CH; ++i)
#pragma unroll
            for (int col = 0; col < NCA; ++col) dmma884(acc[col][i], a1[i], b1[col]);
    }
    if (k0 < NK) {   // NK = 4 (mod 8): one k-step left, its fragments are in set 0
#pragma unroll
        for (int i = 0; i < CH; ++i)
#pragma unroll
            for (int col = 0; col < NCA; ++col) dmma884(acc[col][i], a0[i], b0[col]);
    }
}
template <int CH, int NC>
__device__ __forceinline__ void mmag_chunk_n(int nca, double (&acc)[NC][CH][2], const double* __restrict__ Ap, int a_tile_stride, int nrow,
                                             const double* __restrict__ Bp, int b_k_stride, int NK) {
    if (nca == NC) mmag_chunk<CH, NC, NC>(acc, Ap, a_tile_stride, nrow, Bp, b_k_stride, NK);
    else if (NC > 3 && nca == 3) mmag_chunk<CH, NC, (NC > 3 ? 3 : 1)>(acc, Ap, a_tile_stride, nrow, Bp, b_k_stride, NK);
    else if (NC > 2 && nca == 2) mmag_chunk<CH, NC, (NC > 2 ? 2 : 1)>(acc, Ap, a_tile_stride, nrow, Bp, b_k_stride, NK);
    else mmag_chunk<CH, NC, 1>(acc, Ap, a_tile_stride, nrow, Bp, b_k_stride, NK);
}

template <int CH, int NC, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_mmag2(const __grid_constant__ GMParams gp, const int GT, const int MXi) {
    const MParams& mp = gp.m;
    const KParams& p = mp.k;
    const bool MX = MXi != 0;
    const int GTC = GT + (MX ? 1 : 0);
    const int TJM = MX ? GT : GT - 1;
    const int NPm = 8 * GT, LDB = mp.LDB, LDC = mp.LDC;
    const size_t MATB = static_cast<size_t>(NPm) * LDB, MATG = static_cast<size_t>(NPm) * NPm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* colb = reinterpret_cast<double*>(smem_raw);    // [2][NPm]
    double* const lst = colb + 2 * NPm;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, c4 = lane & 3;
    const int e_sub = blockIdx.y;
    const int N = p.N, D = p.D, NK = mp.NK;
    const int cA = NC * gp.colmap[wid];           // this warp's tile columns cA .. cA + NC - 1 (of the GTC columns of [C | M])
    const int ncw = (GTC - cA < NC) ? GTC - cA : NC;                       // columns it has in P1
    const int ncu = (GT - cA < NC) ? (GT - cA > 0 ? GT - cA : 0) : NC;     // ... and in P2 / update (covariance columns only)
    const bool mown = (cA <= GT - 1) && (GT - 1 < cA + NC);               // owner of the last covariance column carries the mean

    double* Cg = gp.work + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 2 * NPm * LDC;
    double* Tg = Cg + static_cast<size_t>(NPm) * LDC;
    const int ncols = p.ncols[e_sub];
    const double s2 = p.s2[e_sub];
    const int j0 = p.wz_idx[0], j1 = p.wz_idx[1];
    const double w0 = p.wz_val[0], w1 = p.wz_val[1];
    const int q0 = 2 * c4 - (mp.MC0 - 8 * TJM), q1 = q0 + 1;
    const bool qv0 = static_cast<unsigned>(q0) < static_cast<unsigned>(ncols);
    const bool qv1 = static_cast<unsigned>(q1) < static_cast<unsigned>(ncols);
    const int xc0 = p.cols[e_sub][qv0 ? q0 : 0], xc1 = p.cols[e_sub][qv1 ? q1 : 0];
    const int pairoff = g * LDC + 2 * c4;   // accumulator pair of tile (ti, tj): + 8 ti LDC + 8 tj

  for (int pidx = blockIdx.x; pidx < p.P; pidx += gridDim.x) {   // a CTA (and its workspace) serves several filters in turn
    const int tj = gp.prof_traj ? gp.prof_traj[pidx] : 0;
    const int T = p.T[tj];
    const double* __restrict__ xg = p.x[tj];
    const uint32_t* __restrict__ vbits = reinterpret_cast<const uint32_t*>(p.valid[tj] + (T + 3) / 4 * 4);
    uint32_t vword = 0;
    double quad = 0.0;
    if (tid == 0) { lst[0] = 1.0; reinterpret_cast<int*>(lst + 1)[0] = 0; reinterpret_cast<int*>(lst + 1)[1] = 0; }

    int r_cur = 0;
    int s = p.run_states[static_cast<size_t>(pidx) * p.K1];
    int next_sw = (p.K1 > 1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + 1] : 0x7fffffff;

    for (int t = 0; t < T; ++t) {
        while (t >= next_sw) {
            ++r_cur;
            s = p.run_states[static_cast<size_t>(pidx) * p.K1 + r_cur];
            next_sw = (r_cur + 1 < p.K1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + r_cur + 1] : 0x7fffffff;
        }
        if ((t & 31) == 0) vword = __ldg(vbits + (t >> 5));
        const bool is_valid = (vword >> (t & 31)) & 1u;
        const double* Bs = mp.Bm + MATB * s;
        const double* Gsrc = ((t == 0) ? mp.C0m : mp.Sigm) + MATG * s + g * NPm + 2 * c4;

        if (t > 0) {
            // ---------------- P1: T[:, cA..] = B_s Caug[:, cA..]; one A fragment feeds all columns of the warp
            for (int r0 = 0; r0 < GT; r0 += CH) {
                double acc[NC][CH][2];
#pragma unroll
                for (int col = 0; col < NC; ++col)
#pragma unroll
                    for (int i = 0; i < CH; ++i) acc[col][i][0] = acc[col][i][1] = 0.0;
                mmag_chunk_n<CH, NC>(ncw, acc, Bs + static_cast<size_t>(8 * r0 + g) * LDB + c4, 8 * LDB, GT - r0,
                                     Cg + c4 * LDC + 8 * cA + g, LDC, NK);
#pragma unroll
                for (int col = 0; col < NC; ++col)
#pragma unroll
                    for (int i = 0; i < CH; ++i)
                        if (col < ncw && r0 + i < GT)
                            *reinterpret_cast<double2*>(Tg + pairoff + 8 * (r0 + i) * LDC + 8 * (cA + col)) = make_double2(acc[col][i][0], acc[col][i][1]);
            }
            __syncthreads();   // T complete
        }
        // ---------------- P2 (t > 0) / steady state (t = 0): prior C' tiles (ti <= c, c) of the warp's columns, published
        //                  columns, stored to C.  Rows up to the largest column index; tiles below the diagonal are dropped.
        if (ncu > 0) {
            const int cL = cA + ncu - 1;
            for (int r0 = 0; r0 <= cL; r0 += CH) {
                double acc[NC][CH][2];
#pragma unroll
                for (int col = 0; col < NC; ++col)
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        acc[col][i][0] = acc[col][i][1] = 0.0;
                        if (col < ncu && r0 + i <= cA + col) {
                            const double2 v = __ldg(reinterpret_cast<const double2*>(Gsrc + 8 * (r0 + i) * NPm + 8 * (cA + col)));
                            acc[col][i][0] = v.x; acc[col][i][1] = v.y;
                        }
                    }
                if (t > 0)
                    mmag_chunk_n<CH, NC>(ncu, acc, Tg + (8 * r0 + g) * LDC + c4, 8 * LDC, cL - r0 + 1,
                                         Bs + c4 * LDB + 8 * cA + g, LDB, NK);
#pragma unroll
                for (int col = 0; col < NC; ++col) {
                    if (col >= ncu) continue;
                    const int c = cA + col;
                    if (is_valid) {
#pragma unroll
                        for (int z = 0; z < 2; ++z) {
                            const int jz = z ? j1 : j0;
                            const int tjz = jz >> 3, cj = jz & 7;
#pragma unroll
                            for (int i = 0; i < CH; ++i) {
                                const int ti = r0 + i;
                                if (ti <= c) {
                                    if (c == tjz && c4 == (cj >> 1)) colb[z * NPm + 8 * ti + g] = (cj & 1) ? acc[col][i][1] : acc[col][i][0];
                                    if (c > tjz && ti == tjz && g == cj)
                                        *reinterpret_cast<double2*>(colb + z * NPm + 8 * c + 2 * c4) = make_double2(acc[col][i][0], acc[col][i][1]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < CH; ++i)
                        if (r0 + i <= c)
                            *reinterpret_cast<double2*>(Cg + pairoff + 8 * (r0 + i) * LDC + 8 * c) = make_double2(acc[col][i][0], acc[col][i][1]);
                }
            }
        }
        // mean prior: M0 at t = 0, else M' from the T buffer (written with T); kept in Cg's mean columns
        auto mean_prior = [&](int ti, double& m0, double& m1) {
            const int row = 8 * ti + g;
            if (t == 0) {
                m0 = (qv0 && row < N) ? __ldg(p.M0 + (s * N + row) * D + xc0) : 0.0;
                m1 = (qv1 && row < N) ? __ldg(p.M0 + (s * N + row) * D + xc1) : 0.0;
            } else {
                const double2 v = *reinterpret_cast<const double2*>(Tg + pairoff + 8 * ti * LDC + 8 * TJM);
                m0 = qv0 ? v.x : 0.0;
                m1 = qv1 ? v.y : 0.0;
                if (p.hasG && row < N) {
                    if (qv0) m0 += __ldg(p.Gm + (s * N + row) * D + xc0);
                    if (qv1) m1 += __ldg(p.Gm + (s * N + row) * D + xc1);
                }
            }
        };
        double* Msrc = (t == 0) ? Cg : Tg;   // where w . M' is read from
        if (is_valid && t == 0 && mown) {
            for (int ti = 0; ti < GT; ++ti) {
                double m0, m1;
                mean_prior(ti, m0, m1);
                if (qv0) Cg[(8 * ti + g) * LDC + mp.MC0 + q0] = m0;
                if (qv1) Cg[(8 * ti + g) * LDC + mp.MC0 + q1] = m1;
            }
        }
        __syncthreads();   // prior C' (and published columns, M') visible
        double xm0 = 0.0, xm1 = 0.0, Sinv = 0.0;
        if (is_valid) {
            const double cw_j0 = fma(w1, colb[NPm + j0], w0 * colb[j0]);
            const double cw_j1 = fma(w1, colb[NPm + j1], w0 * colb[j1]);
            const double S = fma(w1, cw_j1, fma(w0, cw_j0, s2));
            Sinv = __drcp_rn(S);                                            // pyx:63
            if (mown) {
                if (qv0) {
                    double ma = Msrc[j0 * LDC + mp.MC0 + q0], mb = Msrc[j1 * LDC + mp.MC0 + q0];
                    if (p.hasG && t > 0) { ma += __ldg(p.Gm + (s * N + j0) * D + xc0); mb += __ldg(p.Gm + (s * N + j1) * D + xc0); }
                    xm0 = __ldg(xg + t * D + xc0) - fma(w1, mb, w0 * ma);   // pyx:79
                    if (g == 0) quad = fma(xm0 * xm0, Sinv, quad);
                }
                if (qv1) {
                    double ma = Msrc[j0 * LDC + mp.MC0 + q1], mb = Msrc[j1 * LDC + mp.MC0 + q1];
                    if (p.hasG && t > 0) { ma += __ldg(p.Gm + (s * N + j0) * D + xc1); mb += __ldg(p.Gm + (s * N + j1) * D + xc1); }
                    xm1 = __ldg(xg + t * D + xc1) - fma(w1, mb, w0 * ma);
                    if (g == 0) quad = fma(xm1 * xm1, Sinv, quad);
                }
                if (lane == 0) {
                    double lmant = lst[0] * Sinv;
                    const int ex = ((__double2hiint(lmant) >> 20) & 0x7ff) - 1023;
                    lmant = __hiloint2double(__double2hiint(lmant) - (ex << 20), __double2loint(lmant));
                    lst[0] = lmant;
                    reinterpret_cast<int*>(lst + 1)[0] += ex;
                    reinterpret_cast<int*>(lst + 1)[1] += 1;
                }
                __syncwarp();
            }
        }
        // ---------------- update + write-back of the warp's own tiles (read-modify-write in the C buffer), mirrored
        if (t + 1 < T) {
            for (int col = 0; col < ncu; ++col) {
                const int c = cA + col;
                const bool mcol = mown && (c == GT - 1);
                double c0v = 0.0, c1v = 0.0;
                if (is_valid) {
                    const double2 u = *reinterpret_cast<const double2*>(colb + 8 * c + 2 * c4);
                    const double2 v = *reinterpret_cast<const double2*>(colb + NPm + 8 * c + 2 * c4);
                    c0v = fma(w1, v.x, w0 * u.x);
                    c1v = fma(w1, v.y, w0 * u.y);
                }
                for (int ti = 0; ti <= c; ++ti) {
                    double2 v = *reinterpret_cast<const double2*>(Cg + pairoff + 8 * ti * LDC + 8 * c);
                    double kr = 0.0;
                    if (is_valid) {
                        kr = fma(w1, colb[NPm + 8 * ti + g], w0 * colb[8 * ti + g]) * Sinv;   // pyx:66-67
                        v.x = fma(-kr, c0v, v.x);                                               // pyx:71-75
                        v.y = fma(-kr, c1v, v.y);
                    }
                    if (ti < c) {
                        const int r0 = 8 * c + 2 * c4;
                        Cg[r0 * LDC + 8 * ti + g] = v.x;
                        Cg[(r0 + 1) * LDC + 8 * ti + g] = v.y;
                    }
                    if (mcol) {
                        double m0, m1;
                        mean_prior(ti, m0, m1);
                        if (is_valid) {
                            m0 = fma(kr, xm0, m0);   // pyx:82-85
                            m1 = fma(kr, xm1, m1);
                        }
                        if (!MX) {
                            if (qv0) v.x = m0;
                            if (qv1) v.y = m1;
                        } else {
                            *reinterpret_cast<double2*>(Cg + pairoff + 8 * ti * LDC + 8 * TJM) = make_double2(m0, m1);
                        }
                    }
                    *reinterpret_cast<double2*>(Cg + pairoff + 8 * ti * LDC + 8 * c) = v;
                }
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
    __syncthreads();   // the log-likelihood state and the workspace are reused by the next filter
  }
}

}  // namespace bildk
