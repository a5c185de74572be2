// k_mmarb - k_mmar for N mod 8 in {1, 2} (N = 9, 10, 17, 18, 25, 26: the sweep's N = 10 and N = 25) with the r = N mod 8
// BORDER rows / columns of the covariance taken off the tensor cores.
//
// Why: on 8x8 tiles a polymer of N = 8 (GT - 1) + r monomers pays a whole tile row and a whole tile column for its last r
// monomers.  In k_mmar that is (i) P1 for the last tile-row block - GT tile products with r of 8 rows valid, needed only for
// the corner C'[a][a'] and the prior mean of the border rows - and (ii) the P2 tiles (ti, GT-1) - GT tile products with r of
// 8 columns valid: 56 of 182 DMMAs per frame at N = 25, 12 of 21 at N = 10 (profiles/r01_bench_n25_v9_mmar.json: 0.66 of the
// FP64 peak with the tensor pipe 93 % busy - only less padding helps).
//
// Here the tensor cores work on the 8 (GT - 1) core rows only:
//     T[core][:] = B_s[core][:] [C | M]   (all N columns: the permuted last tile column carries the r border columns and the mean)
//     C'[core][core] = T[core][:] B_s[:][core] + Sig            (upper tiles, contraction over all N)
// and column a = N - r + j of C' (all N rows, corner included) is two matrix-vector products in plain DFMAs, one lane per row,
// using the symmetry of B_s and C:
//     t_j = [C | M]^T b_a  (b_a = row a of B_s; t_j[N + q] = M[:, q] . b_a = prior mean of border row a, dimension q)
//     C'[:, a] = B_s t_j + Sig[:, a]
// 2 r N DFMAs per lane-row and frame instead of 2 GT tile products: N = 25: 126 DMMAs + 52 DFMAs (182 DMMAs before).
// Measured (B200, profiles/r02_border_variants.txt): N = 25 0.674 -> 0.826 of the DMMA peak, N = 17 0.503 -> 0.604, N = 26
// (r = 2) 0.751 -> 0.809; it loses at N = 18 (r = 2, GT = 3: 0.595 -> 0.563) and at GT = 2 (N = 10: 0.33 -> 0.20, nine DMMAs
// per frame cannot hide two dependent DFMA chains): a scalar DFMA occupies the FP64 pipe for 2 cycles but has to win it
// against 16-cycle DMMAs of the other warps each time.  The library therefore selects this kernel for N = 17, 25, 26; the
// other instantiations are compiled for the parity tests (BILDK_MMARB=2).  The border lives in "lane = row" registers (cbv[j] = C'[lane][a_j]); lanes N .. N + ncols - 1
// own the mean of the border rows.  The fragment layouts, the permuted last tile column, the swizzle and the update of the
// core are those of k_mmar (bildk_mmar.cuh).
#pragma once
#include "bildk_mmar.cuh"

namespace bildk {

// The two matrix-vector products read their matrix COLUMN-wise (lane = column, one row per step: 32 consecutive doubles,
// conflict-free).  A first version read row-wise (lane = row, k contiguous: 128-bit loads with a 2-way bank conflict between rows
// i and i + 4); ncu (profiles/r02_ncu_n25.txt) showed the shared-memory pipe 68 % busy with a quarter of its wavefronts
// conflicts - column-wise: N = 25 0.781 -> 0.826, N = 17 0.552 -> 0.604, and r = 2 at GT = 4 now beats k_mmar too (N = 26
// 0.751 -> 0.809).  For this the mean rides as d extra COLUMNS N .. N+d-1 of the buffer rows (padding columns of the last
// tile, which every tensor-core operand multiplies by zero rows / columns of B_s) next to the M^T rows that the permuted tile
// column reads.
template <int GT, int NB, int RB>
__global__ void __launch_bounds__(128, NB) k_mmarb(const __grid_constant__ RParams rp) {
    static_assert(GT >= 2 && GT <= 4 && (RB == 1 || RB == 2), "border kernel: N = 8 (GT - 1) + RB, GT 2..4");
    using G = MmarGeom<GT, false>;
    constexpr int R = G::R, LD = G::LD, MAT = G::MAT;
    constexpr int KT = GT - 1;                     // full k-tiles; the last k-tile (border columns) is a single k-step
    constexpr int GR = GT - 1;                     // core tile rows / columns
    constexpr int NU = GR * (GR + 1) / 2;          // upper tiles of the core of C'
    constexpr int BASE = 8 * GR;                   // first border row = N - RB
    constexpr int KP = BASE + 2;                   // contraction length of the matrix-vector products (even, >= N; zeros beyond N)
#define UIDX(ti, tjj) ((ti) * GR - (ti) * ((ti) - 1) / 2 + ((tjj) - (ti)))
    const KParams& p = rp.k;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    double* Bsm = reinterpret_cast<double*>(smem_raw + 16);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, c4 = lane & 3;
    const int e_sub = blockIdx.y;
    const int N = p.N, D = p.D;

    const int tjx = p.cta_traj ? p.cta_traj[blockIdx.x] : 0;
    const int first = p.cta_first ? p.cta_first[blockIdx.x] : blockIdx.x * rp.WPC;
    const int pend = p.traj_first[tjx + 1];
    const int pidx = first + wid;
    const bool alive = (wid < rp.WPC) && (pidx < pend);

    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(mbar, static_cast<uint32_t>(MAT * p.S * sizeof(double)));
        for (int st = 0; st < p.S; ++st) {
            constexpr uint32_t CH = 32768;
            constexpr uint32_t bytes = MAT * sizeof(double);
            for (uint32_t off = 0; off < bytes; off += CH)
                tma_load_1d(reinterpret_cast<char*>(Bsm + st * MAT) + off, reinterpret_cast<const char*>(rp.Br + static_cast<size_t>(MAT) * st) + off,
                            bytes - off < CH ? bytes - off : CH, mbar);
        }
    }
    if (!alive) return;   // warps are independent from here on (warp-scope barriers only)

    double* const Cb = Bsm + MAT * p.S + wid * rp.fstride;   // [R][LD]: rows < N covariance, spare rows of the last block M^T / zero
    double* const colb = Cb + MAT;                            // [2][R] the two columns of C' that w touches
    double* const mpub = colb + 2 * R;                        // [2][4] prior mean rows 0, N - 1
    double* const tb = mpub + 8;                              // [2][R] t_j = C b_a, logical order (entries >= N stay zero)
    for (int i = lane; i < rp.fstride; i += 32) Cb[i] = 0.0;

    const int T = p.T[tjx];
    const double* __restrict__ xg = p.x[tjx];
    const uint32_t* __restrict__ vbits = reinterpret_cast<const uint32_t*>(p.valid[tjx] + (T + 3) / 4 * 4);
    uint32_t vword = 0;
    const int ncols = p.ncols[e_sub];
    const double s2 = p.s2[e_sub];
    const double w0 = p.wz_val[0], w1 = p.wz_val[1];

    // ---- core (tensor cores): lane-constant fragment offsets as in k_mmar
    const int fx = 4 * ((g >> 1) & 1);
    const int offP = g * LD + ((2 * c4) ^ fx);
    const int offS = g * LD + 8 * KT + c4 + fx;
    const int lr = rp.lastrow[e_sub][g];
    const int lx = 4 * ((lr >> 1) & 1);
    const int offLP = lr * LD + ((2 * c4) ^ lx);
    const int offLS = lr * LD + 8 * KT + c4 + lx;
    const int offMir = 2 * c4 * LD + (g ^ (4 * (c4 & 1)));
    const bool hasq = c4 < ncols;
    const int mr = rp.mrow[e_sub][hasq ? c4 : 0];
    const int offM = mr * LD + (g ^ (4 * ((mr >> 1) & 1)));
    const int xcol = p.cols[e_sub][hasq ? c4 : 0];

    // ---- border (DFMA): lane i < N owns row i of the border columns; lane N + q the mean of the border rows, dimension q
    const bool brow = lane < N;
    const int bq = lane - N;
    const bool bmean = bq >= 0 && bq < ncols;
    const int rowi = brow ? lane : 0;
    const int fi = 4 * ((rowi >> 1) & 1);
    const int mrb = rp.mrow[e_sub][bmean ? bq : 0];
    const int xcolb = p.cols[e_sub][bmean ? bq : 0];
    const int lane4 = lane ^ 4;                                    // column `lane` in rows whose column bit 2 is flipped
    const int offMc = g * LD + ((N + c4) ^ fx);                    // M[8 ti + g][q = c4] as column N + q of row 8 ti + g: + 8 ti LD

    double quad = 0.0, lmant = 1.0;
    int lexp = 0;

    int r_cur = 0;
    int s = p.run_states[static_cast<size_t>(pidx) * p.K1];
    int next_sw = (p.K1 > 1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + 1] : 0x7fffffff;

    double acc[NU][2];
    double mu[GR];            // prior / posterior mean M[8 ti + g][q = c4] of the core rows
    double cbv[RB];           // C'[lane][BASE + j]   (lanes < N)
    double mb[RB];            // M[BASE + j][q]       (lanes N + q)

    __syncwarp();
    mbar_wait(mbar, 0);

    for (int t = 0; t < T; ++t) {
        while (t >= next_sw) {
            ++r_cur;
            s = p.run_states[static_cast<size_t>(pidx) * p.K1 + r_cur];
            next_sw = (r_cur + 1 < p.K1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + r_cur + 1] : 0x7fffffff;
        }
        if ((t & 31) == 0) vword = __ldg(vbits + (t >> 5));
        const bool is_valid = (vword >> (t & 31)) & 1u;

        if (t > 0) {
            const double* __restrict__ Bs = Bsm + s * MAT;
            const double* __restrict__ Gs = rp.Sigm + static_cast<size_t>(R * R) * s + g * R + 2 * c4;
            // ---------------- border, first product: t_j = [C | M]^T b_a, a = BASE + j  (the border rows of T = B_s [C | M], pyx:206-241)
            // Both border products stand AHEAD of the tile rows in program order (ptxas lets the second one sink behind the last
            // DMMAs on its own).  Measured alternatives (profiles/r02_border_variants.txt): pinning the second product ahead of
            // the tile rows with a warp barrier, or placing the products behind tile rows 0 and 1 so that their DFMA chains
            // interleave with DMMAs, both cost 6-7 % at N = 25 - a scalar DFMA waits behind the other warps' DMMAs in the one
            // FP64 pipe wherever it stands, and interleaving delays this warp's own DMMAs as well.
            {
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const double* __restrict__ brp = Bs + (BASE + j) * LD;
                    const int fa = 4 * (((BASE + j) >> 1) & 1);
                    double ax = 0.0, ay = 0.0, az = 0.0, aw = 0.0;   // four independent chains
#pragma unroll
                    for (int k = 0; k < KP; k += 2) {
                        const double2 b = *reinterpret_cast<const double2*>(brp + (k ^ fa));        // B_s[a][k], B_s[a][k+1] (broadcast)
                        const int cc = ((k >> 1) & 1) ? lane4 : lane;                                // column `lane` of rows k, k + 1
                        const double c0 = Cb[k * LD + cc];
                        const double c1 = (k + 1 < BASE + RB) ? Cb[(k + 1) * LD + cc] : 0.0;
                        if (k & 2) { az = fma(c0, b.x, az); aw = fma(c1, b.y, aw); }
                        else { ax = fma(c0, b.x, ax); ay = fma(c1, b.y, ay); }
                    }
                    const double tj = (ax + az) + (ay + aw);
                    if (brow) tb[j * R + lane] = tj;
                    mb[j] = tj;                       // lanes N + q: column N + q of the buffer rows is M[:, q]
                }
            }
            __syncwarp();   // t_j complete
            // ---------------- border, second product: C'[:, a] = B_s t_j + Sig[:, a]
            {
                const double* __restrict__ sgp = rp.Sigm + static_cast<size_t>(R * R) * s + rowi * R + BASE;
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    double ax = __ldg(sgp + j), ay = 0.0, az = 0.0, aw = 0.0;
#pragma unroll
                    for (int k = 0; k < KP; k += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(tb + j * R + k);
                        const int cc = ((k >> 1) & 1) ? lane4 : lane;                                // B_s[k][i] = B_s[i][k]: column `lane` of rows k, k + 1
                        const double2 b = make_double2(Bs[k * LD + cc], (k + 1 < BASE + RB) ? Bs[(k + 1) * LD + cc] : 0.0);
                        if (k & 2) { az = fma(b.x, v.x, az); aw = fma(b.y, v.y, aw); }
                        else { ax = fma(b.x, v.x, ax); ay = fma(b.y, v.y, ay); }
                    }
                    cbv[j] = (ax + az) + (ay + aw);
                }
            }
#pragma unroll
            for (int ti = 0; ti < GR; ++ti) {
                // ---------------- P1, core tile-row block ti: T[ti][:] = B_s[ti][:] [C | M]
                double Tt[GT][2];
#pragma unroll
                for (int tj = 0; tj < GT; ++tj) Tt[tj][0] = Tt[tj][1] = 0.0;
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    const double2 a = *reinterpret_cast<const double2*>(Bs + offP + 8 * ti * LD + 8 * kt);
                    double2 b[GT];
#pragma unroll
                    for (int tj = 0; tj < GT; ++tj)
                        b[tj] = *reinterpret_cast<const double2*>(tj < GT - 1 ? Cb + offP + 8 * tj * LD + 8 * kt : Cb + offLP + 8 * kt);
#pragma unroll
                    for (int tj = 0; tj < GT; ++tj) dmma884(Tt[tj], a.x, b[tj].x);
#pragma unroll
                    for (int tj = 0; tj < GT; ++tj) dmma884(Tt[tj], a.y, b[tj].y);
                }
                {
                    const double a = Bs[offS + 8 * ti * LD];
                    double b[GT];
#pragma unroll
                    for (int tj = 0; tj < GT; ++tj) b[tj] = tj < GT - 1 ? Cb[offS + 8 * tj * LD] : Cb[offLS];
#pragma unroll
                    for (int tj = 0; tj < GT; ++tj) dmma884(Tt[tj], a, b[tj]);
                }
                mu[ti] = Tt[GT - 1][1];   // M'[8 ti + g][c4]
                // ---------------- P2, upper core tiles of tile row ti: C'[ti][tj] = Sig + T[ti][:] B_s[:][tj]
#pragma unroll
                for (int tj = ti; tj < GR; ++tj) {
                    const double2 v = __ldg(reinterpret_cast<const double2*>(Gs + 8 * ti * R + 8 * tj));
                    acc[UIDX(ti, tj)][0] = v.x;
                    acc[UIDX(ti, tj)][1] = v.y;
                }
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    double2 b[GR];
#pragma unroll
                    for (int tj = ti; tj < GR; ++tj) b[tj] = *reinterpret_cast<const double2*>(Bs + offP + 8 * tj * LD + 8 * kt);
#pragma unroll
                    for (int tj = ti; tj < GR; ++tj) dmma884(acc[UIDX(ti, tj)], Tt[kt][0], b[tj].x);
#pragma unroll
                    for (int tj = ti; tj < GR; ++tj) dmma884(acc[UIDX(ti, tj)], Tt[kt][1], b[tj].y);
                }
                {
                    double b[GR];
#pragma unroll
                    for (int tj = ti; tj < GR; ++tj) b[tj] = Bs[offS + 8 * tj * LD];
#pragma unroll
                    for (int tj = ti; tj < GR; ++tj) dmma884(acc[UIDX(ti, tj)], Tt[GT - 1][0], b[tj]);
                }
            }
            if (p.hasG) {   // M' = B M + G  (pyx:209-214)
#pragma unroll
                for (int ti = 0; ti < GR; ++ti)
                    if (hasq) mu[ti] += __ldg(p.Gm + (s * N + 8 * ti + g) * D + xcol);
                if (bmean) {
#pragma unroll
                    for (int j = 0; j < RB; ++j) mb[j] += __ldg(p.Gm + (s * N + BASE + j) * D + xcolb);
                }
            }
        } else {
            // frame 0: steady state of the first state (pyx:160-163), no propagation
            const double* __restrict__ Gs = rp.C0m + static_cast<size_t>(R * R) * s + g * R + 2 * c4;
#pragma unroll
            for (int ti = 0; ti < GR; ++ti) {
#pragma unroll
                for (int tj = ti; tj < GR; ++tj) {
                    const double2 v = __ldg(reinterpret_cast<const double2*>(Gs + 8 * ti * R + 8 * tj));
                    acc[UIDX(ti, tj)][0] = v.x;
                    acc[UIDX(ti, tj)][1] = v.y;
                }
                mu[ti] = hasq ? __ldg(p.M0 + (s * N + 8 * ti + g) * D + xcol) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                cbv[j] = __ldg(rp.C0m + static_cast<size_t>(R * R) * s + rowi * R + BASE + j);
                mb[j] = bmean ? __ldg(p.M0 + (s * N + BASE + j) * D + xcolb) : 0.0;
            }
        }

        double x = 0.0, xb = 0.0;
        if (is_valid) {
            if (hasq) x = __ldg(xg + t * D + xcol);
            if (bmean) xb = __ldg(xg + t * D + xcolb);
            // publish the two columns of C' that w touches: column 0 (core rows from tile (0,0) and, by symmetry, from row 0 of
            // the tiles (0, ti); border rows = C'[0][a_j], lane 0) and column N - 1 = the last border column (lane i: row i)
            if (c4 == 0) colb[g] = acc[UIDX(0, 0)][0];
            if (g == 0) {
#pragma unroll
                for (int ti = 1; ti < GR; ++ti)
                    *reinterpret_cast<double2*>(colb + 8 * ti + 2 * c4) = make_double2(acc[UIDX(0, ti)][0], acc[UIDX(0, ti)][1]);
                mpub[c4] = mu[0];
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < RB; ++j) colb[BASE + j] = cbv[j];
            }
            if (brow) colb[R + lane] = cbv[RB - 1];
            if (bmean) mpub[4 + bq] = mb[RB - 1];          // prior mean row N - 1
        }
        __syncwarp();   // every lane is done reading C / M^T / t; published columns and mean rows visible

        if (is_valid) {
            if (lane < R) colb[lane] = fma(w1, colb[R + lane], w0 * colb[lane]);   // C' w
            __syncwarp();
            // S = s2 + w^T C' w = s2 + w0 (C' w)[0] + w1 (C' w)[N-1]  (pyx:55-63), then 1/S
            const double Sinv = rcp3(fma(w1, colb[BASE + RB - 1], fma(w0, colb[0], s2)));
            double kr[GR];
#pragma unroll
            for (int ti = 0; ti < GR; ++ti) kr[ti] = colb[8 * ti + g] * Sinv;   // K = C' w / S (pyx:66-67)
#pragma unroll
            for (int tjj = 0; tjj < GR; ++tjj) {
                const double2 cw = *reinterpret_cast<const double2*>(colb + 8 * tjj + 2 * c4);
#pragma unroll
                for (int ti = 0; ti <= tjj; ++ti) {
                    acc[UIDX(ti, tjj)][0] = fma(-kr[ti], cw.x, acc[UIDX(ti, tjj)][0]);   // pyx:71-75
                    acc[UIDX(ti, tjj)][1] = fma(-kr[ti], cw.y, acc[UIDX(ti, tjj)][1]);
                }
            }
            // border columns: C+[i][a_j] = C'[i][a_j] - K[i] (C' w)[a_j]
            {
                const double ki = colb[rowi] * Sinv;
#pragma unroll
                for (int j = 0; j < RB; ++j) cbv[j] = fma(-ki, colb[BASE + j], cbv[j]);
            }
            // innovation (pyx:79) and mean update (pyx:82-85) of this lane's dimension
            const double xm = x - fma(w1, mpub[4 + c4], w0 * mpub[c4]);
            quad = fma(xm * xm, Sinv, quad);
#pragma unroll
            for (int ti = 0; ti < GR; ++ti) mu[ti] = fma(kr[ti], xm, mu[ti]);
            if (bmean) {
                const double xmb = xb - fma(w1, mpub[4 + bq], w0 * mpub[bq]);
#pragma unroll
                for (int j = 0; j < RB; ++j) mb[j] = fma(colb[BASE + j] * Sinv, xmb, mb[j]);
            }
            lmant *= Sinv;
            const int ex = ((__double2hiint(lmant) >> 20) & 0x7ff) - 1023;
            lmant = __hiloint2double(__double2hiint(lmant) - (ex << 20), __double2loint(lmant));
            lexp += ex;
        }
        // ---------------- C+ and M+^T become the operands of the next propagation
        if (t + 1 < T) {
#pragma unroll
            for (int ti = 0; ti < GR; ++ti) {
#pragma unroll
                for (int tj = ti; tj < GR; ++tj) {
                    const double v0 = acc[UIDX(ti, tj)][0], v1 = acc[UIDX(ti, tj)][1];
                    if (tj > ti) {   // C[8 tj + 2 c4 + e][8 ti + g] = C[8 ti + g][8 tj + 2 c4 + e]
                        Cb[offMir + (8 * tj) * LD + 8 * ti] = v0;
                        Cb[offMir + (8 * tj + 1) * LD + 8 * ti] = v1;
                    }
                    *reinterpret_cast<double2*>(Cb + offP + 8 * ti * LD + 8 * tj) = make_double2(v0, v1);
                }
                if (hasq) Cb[offM + 8 * ti] = mu[ti];
                if (hasq) Cb[offMc + 8 * ti * LD] = mu[ti];
            }
            // border: element (i, a_j) of every row i < N (corner rows included), element (a_j, i) of the border rows for the
            // core columns i (the corner entries are written once, by their row's lane)
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int a = BASE + j;
                const int fa = 4 * ((a >> 1) & 1);
                if (brow) {
                    Cb[rowi * LD + (a ^ fi)] = cbv[j];
                    if (lane < BASE) Cb[a * LD + (lane ^ fa)] = cbv[j];
                }
                if (bmean) Cb[mrb * LD + (a ^ (4 * ((mrb >> 1) & 1)))] = mb[j];
                if (bmean) Cb[a * LD + ((N + bq) ^ fa)] = mb[j];
            }
        }
        __syncwarp();   // C+ / M+^T complete before the next frame's products
    }

    // logL = -1/2 [ sum xmm^2 Sinv - ncols * sum_t log Sinv_t + nvalid * ncols * log 2 pi ]   (pyx:88, 251-256)
    quad += __shfl_xor_sync(0xffffffffu, quad, 1);
    quad += __shfl_xor_sync(0xffffffffu, quad, 2);
    if (lane == 0) {
        int nvalid = 0;
        for (int wv = 0; wv < (T + 31) / 32; ++wv) nvalid += __popc(__ldg(vbits + wv));
        const double logdet = log(lmant) + lexp * 0.6931471805599453;
        p.out[static_cast<size_t>(e_sub) * p.P + pidx] = -0.5 * (quad - ncols * logdet + static_cast<double>(nvalid) * ncols * LOG_2PI);
    }
#undef UIDX
}

}  // namespace bildk
