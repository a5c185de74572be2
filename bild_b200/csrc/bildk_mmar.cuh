// k_mmar - FP64 tensor-core filter kernel with the intermediate T = B_s [C | M] CHAINED THROUGH REGISTERS
// (one warp per filter, N <= 32 with N mod 8 in 1..4: BASELINE configs[1] N=20, the sweep's N=10 and N=25).
//
// Why a second one-warp kernel: the ncu source view of k_mma<3> on configs[1] (profiles/r01_ncu_c2_mma_v7_final.txt)
// shows 862 instructions per filter-frame of which only 75 are DMMAs; half of the warp-time is spent outside the
// two product loops (T written to shared memory and read back, five warp barriers, address arithmetic that the
// 72-register cap forces to be rematerialised every frame), and the FP64 tensor pipe idles a third of the time.
// This kernel removes the round trip:
//
//   * every operand is read as "row n, contraction index k contiguous".  B_s and C are symmetric, so the
//     B-fragment X[k][n] of a DMMA is row n of the matrix; the mean rides as M^T in spare rows of the last
//     tile-row block.  A lane may pick ANY assignment of contraction indices to (k-step, lane) as long as the
//     A and the B fragment agree, so a lane loads TWO consecutive k (one 128-bit load) and feeds two DMMAs.
//   * the accumulator (D) fragment of a tile - lane (g, c4) holds row g, columns 2 c4, 2 c4 + 1 - is exactly
//     such a two-k A fragment.  T is therefore produced one tile-ROW block at a time (GT tiles, in registers)
//     and consumed on the spot by P2 (C'[ti][tj >= ti] += T[ti][kt] B_s[kt][tj]); it never exists in memory.
//   * the last tile column of T is computed with its columns PERMUTED (the B-fragment lane g reads buffer row
//     lastrow[g]): even slots take the r = N - 8 (GT - 1) <= 4 covariance columns, odd slots the mean columns.
//     Element 0 of the D fragment is then a complete single k-step for P2 (no padded second step), and element
//     1 is this lane's prior mean M'[8 ti + g][q = c4].
//   * frame-constant addressing: all strides are compile-time, the row stride is == 8 (mod 16) doubles and bit 2
//     of the column index is flipped in rows with (row >> 1) & 1, which makes every 128-bit and 64-bit
//     fragment access conflict-free with ONE lane-constant offset per operand class.
// Per frame: 2 warp barriers, no T traffic, ~55 fragment loads for 75 DMMAs (GT = 3).
#pragma once
#include "bildk_mma.cuh"

namespace bildk {

struct RParams {
    KParams k;
    const double* Br;      // [S][8 GT][LD] zero padded, column bit 2 flipped in rows with (row >> 1) & 1
    const double* Sigm;    // [S][8 GT][8 GT] zero padded, plain row-major
    const double* C0m;     // [S][8 GT][8 GT]
    int WPC;               // warps (= filters) per CTA
    int fstride;           // doubles of shared memory per filter
    int r;                 // N - 8 (GT - 1), 1..4
    double ww00, ww11, ww01;   // w0 w0, w1 w1, 2 w0 w1 (host-computed with the same IEEE operations): constant-bank operands of
                               // the S = s2 + w^T C' w FMAs instead of six live registers (k_mmar<3,4> spilled one double per frame)
    unsigned char lastrow[DMAX][8];   // per sub-filter: buffer row read by B-fragment lane g for the last tile column
    unsigned char mrow[DMAX][4];      // per sub-filter: buffer row of mean column q (M^T)
};

// MX ("mean in an extra row block", N mod 8 in {0, 5, 6, 7}): the last tile-row block has no room for M^T next to its
// r = N - 8 (GT - 1) >= 5 covariance rows, so the filter buffer gets eight more rows R .. R+7 (M^T in the odd ones, zeros
// in the even ones), T one more tile column - again read with its columns permuted, even slots = the zero row, odd slots
// = the mean columns, so that element 1 of its D fragment is the lane's prior mean exactly as without MX - and every
// k-tile of both products is a full one (no single k-step).
template <int GT, bool MX = false>
struct MmarGeom {
    static constexpr int R = 8 * GT;
    static constexpr int LD = (R % 16 == 8) ? R : R + 8;   // == 8 (mod 16)
    static constexpr int MAT = R * LD;                     // one propagator
    static constexpr int MATC = (R + (MX ? 8 : 0)) * LD;   // the filter buffer [C ; M^T]
    static constexpr int KT = MX ? GT : GT - 1;            // full k-tiles (two DMMAs per 128-bit fragment pair)
    static constexpr int GTC = GT + (MX ? 1 : 0);          // tile columns of T = B_s [C | M]
    static constexpr int FSTRIDE = MATC + 2 * R + 8;       // buffer | two published columns | published means
};

// 1 / S to within an ulp or two in three dependent FMAs: y0 = rcp.approx (relative error <= 2^-23), e = 1 - S y0,
// y = y0 (1 + e + e^2).  (__drcp_rn's correctly rounded result costs five; the filter's 1e-9 gate does not need it.)
__device__ __forceinline__ double rcp3(double S) {
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(S));
    const double e = fma(-S, y0, 1.0);
    return fma(y0, fma(e, e, e), y0);
}

// CTAs are 4 warps (one per warp scheduler; measured best: fine-grained CTA scheduling keeps the four schedulers of
// an SM evenly loaded).  NB = resident CTAs per SM the kernel is compiled for, i.e. warps per scheduler; registers per
// thread = 65536 / (128 NB).  HIDE: re-read the fragments for every tile row instead of keeping them in registers.
template <int GT, int NB, bool MX = false>
__global__ void __launch_bounds__(128, NB) k_mmar(const __grid_constant__ RParams rp) {
    constexpr bool HIDE = (65536 / (128 * NB)) / 8 * 8 < 40 + 12 * GT * GT / 3 + 28;
    using G = MmarGeom<GT, MX>;
    constexpr int R = G::R, LD = G::LD, MAT = G::MAT, MATC = G::MATC;
    constexpr int KT = G::KT;                      // full k-tiles (two DMMAs per 128-bit fragment pair)
    constexpr int GTC = G::GTC;                    // tile columns of T; the last one is read with permuted columns
    constexpr int NU = GT * (GT + 1) / 2;          // upper tiles of C'
#define UIDX(ti, tjj) ((ti) * GT - (ti) * ((ti) - 1) / 2 + ((tjj) - (ti)))
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
    double* const colb = Cb + MATC;                           // [2][R] the two columns of C' that w touches
    double* const mpub = colb + 2 * R;                        // [2][4] prior mean rows j0, j1
    for (int i = lane; i < rp.fstride; i += 32) Cb[i] = 0.0;  // padding columns and zero rows must stay finite / zero

    const int T = p.T[tjx];
    const double* __restrict__ xg = p.x[tjx];
    const uint32_t* __restrict__ vbits = reinterpret_cast<const uint32_t*>(p.valid[tjx] + (T + 3) / 4 * 4);
    uint32_t vword = 0;
    const int ncols = p.ncols[e_sub];
    const double s2 = p.s2[e_sub];
    const double w0 = p.wz_val[0], w1 = p.wz_val[1];
    const int rr = rp.r;
    // The measurement is BILD's end-to-end distance (models.py:230-233): w = w0 e_0 + w1 e_{N-1} (other 2-sparse
    // vectors use k_mma).  Column 0 of C' lives in tile column 0, column N - 1 = 8 (GT - 1) + cj1 in the last one.
    // (Broadcasting the 2x2 block of C' at (0, N-1) by shuffle, so that 1/S overlaps the publish barrier, was
    // measured 2-3 % SLOWER than reading it back from the published columns: SHFL shares the LSU pipe.)
    const int cj1 = rr - 1;
    const bool e1 = cj1 & 1;

    // lane-constant fragment offsets (doubles).  Rows 8 t + g flip column bit 2 when (g >> 1) & 1.
    const int fx = 4 * ((g >> 1) & 1);
    const int offP = g * LD + ((2 * c4) ^ fx);                 // two-k fragment / accumulator pair: + 8 t LD + 8 kt
    const int offS = g * LD + 8 * KT + c4 + fx;                // single-k fragment of the last k-tile: + 8 t LD
    const int lr = rp.lastrow[e_sub][g];
    const int lx = 4 * ((lr >> 1) & 1);
    const int offLP = lr * LD + ((2 * c4) ^ lx);               // B fragments of the permuted last tile column
    const int offLS = lr * LD + 8 * KT + c4 + lx;
    const int offMir = 2 * c4 * LD + (g ^ (4 * (c4 & 1)));     // mirrored element e of tile (ti, tj): + (8 tj + e) LD + 8 ti
    const bool hasq = c4 < ncols;                               // this lane owns mean column q = c4
    const int mr = rp.mrow[e_sub][hasq ? c4 : 0];
    const int offM = mr * LD + (g ^ (4 * ((mr >> 1) & 1)));    // M^T[q][8 ti + g]: + 8 ti
    const int xcol = p.cols[e_sub][hasq ? c4 : 0];
    const bool lastrow_ok = g < rr;                             // row 8 (GT-1) + g is a covariance row
    const bool mir0_ok = 2 * c4 < rr, mir1_ok = 2 * c4 + 1 < rr;

    double quad = 0.0;        // sum xmm^2 Sinv of this lane's dimension (identical on the 8 lanes that share c4)
    double lmant = 1.0;       // running product of Sinv, exponent split off
    int lexp = 0;

    int r_cur = 0;
    int s = p.run_states[static_cast<size_t>(pidx) * p.K1];
    int next_sw = (p.K1 > 1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + 1] : 0x7fffffff;

    double acc[NU][2];
    double mu[GT];            // prior / posterior mean M[8 ti + g][q = c4]

    __syncwarp();
    mbar_wait(mbar, 0);

    for (int t = 0; t < T; ++t) {
        while (t >= next_sw) {   // rare: at most K1 - 1 times per filter
            ++r_cur;
            s = p.run_states[static_cast<size_t>(pidx) * p.K1 + r_cur];
            next_sw = (r_cur + 1 < p.K1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + r_cur + 1] : 0x7fffffff;
        }
        if ((t & 31) == 0) vword = __ldg(vbits + (t >> 5));
        const bool is_valid = (vword >> (t & 31)) & 1u;

        if (t > 0) {
            const double* __restrict__ Bs = Bsm + s * MAT;
            const double* __restrict__ Gs = rp.Sigm + static_cast<size_t>(R * R) * s + g * R + 2 * c4;
#pragma unroll
            for (int ti = 0; ti < GT; ++ti) {
                // ---------------- P1, tile-row block ti: T[ti][:] = B_s[ti][:] [C | M]   (MSRouse_logL.pyx:206-241)
                // The fragment loads of C and B_s repeat for every ti; keeping them in registers across tile rows
                // (what the compiler does when it can prove the addresses equal) costs 30+ registers and spills
                // the accumulators.  Re-reading shared memory is cheaper than a spill / fill pair: hide the bases.
                int oq = 0;
                if (HIDE) asm volatile("" : "+r"(oq));
                const double* Cq = Cb + oq;
                const double* Bq = Bs + oq;
                double Tt[GTC][2];
#pragma unroll
                for (int tj = 0; tj < GTC; ++tj) Tt[tj][0] = Tt[tj][1] = 0.0;
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    const double2 a = *reinterpret_cast<const double2*>(Bq + offP + 8 * ti * LD + 8 * kt);
                    double2 b[GTC];
#pragma unroll
                    for (int tj = 0; tj < GTC; ++tj)
                        b[tj] = *reinterpret_cast<const double2*>(tj < GTC - 1 ? Cq + offP + 8 * tj * LD + 8 * kt : Cq + offLP + 8 * kt);
#pragma unroll
                    for (int tj = 0; tj < GTC; ++tj) dmma884(Tt[tj], a.x, b[tj].x);
#pragma unroll
                    for (int tj = 0; tj < GTC; ++tj) dmma884(Tt[tj], a.y, b[tj].y);
                }
                if constexpr (!MX) {
                    const double a = Bq[offS + 8 * ti * LD];
                    double b[GT];
#pragma unroll
                    for (int tj = 0; tj < GT; ++tj) b[tj] = tj < GT - 1 ? Cq[offS + 8 * tj * LD] : Cq[offLS];
#pragma unroll
                    for (int tj = 0; tj < GT; ++tj) dmma884(Tt[tj], a, b[tj]);
                }
                mu[ti] = Tt[GTC - 1][1];   // M'[8 ti + g][c4]  (zero for lanes without a mean column)
                // ---------------- P2, upper tiles of tile row ti: C'[ti][tj] = Sig + T[ti][:] B_s[:][tj]
#pragma unroll
                for (int tj = ti; tj < GT; ++tj) {
                    const double2 v = __ldg(reinterpret_cast<const double2*>(Gs + 8 * ti * R + 8 * tj));
                    acc[UIDX(ti, tj)][0] = v.x;
                    acc[UIDX(ti, tj)][1] = v.y;
                }
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    double2 b[GT];
#pragma unroll
                    for (int tj = ti; tj < GT; ++tj) b[tj] = *reinterpret_cast<const double2*>(Bq + offP + 8 * tj * LD + 8 * kt);
#pragma unroll
                    for (int tj = ti; tj < GT; ++tj) dmma884(acc[UIDX(ti, tj)], Tt[kt][0], b[tj].x);
#pragma unroll
                    for (int tj = ti; tj < GT; ++tj) dmma884(acc[UIDX(ti, tj)], Tt[kt][1], b[tj].y);
                }
                if constexpr (!MX) {
                    double b[GT];
#pragma unroll
                    for (int tj = ti; tj < GT; ++tj) b[tj] = Bq[offS + 8 * tj * LD];
#pragma unroll
                    for (int tj = ti; tj < GT; ++tj) dmma884(acc[UIDX(ti, tj)], Tt[GT - 1][0], b[tj]);
                }
            }
            if (p.hasG) {   // M' = B M + G  (pyx:209-214); G = 0 for BILD's force-free chains
#pragma unroll
                for (int ti = 0; ti < GT; ++ti)
                    if (hasq && 8 * ti + g < N) mu[ti] += __ldg(p.Gm + (s * N + 8 * ti + g) * D + xcol);
            }
        } else {
            // frame 0: steady state of the first state (pyx:160-163), no propagation
            const double* __restrict__ Gs = rp.C0m + static_cast<size_t>(R * R) * s + g * R + 2 * c4;
#pragma unroll
            for (int ti = 0; ti < GT; ++ti) {
#pragma unroll
                for (int tj = ti; tj < GT; ++tj) {
                    const double2 v = __ldg(reinterpret_cast<const double2*>(Gs + 8 * ti * R + 8 * tj));
                    acc[UIDX(ti, tj)][0] = v.x;
                    acc[UIDX(ti, tj)][1] = v.y;
                }
                mu[ti] = (hasq && 8 * ti + g < N) ? __ldg(p.M0 + (s * N + 8 * ti + g) * D + xcol) : 0.0;
            }
        }

        double x = 0.0;
        if (is_valid) {
            if (hasq) x = __ldg(xg + t * D + xcol);
            // publish the two columns of C' that w touches.  Column 0: rows 0..7 from tile (0,0) (lanes c4 == 0,
            // element 0), rows of tile rows ti > 0 by symmetry from row 0 of the upper tiles (0, ti) (lanes g == 0).
            // Column N-1: all rows from the upper tiles (ti, GT-1) (lanes c4 == cj1 >> 1, element cj1 & 1).
            if (c4 == 0) colb[g] = acc[UIDX(0, 0)][0];
            if (g == 0) {
#pragma unroll
                for (int ti = 1; ti < GT; ++ti)
                    *reinterpret_cast<double2*>(colb + 8 * ti + 2 * c4) = make_double2(acc[UIDX(0, ti)][0], acc[UIDX(0, ti)][1]);
                mpub[c4] = mu[0];              // prior mean row 0 (zero on lanes without a mean column)
            }
            if (c4 == (cj1 >> 1)) {
#pragma unroll
                for (int ti = 0; ti < GT; ++ti) colb[R + 8 * ti + g] = e1 ? acc[UIDX(ti, GT - 1)][1] : acc[UIDX(ti, GT - 1)][0];
            }
            if (g == cj1) mpub[4 + c4] = mu[GT - 1];   // prior mean row N-1
        }
        __syncwarp();   // every lane is done reading C / M^T (P1); published columns and mean rows visible

        if (is_valid) {
            // C' w once per row instead of once per use: lane i < R combines the two published columns in place
            // (colb[i] <- w0 C'[i][0] + w1 C'[i][N-1]); every later use - the gain, the column pairs of the rank-1 update, S -
            // reads the finished vector.  Scalar FP64 instructions queue behind the other warps' DMMAs in the one FP64 pipe
            // (ncu: 50 of them cost 30 % of the warp time of k_mmar<3>), so their NUMBER is what matters: 50 -> 31 per frame.
            if (lane < R) colb[lane] = fma(w1, colb[R + lane], w0 * colb[lane]);
            __syncwarp();
            // S = s2 + w^T C' w = s2 + w0 (C' w)[0] + w1 (C' w)[N-1]  (pyx:55-63), then 1/S
            const double Sinv = rcp3(fma(w1, colb[8 * (GT - 1) + cj1], fma(w0, colb[0], s2)));
            double kr[GT];
#pragma unroll
            for (int ti = 0; ti < GT; ++ti) kr[ti] = colb[8 * ti + g] * Sinv;   // K = C' w / S (pyx:66-67)
#pragma unroll
            for (int tjj = 0; tjj < GT; ++tjj) {
                const double2 cw = *reinterpret_cast<const double2*>(colb + 8 * tjj + 2 * c4);   // (C' w)[column pair]
#pragma unroll
                for (int ti = 0; ti <= tjj; ++ti) {
                    acc[UIDX(ti, tjj)][0] = fma(-kr[ti], cw.x, acc[UIDX(ti, tjj)][0]);   // pyx:71-75
                    acc[UIDX(ti, tjj)][1] = fma(-kr[ti], cw.y, acc[UIDX(ti, tjj)][1]);
                }
            }
            // innovation (pyx:79) and mean update (pyx:82-85) of this lane's dimension
            const double xm = x - fma(w1, mpub[4 + c4], w0 * mpub[c4]);
            quad = fma(xm * xm, Sinv, quad);
#pragma unroll
            for (int ti = 0; ti < GT; ++ti) mu[ti] = fma(kr[ti], xm, mu[ti]);
            // running product of Sinv with the exponent split off (no overflow over thousands of frames)
            lmant *= Sinv;
            const int ex = ((__double2hiint(lmant) >> 20) & 0x7ff) - 1023;
            lmant = __hiloint2double(__double2hiint(lmant) - (ex << 20), __double2loint(lmant));
            lexp += ex;
        }
        // ---------------- C+ and M+^T become the operands of the next propagation: upper tiles as accumulator
        //                  pairs, strictly-upper tiles also mirrored; spare rows of the last block are not touched
        if (t + 1 < T) {
#pragma unroll
            for (int ti = 0; ti < GT; ++ti) {
#pragma unroll
                for (int tj = ti; tj < GT; ++tj) {
                    const double v0 = acc[UIDX(ti, tj)][0], v1 = acc[UIDX(ti, tj)][1];
                    if (tj > ti) {   // C[8 tj + 2 c4 + e][8 ti + g] = C[8 ti + g][8 tj + 2 c4 + e]
                        if (tj < GT - 1 || mir0_ok) Cb[offMir + (8 * tj) * LD + 8 * ti] = v0;
                        if (tj < GT - 1 || mir1_ok) Cb[offMir + (8 * tj + 1) * LD + 8 * ti] = v1;
                    }
                    if (ti < GT - 1 || lastrow_ok) *reinterpret_cast<double2*>(Cb + offP + 8 * ti * LD + 8 * tj) = make_double2(v0, v1);
                }
                if (hasq) Cb[offM + 8 * ti] = mu[ti];
            }
        }
        __syncwarp();   // C+ / M+^T complete before the next frame's fragment loads
    }

    // logL = -1/2 [ sum xmm^2 Sinv - ncols * sum_t log Sinv_t + nvalid * ncols * log 2 pi ]   (pyx:88, 251-256)
    quad += __shfl_xor_sync(0xffffffffu, quad, 1);   // lanes 0..3 hold the per-dimension sums
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
