// k_mmar2 / k_mmar8 - the register-chained FP64 tensor-core filter (k_mmar) with SEVERAL WARPS PER FILTER:
//   33 <= N <= 56  (GT 5..7; the north-star shape N = 50 is GT = 7, r = 2)   two warps, up to 6 filters per CTA     k_mmar2<GT, MAXF, MX, 2>
//   57 <= N <= 64  (GT = 8)                                                  four warps (complementary row pairs)   k_mmar2<8, 2, MX, 4>
//   65 <= N <= 72  (GT = 9)                                                  five warps (four pairs + middle row)   k_mmar2<9, 2, MX, 5>
//   73 <= N <= 104 (GT 10..13; BASELINE configs[2] N = 100)                  eight warps, one filter per CTA, one resident propagator  k_mmar8<GT, MX>
// All of them run the same frame, mmar2_run<GT, ROLE, MX, NW, BONE>; only the row -> warp tables (Mmar2Rows) differ.
// MX: N mod 8 in {0, 5, 6, 7}, the mean in an extra row block (bildk_mmar.cuh).
//
// Why: ncu of k_mma2<7> on the N = 50 target (profiles/r02_ncu_n50.txt, source view) shows the tensor pipe 79 % busy
// and each warp away from it 40-55 % of its time: k_mma2 splits the work of a frame by tile COLUMNS - P1 (T = B C, in
// place in shared memory) 28 : 21 tiles, a pair barrier, P2 (C' = T B) 10 : 18 tiles, a pair barrier, update, a pair
// barrier - so each warp waits ~20 % of the frame for its partner at the barrier in the middle, T makes a round trip
// through shared memory, and the three warps per scheduler cannot cover each other.
//
// Here the pair splits the frame by tile ROWS, as k_mmar does inside one warp: row block ti of T = B_s [C | M] is
// produced in registers (GT tiles) and consumed on the spot by the upper tiles of row ti of C' = T B_s + Sig.  Rows are
// independent, so there is NO barrier between the two products; the rows are dealt out so that the DMMA counts match
// (GT = 7: rows {0,1,2} = 39 tile products, rows {3,4,5,6} = 38).  Per frame: two pair barriers (everybody done reading
// C / published columns visible; C+ complete), no T traffic.  Fragment layout, the permuted last tile column (even slots
// = covariance columns, odd slots = mean columns), the swizzle and the update are those of k_mmar (bildk_mmar.cuh).
#pragma once
#include "bildk_mmar.cuh"

namespace bildk {

struct R2Params {
    RParams r;
    int FPC2;   // filters per CTA (NW warps each)
};

// NW = warps per filter: 2 for GT 5..7; 4 for GT = 8 (N = 57..64), where the four COMPLEMENTARY row pairs (ti, GT-1-ti) cost
// exactly the same - 2 GTC tile products in P1 and GT + 1 in P2 each - and the four warps of a filter sit on the four schedulers;
// 5 for GT = 9 (N = 65..72): four pairs and the middle row, two filters per CTA so that the schedulers carry 70 | 70 | 56 | 56.
template <int GT, int NW = 2>
struct Mmar2Rows {
    // role that owns tile-row block ti; cost of a row = GT (P1) + GT - ti (P2) tile products
    // (with the mean in an extra tile column, MX, every row costs one more: the same splits stay the balanced ones)
    // NW = 8 (GT = 10..13, N = 73..104; one filter per CTA, warp w on scheduler w % 4): rows dealt out so that the four schedulers
    // carry (almost) equal numbers of tile products (tables below)
    __host__ __device__ static constexpr int role(int ti) {
        return NW == 8 ? role8(ti)
             : NW >= 4 ? (ti < GT - 1 - ti ? ti : GT - 1 - ti)   // GT = 8: rows {0,7} {1,6} {2,5} {3,4}: 25 tile products each (MX: 27);
                                                                 // GT = 9 (NW = 5): {0,8} {1,7} {2,6} {3,5} 28 (30) each and the middle row {4} 14 (15)
             : GT == 7 ? (ti <= 2 ? 0 : 1)                       // 14+13+12 = 39 | 11+10+9+8 = 38      MX: 42 | 42
             : GT == 6 ? ((ti == 0 || ti == 3 || ti == 5) ? 0 : 1)   // 12+9+7 = 28 | 11+10+8 = 29          MX: 31 | 32
             : (ti <= 1 ? 0 : 1);                                // GT = 5: 10+9 = 19 | 8+7+6 = 21      MX: 21 | 24
    }
    // NW = 8: closed forms (a role owns one or two rows).  The loops below fold to constants for two and four roles, but with
    // thirteen rows and eight roles the optimiser gave up on them: run-time indexed accumulators in local memory, jump tables,
    // a 40-minute compilation and a kernel 30x slower than the one it was to replace.
    // row tables of the eight-warp kernel (one or two rows per warp; warps w and w + 4 share a scheduler), found by a greedy search
    // over the per-scheduler sums of the row costs GTC + GT - ti (tools/mmar8_tables.py): per-scheduler maximum / mean
    //   GT = 10: {6} {5,9} {3} {1} | {0} {2} {4} {7,8}          44 / 38.75      GT = 12: {7,10} {4,9} {2} {1} | {0} {3} {6,8} {5,11}   56 / 55.5
    //   GT = 11: {5} {0} {1} {2} | {3,8} {7,9} {6,10} {4}       50 / 46.75      GT = 13: {0,12} {3,9} {4,6} {7,11} | {1} {2} {5} {8,10} 68 / 65
    __host__ __device__ static constexpr int first8(int r) {
        return GT == 10 ? (r == 0 ? 6 : r == 1 ? 5 : r == 2 ? 3 : r == 3 ? 1 : r == 4 ? 0 : r == 5 ? 2 : r == 6 ? 4 : 7)
             : GT == 11 ? (r == 0 ? 5 : r == 1 ? 0 : r == 2 ? 1 : r == 3 ? 2 : r == 4 ? 3 : r == 5 ? 7 : r == 6 ? 6 : 4)
             : GT == 12 ? (r == 0 ? 7 : r == 1 ? 4 : r == 2 ? 2 : r == 3 ? 1 : r == 4 ? 0 : r == 5 ? 3 : r == 6 ? 6 : 5)
             : (r == 0 ? 0 : r == 1 ? 3 : r == 2 ? 4 : r == 3 ? 7 : r == 4 ? 1 : r == 5 ? 2 : r == 6 ? 5 : 8);
    }
    __host__ __device__ static constexpr int second8(int r) {
        return GT == 10 ? (r == 1 ? 9 : r == 7 ? 8 : -1)
             : GT == 11 ? (r == 4 ? 8 : r == 5 ? 9 : r == 6 ? 10 : -1)
             : GT == 12 ? (r == 0 ? 10 : r == 1 ? 9 : r == 6 ? 8 : r == 7 ? 11 : -1)
             : (r == 0 ? 12 : r == 1 ? 9 : r == 2 ? 6 : r == 3 ? 11 : r == 7 ? 10 : -1);
    }
    __host__ __device__ static constexpr bool owns8(int r, int ti) { return ti == first8(r) || ti == second8(r); }
    __host__ __device__ static constexpr int role8(int ti) {
        return owns8(0, ti) ? 0 : owns8(1, ti) ? 1 : owns8(2, ti) ? 2 : owns8(3, ti) ? 3 : owns8(4, ti) ? 4 : owns8(5, ti) ? 5 : owns8(6, ti) ? 6 : 7;
    }
    // NW = 4, 5: complementary pairs - role r owns rows r and GT-1-r (the middle row of an odd grid alone)
    static constexpr bool PAIRS = NW == 4 || NW == 5;
    __host__ __device__ static constexpr int nacc(int r) {          // upper tiles owned by role r
        if (NW == 8) return (GT - first8(r)) + (second8(r) >= 0 ? GT - second8(r) : 0);
        if (PAIRS) return (GT - r) + (GT - 1 - r != r ? r + 1 : 0);
        int n = 0;
        for (int ti = 0; ti < GT; ++ti) if (role(ti) == r) n += GT - ti;
        return n;
    }
    __host__ __device__ static constexpr int nrows(int r) {
        if (NW == 8) return second8(r) >= 0 ? 2 : 1;
        if (PAIRS) return GT - 1 - r != r ? 2 : 1;
        int n = 0;
        for (int ti = 0; ti < GT; ++ti) if (role(ti) == r) ++n;
        return n;
    }
    __host__ __device__ static constexpr int first(int r) {          // lowest row of role r
        if (NW == 8) return first8(r);
        if (PAIRS) return r;
        for (int ti = 0; ti < GT; ++ti) if (role(ti) == r) return ti;
        return GT;
    }
    __host__ __device__ static constexpr int aidx(int ti, int tj) {  // accumulator slot of upper tile (ti, tj) within its owner
        if (NW == 8) return (ti == first8(role(ti)) ? 0 : GT - first8(role(ti))) + tj - ti;
        if (PAIRS) return (ti <= GT - 1 - ti ? 0 : GT - role(ti)) + tj - ti;
        int n = 0;
        for (int t = 0; t < ti; ++t) if (role(t) == role(ti)) n += GT - t;
        return n + tj - ti;
    }
    __host__ __device__ static constexpr int ridx(int ti) {          // slot of row ti within its owner (mu, kr)
        if (NW == 8) return ti == first8(role(ti)) ? 0 : 1;
        if (PAIRS) return ti <= GT - 1 - ti ? 0 : 1;
        int n = 0;
        for (int t = 0; t < ti; ++t) if (role(t) == role(ti)) ++n;
        return n;
    }
};

// one k-tile of P1 for this warp's tile rows: the GTC fragments of [C | M] are loaded once and serve all of them
template <int GT, int ROLE, bool MX, int NW>
__device__ __forceinline__ void mmar2_p1_ktile(double (&Tt)[Mmar2Rows<GT, NW>::nrows(ROLE)][MmarGeom<GT, MX>::GTC][2], const double* __restrict__ Bs,
                                               const double* __restrict__ Cb, int offP, int offLP, int kt) {
    using RW = Mmar2Rows<GT, NW>;
    constexpr int LD = MmarGeom<GT, MX>::LD, GTC = MmarGeom<GT, MX>::GTC;
    double2 b[GTC];
#pragma unroll
    for (int tj = 0; tj < GTC; ++tj)
        b[tj] = *reinterpret_cast<const double2*>(tj < GTC - 1 ? Cb + offP + 8 * tj * LD + 8 * kt : Cb + offLP + 8 * kt);
#pragma unroll
    for (int ti = 0; ti < GT; ++ti) {
        if (RW::role(ti) != ROLE) continue;
        const double2 a = *reinterpret_cast<const double2*>(Bs + offP + 8 * ti * LD + 8 * kt);
#pragma unroll
        for (int tj = 0; tj < GTC; ++tj) dmma884(Tt[RW::ridx(ti)][tj], a.x, b[tj].x);
#pragma unroll
        for (int tj = 0; tj < GTC; ++tj) dmma884(Tt[RW::ridx(ti)][tj], a.y, b[tj].y);
    }
}

template <int NW>
__device__ __forceinline__ void filter_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(32 * NW) : "memory"); }

// BONE: ONE propagator resident in shared memory (two do not fit next to the filter for N >= 89); it is swapped by TMA when the
// profile changes state (all readers of the old one have passed the barrier that ends the previous frame).
template <int GT, int ROLE, bool MX, int NW = 2, bool BONE = false>
__device__ __forceinline__ void mmar2_run(const RParams& rp, double* __restrict__ Bsm, double* __restrict__ Cb, int pidx, int tjx,
                                          int e_sub, int barid, int lane, uint64_t* mbar = nullptr) {
    using G = MmarGeom<GT, MX>;
    using RW = Mmar2Rows<GT, NW>;
    constexpr int R = G::R, LD = G::LD, MAT = G::MAT, MATC = G::MATC;
    constexpr int KT = G::KT, GTC = G::GTC;
    constexpr int NACC = RW::nacc(ROLE), NROW = RW::nrows(ROLE);
    constexpr bool OWN_FIRST = RW::role(0) == ROLE, OWN_LAST = RW::role(GT - 1) == ROLE;
    constexpr int FIRSTROW = RW::first(ROLE);
#define MINE(ti) (RW::role(ti) == ROLE)
#define AIDX(ti, tjj) (RW::aidx(ti, tjj))
#define RIDX(ti) (RW::ridx(ti))
    const KParams& p = rp.k;
    const int g = lane >> 2, c4 = lane & 3;
    const int N = p.N, D = p.D;
    double* const colb = Cb + MATC;             // [2][R] the two columns of C' that w touches
    double* const mpub = colb + 2 * R;          // [2][4] prior mean rows 0 and N - 1
    double* const cwb = mpub + 8 + ROLE * R;    // [R] this warp's copy of C' w

    const int T = p.T[tjx];
    const double* __restrict__ xg = p.x[tjx];
    const uint32_t* __restrict__ vbits = reinterpret_cast<const uint32_t*>(p.valid[tjx] + (T + 3) / 4 * 4);
    uint32_t vword = 0;
    const int ncols = p.ncols[e_sub];
    const double s2 = p.s2[e_sub];
    const double w0 = p.wz_val[0], w1 = p.wz_val[1];
    const int rr = rp.r;
    const int cj1 = rr - 1;                     // column N - 1 = 8 (GT - 1) + cj1
    const bool e1 = cj1 & 1;

    // lane-constant fragment offsets (doubles); rows 8 t + g flip column bit 2 when (g >> 1) & 1  (bildk_mmar.cuh)
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
    const bool lastrow_ok = g < rr;
    const bool mir0_ok = 2 * c4 < rr, mir1_ok = 2 * c4 + 1 < rr;

    double quad = 0.0, lmant = 1.0;             // role 1 accumulates the log-likelihood
    int lexp = 0;

    int r_cur = 0;
    int s = p.run_states[static_cast<size_t>(pidx) * p.K1];
    int next_sw = (p.K1 > 1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + 1] : 0x7fffffff;

    double acc[NACC][2];
    double mu[NROW];
    int s_loaded = -1;
    uint32_t bphase = 0;

    for (int t = 0; t < T; ++t) {
        while (t >= next_sw) {
            ++r_cur;
            s = p.run_states[static_cast<size_t>(pidx) * p.K1 + r_cur];
            next_sw = (r_cur + 1 < p.K1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + r_cur + 1] : 0x7fffffff;
        }
        if ((t & 31) == 0) vword = __ldg(vbits + (t >> 5));
        const bool is_valid = (vword >> (t & 31)) & 1u;

        if (BONE && t > 0 && s != s_loaded) {   // the same decision in every warp of the filter
            if (ROLE == 0 && lane == 0) tma_stage(Bsm, rp.Br + static_cast<size_t>(MAT) * s, MAT * sizeof(double), mbar);
            mbar_wait(mbar, bphase);
            bphase ^= 1;
            s_loaded = s;
        }
        if (t > 0) {
            const double* __restrict__ Bs = BONE ? Bsm : Bsm + s * MAT;
            const double* __restrict__ Gs = rp.Sigm + static_cast<size_t>(R * R) * s + g * R + 2 * c4;
            // ---------------- P1, this warp's tile-row blocks: T[ti][:] = B_s[ti][:] [C | M]   (MSRouse_logL.pyx:206-241)
            // k-tile outermost: the GT fragments of [C | M] of a k-tile are loaded once and serve all of this warp's rows
            double Tt[NROW][GTC][2];
#pragma unroll
            for (int r = 0; r < NROW; ++r)
#pragma unroll
                for (int tj = 0; tj < GTC; ++tj) Tt[r][tj][0] = Tt[r][tj][1] = 0.0;
            if constexpr (NW == 8) {   // eight roles: the k-tile loop of P1 stays rolled (code size; T is indexed by tile column only)
#pragma unroll 1
                for (int kt = 0; kt < KT; ++kt) mmar2_p1_ktile<GT, ROLE, MX, NW>(Tt, Bs, Cb, offP, offLP, kt);
            } else {
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) mmar2_p1_ktile<GT, ROLE, MX, NW>(Tt, Bs, Cb, offP, offLP, kt);
            }
            if constexpr (!MX) {
                double b[GT];
#pragma unroll
                for (int tj = 0; tj < GT; ++tj) b[tj] = tj < GT - 1 ? Cb[offS + 8 * tj * LD] : Cb[offLS];
#pragma unroll
                for (int ti = 0; ti < GT; ++ti) {
                    if (!MINE(ti)) continue;
                    const double a = Bs[offS + 8 * ti * LD];
#pragma unroll
                    for (int tj = 0; tj < GT; ++tj) dmma884(Tt[RIDX(ti)][tj], a, b[tj]);
                }
            }
            // ---------------- P2, upper tiles of this warp's rows: C'[ti][tj] = Sig + T[ti][:] B_s[:][tj]; again k-tile outermost
#pragma unroll
            for (int ti = 0; ti < GT; ++ti) {
                if (!MINE(ti)) continue;
                mu[RIDX(ti)] = Tt[RIDX(ti)][GTC - 1][1];   // M'[8 ti + g][c4]
#pragma unroll
                for (int tj = ti; tj < GT; ++tj) {
                    const double2 v = __ldg(reinterpret_cast<const double2*>(Gs + 8 * ti * R + 8 * tj));
                    acc[AIDX(ti, tj)][0] = v.x;
                    acc[AIDX(ti, tj)][1] = v.y;
                }
            }
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                double2 b[GT];
#pragma unroll
                for (int tj = FIRSTROW; tj < GT; ++tj) b[tj] = *reinterpret_cast<const double2*>(Bs + offP + 8 * tj * LD + 8 * kt);
#pragma unroll
                for (int ti = 0; ti < GT; ++ti) {
                    if (!MINE(ti)) continue;
#pragma unroll
                    for (int tj = ti; tj < GT; ++tj) dmma884(acc[AIDX(ti, tj)], Tt[RIDX(ti)][kt][0], b[tj].x);
#pragma unroll
                    for (int tj = ti; tj < GT; ++tj) dmma884(acc[AIDX(ti, tj)], Tt[RIDX(ti)][kt][1], b[tj].y);
                }
            }
            if constexpr (!MX) {
                double b[GT];
#pragma unroll
                for (int tj = FIRSTROW; tj < GT; ++tj) b[tj] = Bs[offS + 8 * tj * LD];
#pragma unroll
                for (int ti = 0; ti < GT; ++ti) {
                    if (!MINE(ti)) continue;
#pragma unroll
                    for (int tj = ti; tj < GT; ++tj) dmma884(acc[AIDX(ti, tj)], Tt[RIDX(ti)][GT - 1][0], b[tj]);
                }
            }
            if (p.hasG) {   // M' = B M + G  (pyx:209-214)
#pragma unroll
                for (int ti = 0; ti < GT; ++ti)
                    if (MINE(ti) && hasq && 8 * ti + g < N) mu[RIDX(ti)] += __ldg(p.Gm + (s * N + 8 * ti + g) * D + xcol);
            }
        } else {
            // frame 0: steady state of the first state (pyx:160-163), no propagation
            const double* __restrict__ Gs = rp.C0m + static_cast<size_t>(R * R) * s + g * R + 2 * c4;
#pragma unroll
            for (int ti = 0; ti < GT; ++ti) {
                if (!MINE(ti)) continue;
#pragma unroll
                for (int tj = ti; tj < GT; ++tj) {
                    const double2 v = __ldg(reinterpret_cast<const double2*>(Gs + 8 * ti * R + 8 * tj));
                    acc[AIDX(ti, tj)][0] = v.x;
                    acc[AIDX(ti, tj)][1] = v.y;
                }
                mu[RIDX(ti)] = (hasq && 8 * ti + g < N) ? __ldg(p.M0 + (s * N + 8 * ti + g) * D + xcol) : 0.0;
            }
        }

        double x = 0.0;
        if (is_valid) {
            if (hasq) x = __ldg(xg + t * D + xcol);
            // publish the two columns of C' that w touches (bildk_mmar.cuh): column 0 comes from tile row 0 alone (its owner
            // publishes it), column N - 1 from the last tile of every row (each warp publishes its rows)
            if (OWN_FIRST) {
                if (c4 == 0) colb[g] = acc[AIDX(0, 0)][0];
                if (g == 0) {
#pragma unroll
                    for (int ti = 1; ti < GT; ++ti)
                        *reinterpret_cast<double2*>(colb + 8 * ti + 2 * c4) = make_double2(acc[AIDX(0, ti)][0], acc[AIDX(0, ti)][1]);
                    mpub[c4] = mu[RIDX(0)];
                }
            }
            if (c4 == (cj1 >> 1)) {
#pragma unroll
                for (int ti = 0; ti < GT; ++ti)
                    if (MINE(ti)) colb[R + 8 * ti + g] = e1 ? acc[AIDX(ti, GT - 1)][1] : acc[AIDX(ti, GT - 1)][0];
            }
            if (OWN_LAST && g == cj1) mpub[4 + c4] = mu[RIDX(GT - 1)];
        }
        filter_sync<NW>(barid);   // both warps are done reading C / M^T; published columns and mean rows visible

        if (is_valid) {
            // C' w once per row (bildk_mmar.cuh): each warp combines the two published columns into its own copy of the vector
            // (a shared copy would need a third pair barrier), then reads gains and column pairs from it
#pragma unroll
            for (int i = lane; i < R; i += 32) cwb[i] = fma(w1, colb[R + i], w0 * colb[i]);
            __syncwarp();
            // S = s2 + w^T C' w = s2 + w0 (C' w)[0] + w1 (C' w)[N-1]  (pyx:55-63), then 1/S
            const double Sinv = rcp3(fma(w1, cwb[8 * (GT - 1) + cj1], fma(w0, cwb[0], s2)));
            double kr[NROW];
#pragma unroll
            for (int ti = 0; ti < GT; ++ti)
                if (MINE(ti)) kr[RIDX(ti)] = cwb[8 * ti + g] * Sinv;   // K = C' w / S (pyx:66-67)
#pragma unroll
            for (int tjj = FIRSTROW; tjj < GT; ++tjj) {
                const double2 cw = *reinterpret_cast<const double2*>(cwb + 8 * tjj + 2 * c4);   // (C' w)[column pair]
#pragma unroll
                for (int ti = 0; ti <= tjj; ++ti) {
                    if (!MINE(ti)) continue;
                    acc[AIDX(ti, tjj)][0] = fma(-kr[RIDX(ti)], cw.x, acc[AIDX(ti, tjj)][0]);   // pyx:71-75
                    acc[AIDX(ti, tjj)][1] = fma(-kr[RIDX(ti)], cw.y, acc[AIDX(ti, tjj)][1]);
                }
            }
            // innovation (pyx:79) and mean update (pyx:82-85) of this lane's dimension
            const double xm = x - fma(w1, mpub[4 + c4], w0 * mpub[c4]);
#pragma unroll
            for (int ti = 0; ti < GT; ++ti)
                if (MINE(ti)) mu[RIDX(ti)] = fma(kr[RIDX(ti)], xm, mu[RIDX(ti)]);
            if (ROLE == 1) {
                quad = fma(xm * xm, Sinv, quad);
                lmant *= Sinv;   // running product of Sinv with the exponent split off
                const int ex = ((__double2hiint(lmant) >> 20) & 0x7ff) - 1023;
                lmant = __hiloint2double(__double2hiint(lmant) - (ex << 20), __double2loint(lmant));
                lexp += ex;
            }
        }
        // ---------------- C+ and M+^T become the operands of the next propagation
        if (t + 1 < T) {
#pragma unroll
            for (int ti = 0; ti < GT; ++ti) {
                if (!MINE(ti)) continue;
#pragma unroll
                for (int tj = ti; tj < GT; ++tj) {
                    const double v0 = acc[AIDX(ti, tj)][0], v1 = acc[AIDX(ti, tj)][1];
                    if (tj > ti) {   // C[8 tj + 2 c4 + e][8 ti + g] = C[8 ti + g][8 tj + 2 c4 + e]
                        if (tj < GT - 1 || mir0_ok) Cb[offMir + (8 * tj) * LD + 8 * ti] = v0;
                        if (tj < GT - 1 || mir1_ok) Cb[offMir + (8 * tj + 1) * LD + 8 * ti] = v1;
                    }
                    if (ti < GT - 1 || lastrow_ok) *reinterpret_cast<double2*>(Cb + offP + 8 * ti * LD + 8 * tj) = make_double2(v0, v1);
                }
                if (hasq) Cb[offM + 8 * ti] = mu[RIDX(ti)];
            }
        }
        filter_sync<NW>(barid);   // C+ / M+^T complete before the next frame's fragment loads
    }

    if (ROLE == 1) {
        // logL = -1/2 [ sum xmm^2 Sinv - ncols * sum_t log Sinv_t + nvalid * ncols * log 2 pi ]   (pyx:88, 251-256)
        quad += __shfl_xor_sync(0xffffffffu, quad, 1);
        quad += __shfl_xor_sync(0xffffffffu, quad, 2);
        if (lane == 0) {
            int nvalid = 0;
            for (int wv = 0; wv < (T + 31) / 32; ++wv) nvalid += __popc(__ldg(vbits + wv));
            const double logdet = log(lmant) + lexp * 0.6931471805599453;
            p.out[static_cast<size_t>(e_sub) * p.P + pidx] = -0.5 * (quad - ncols * logdet + static_cast<double>(nvalid) * ncols * LOG_2PI);
        }
    }
#undef MINE
#undef AIDX
#undef RIDX
}

// MAXF: filters per CTA the kernel is compiled for (registers per thread = 65536 / (64 MAXF)).
template <int GT, int MAXF, bool MX, int NW>
__device__ __forceinline__ void mmar2_kernel(const R2Params& rp2) {
    constexpr int MAT = MmarGeom<GT, MX>::MAT;
    const RParams& rp = rp2.r;
    const KParams& p = rp.k;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    double* Bsm = reinterpret_cast<double*>(smem_raw + 16);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int fl = wid / NW;                          // filter within the CTA
    // NW = 2: roles alternate so that every scheduler holds warps of either role; NW = 4: warp w of a filter runs on scheduler w
    const int role = NW == 2 ? ((wid & 1) ^ ((fl >> 1) & 1)) : (wid % NW);
    const int tjx = p.cta_traj ? p.cta_traj[blockIdx.x] : 0;
    const int first = p.cta_first ? p.cta_first[blockIdx.x] : blockIdx.x * rp2.FPC2;
    const int pend = p.traj_first[tjx + 1];
    const int pidx = first + fl;
    const bool alive = (fl < rp2.FPC2) && (pidx < pend);

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
    if (!alive) return;   // both warps of a pair leave together; the named barriers below are per pair
    double* Cb = Bsm + MAT * p.S + fl * rp.fstride;
    // padding columns, zero rows and M^T rows must start finite / zero: the pair clears its filter's buffer
    for (int i = (wid % NW) * 32 + lane; i < rp.fstride; i += 32 * NW) Cb[i] = 0.0;
    filter_sync<NW>(1 + fl);
    mbar_wait(mbar, 0);
    if (role == 0) mmar2_run<GT, 0, MX, NW>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, 1 + fl, lane);
    else if (NW == 2 || role == 1) mmar2_run<GT, 1, MX, NW>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, 1 + fl, lane);
    else if (NW == 3 || role == 2) mmar2_run<GT, (NW >= 3 ? 2 : 0), MX, NW>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, 1 + fl, lane);
    else if (NW == 4 || role == 3) mmar2_run<GT, (NW >= 4 ? 3 : 0), MX, NW>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, 1 + fl, lane);
    else mmar2_run<GT, (NW >= 5 ? 4 : 0), MX, NW>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, 1 + fl, lane);
}

template <int GT, int MAXF, bool MX = false, int NW = 2>
__global__ void __launch_bounds__(32 * NW * MAXF, 1) k_mmar2(const __grid_constant__ R2Params rp2) {
    mmar2_kernel<GT, MAXF, MX, NW>(rp2);
}
// (GT = 9 runs ten warps per CTA: the register file is handed out to warps in groups of four, so ptxas budgets 320 threads like
// 384 - 168 registers, ~1 KB of spills per thread.  Compiling for 200 registers with __maxnreg__ halves the spills but the launch
// fails with "too many resources requested": measured, reverted.)

// k_mmar8 - the register-chained scheme for GT = 13 (N = 97..104; BASELINE configs[2] N = 100): ONE filter per CTA, eight warps,
// the tile rows dealt out so that the four schedulers carry (almost) equal DMMA counts (Mmar2Rows<GT, 8>), one resident
// propagator swapped by TMA at a state switch.  Against k_mmact (one CTA per filter as well, T through shared memory, the
// phases separated by CTA barriers; tensor pipe 68 % busy): no barrier between the two products, two CTA barriers per frame, a
// third of the shared-memory traffic.
template <int GT, int ROLE, bool MX>
__device__ __forceinline__ void mmar8_role(const RParams& rp, double* __restrict__ Bsm, double* __restrict__ Cb, int pidx, int tjx, int e_sub,
                                        int lane, uint64_t* mbar) {
    mmar2_run<GT, ROLE, MX, 8, true>(rp, Bsm, Cb, pidx, tjx, e_sub, 1, lane, mbar);
}

template <int GT, bool MX>
__global__ void __launch_bounds__(256, 1) k_mmar8(const __grid_constant__ R2Params rp2) {
    constexpr int MAT = MmarGeom<GT, MX>::MAT;
    const RParams& rp = rp2.r;
    const KParams& p = rp.k;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    double* Bsm = reinterpret_cast<double*>(smem_raw + 16);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tjx = p.cta_traj ? p.cta_traj[blockIdx.x] : 0;
    const int pidx = p.cta_first ? p.cta_first[blockIdx.x] : blockIdx.x;
    if (pidx >= p.traj_first[tjx + 1]) return;
    if (tid == 0) mbar_init(mbar, 1);
    double* Cb = Bsm + MAT;
    for (int i = tid; i < rp.fstride; i += 256) Cb[i] = 0.0;   // padding columns, zero rows and M^T rows must start finite / zero
    __syncthreads();
    switch (wid) {   // warp w runs on scheduler w % 4: roles w and w + 4 share one
        case 0: mmar8_role<GT, 0, MX>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, lane, mbar); break;
        case 1: mmar8_role<GT, 1, MX>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, lane, mbar); break;
        case 2: mmar8_role<GT, 2, MX>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, lane, mbar); break;
        case 3: mmar8_role<GT, 3, MX>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, lane, mbar); break;
        case 4: mmar8_role<GT, 4, MX>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, lane, mbar); break;
        case 5: mmar8_role<GT, 5, MX>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, lane, mbar); break;
        case 6: mmar8_role<GT, 6, MX>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, lane, mbar); break;
        default: mmar8_role<GT, 7, MX>(rp, Bsm, Cb, pidx, tjx, blockIdx.y, lane, mbar); break;
    }
}

}  // namespace bildk
