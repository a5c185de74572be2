// FP64 tensor-core (DMMA m8n8k4) variant of the filter kernel: ONE WARP PER FILTER.
//
// Why: tools/lds_patterns.cu + the ncu capture of the DFMA tile kernel (profiles/) show that kernel bound
// by shared-memory operand delivery (0.4 doubles loaded per FMA with 5x5 register tiles).  A DMMA
// consumes 2 doubles per lane for 8 FMAs per lane and fragments are reused across the warp's tile
// block, so a 3x3-tile filter (N <= 24) needs 0.083 doubles per FMA; the FP64 pipe becomes the limit.
//
// Per frame (MSRouse_logL.pyx:203-248), with the covariance as GT x GT tiles of 8 x 8 in registers
// (lane l holds D[l>>2][2*(l&3) + {0,1}] of every tile):
//   P1   T    = B_s [C | M]      A-fragments from B_s (row-major), B-fragments from the C buffer; the mean
//                                columns ride in the padding columns N..N+d-1 of the last tile column
//                                (or in one extra tile column when 8*GT - N < d), so M' = B_s M is free
//   ---- T (and with it M') overwrites the C buffer, same row-major layout
//   P2   C'   = T B_s + Sig      A-fragments from the buffer, B-fragments from B_s, accumulators start at Sig
//   upd  rank-1 measurement update in the fragment layout (pyx:19-90); sparse w only
//   ---- C+ (with M+ merged into its columns) overwrites the buffer
// Shared-memory layout: row stride a multiple of 128 bytes, 32-byte column groups XOR-swizzled by the
// row ( group' = group ^ swz(row), swz(r) = ((r&1)<<1) | ((r>>1)&1) ).  Fragment loads (8 rows x 32 B or
// 4 rows x 64 B per half warp) and accumulator-pair stores (2 rows x 64 B per quarter warp) are then all
// conflict-free; the first version (stride == 4 mod 8, no swizzle) had ideal loads but 2x the ideal
// wavefronts on every store (ncu: profiles/r01_ncu_c2_mma_v2.txt, tools/lds_patterns.cu).
#pragma once
#include "bildk_kernels.cuh"

#ifndef BILDK_MMA_SWZ
#define BILDK_MMA_SWZ 1   // XOR-swizzled shared layout for GT <= 4 (conflict-free stores); measured +1-2% over the plain stride
#endif

namespace bildk {

struct MParams {
    KParams k;                 // shared fields (model sizes, trajectories, batch); tile-layout fields unused
    int NPm;                   // 8*GT
    int LDB;                   // row stride of B (global pre-swizzled copy and shared), multiple of 16
    int LDC;                   // row stride of the per-filter buffer, multiple of 16, >= 8*GTC
    int MC0;                   // first mean column inside the buffer
    int NK;                    // 4*ceil(N/4): contraction length
    const double* Bm;          // [S][NPm][LDB] zero padded, swizzled like the shared copy
    const double* Sigm;        // [S][NPm][NPm] zero padded, plain row-major
    const double* C0m;         // [S][NPm][NPm]
    int WPC;                   // warps (= filters) per CTA
    int fstride_m;             // doubles of shared memory per filter
    int bstride_m;             // doubles per state of B in shared memory
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// GT: tiles per edge; MX: mean columns live in an extra tile column.  The measurement vector has exactly
// two non-zeros (BILD's end-to-end distance, models.py:230-233); other vectors use the tile kernel.
//
// Symmetry: C' = B C B + Sig is symmetric, so P2 computes only the NU = GT (GT+1)/2 upper tiles (ti <= tj)
// and mirrors them when the posterior is written back: 3 N^3 instead of 4 N^3 executed flops, and a
// third fewer accumulator registers.  T = B [C | M] is not symmetric; it is produced IN PLACE by column
// blocks of CB = (GT+1)/2 tile columns (block J reads only C[:, J] and overwrites it with T[:, J]), which
// needs GT * CB <= NU accumulators - the same registers.
//
// Strides are compile-time so that all fragment addressing folds into immediates.  For GT <= 4 the row
// stride is a multiple of 128 bytes with XOR-swizzled 32-byte groups (all loads and stores conflict-free);
// for larger GT the stride is == 4 (mod 8) doubles without swizzle (ideal fragment loads, 2x wavefronts on
// the comparatively rare stores) because the padded stride would cost a resident filter per SM.
// Register budget: GT <= 3 must keep 28 warps per SM resident (BASELINE config 2 is 27.7 filters per SM).
template <int GT, bool MX>
__global__ void __launch_bounds__(GT <= 3 ? 896 : 256, GT == 4 ? 2 : 1) k_mma(const __grid_constant__ MParams mp) {
    constexpr int GTC = GT + (MX ? 1 : 0);
    constexpr int TJM = MX ? GT : GT - 1;   // tile column that contains the mean columns
    constexpr bool SWZ = BILDK_MMA_SWZ && GT <= 4;
    constexpr int NPm = 8 * GT;
    constexpr int LDB = SWZ ? (NPm + 15) / 16 * 16 : NPm + 4;
    constexpr int LDC = SWZ ? (8 * GTC + 15) / 16 * 16 : 8 * GTC + 4;
    constexpr int MATB = NPm * LDB, MATG = NPm * NPm;
    constexpr int NU = GT * (GT + 1) / 2;          // upper tiles
    constexpr int CB = (GT + 1) / 2;               // tile columns per P1 pass
    constexpr int NPASS = (GTC + CB - 1) / CB;
    static_assert(GT * CB <= NU || GT == 2, "P1 pass must fit the accumulator file");
    constexpr int NACC = (GT * CB > NU) ? GT * CB : NU;
#define UIDX(ti, tjj) ((ti) * GT - (ti) * ((ti) - 1) / 2 + ((tjj) - (ti)))
    const KParams& p = mp.k;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    double* Bsm = reinterpret_cast<double*>(smem_raw + 16);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, c4 = lane & 3;
    const int e_sub = blockIdx.y;
    const int N = p.N, D = p.D, NK = mp.NK;

    const int tj = p.cta_traj ? p.cta_traj[blockIdx.x] : 0;
    const int first = p.cta_first ? p.cta_first[blockIdx.x] : blockIdx.x * mp.WPC;
    const int pend = p.traj_first[tj + 1];
    const int pidx = first + wid;
    const bool alive = (wid < mp.WPC) && (pidx < pend);

    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(mbar, static_cast<uint32_t>(MATB * p.S * sizeof(double)));
        for (int st = 0; st < p.S; ++st) {
            constexpr uint32_t CH = 32768;
            constexpr uint32_t bytes = MATB * sizeof(double);
            for (uint32_t off = 0; off < bytes; off += CH)
                tma_load_1d(reinterpret_cast<char*>(Bsm + st * MATB) + off, reinterpret_cast<const char*>(mp.Bm + static_cast<size_t>(MATB) * st) + off,
                            bytes - off < CH ? bytes - off : CH, mbar);
        }
    }
    if (!alive) return;   // warps are independent after this point (warp-scope barriers only)

    double* Cb = Bsm + MATB * p.S + wid * mp.fstride_m;   // [NPm][LDC]
    double* colb = Cb + NPm * LDC;                         // [2][NPm]

    const int T = p.T[tj];
    const double* __restrict__ xg = p.x[tj];
    // packed valid-frame bits live behind the byte flags (bildk_traj_create); one word per 32 frames
    const uint32_t* __restrict__ vbits = reinterpret_cast<const uint32_t*>(p.valid[tj] + (T + 3) / 4 * 4);
    uint32_t vword = 0;
    const int ncols = p.ncols[e_sub];
    const double s2 = p.s2[e_sub];

    // the two non-zeros of w
    const int j0 = p.wz_idx[0], j1 = p.wz_idx[1];
    const double w0 = p.wz_val[0], w1 = p.wz_val[1];
    auto swz = [](int r) { return SWZ ? (((r & 1) << 1) | ((r >> 1) & 1)) : 0; };
    // physical column of logical column c in a row whose swizzle is sw
    auto pcol = [](int c, int sw) { return (((c >> 2) ^ sw) << 2) | (c & 3); };

    // mean columns owned by this lane: buffer column MC0 + q  <->  tile TJM, local column 2*c4 + e
    const int q0 = 2 * c4 - (mp.MC0 - 8 * TJM), q1 = q0 + 1;
#define qv0 (static_cast<unsigned>(q0) < static_cast<unsigned>(ncols))
#define qv1 (static_cast<unsigned>(q1) < static_cast<unsigned>(ncols))
#define xc0 (p.cols[e_sub][qv0 ? q0 : 0])
#define xc1 (p.cols[e_sub][qv1 ? q1 : 0])

    // log-likelihood pieces (pyx:88), summed at the end: quad = sum xmm^2 Sinv over this lane's dimensions;
    // sum_t log Sinv_t is kept as log(mantissa product) + exponent sum (one log per filter instead of per frame)
    double quad = 0.0;
    // running product of Sinv (mantissa) and its exponent sum / frame count live in shared memory: they are
    // warp-uniform, touched once per valid frame, and registers are the scarce resource here
    double* const lst = colb + 2 * NPm;          // [0] mantissa  [1] (int2) exponent sum, valid frames
    if (lane == 0) { lst[0] = 1.0; reinterpret_cast<int*>(lst + 1)[0] = 0; reinterpret_cast<int*>(lst + 1)[1] = 0; }

    int r_cur = 0;
    int s = p.run_states[static_cast<size_t>(pidx) * p.K1];
    int next_sw = (p.K1 > 1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + 1] : 0x7fffffff;

    double acc[NACC][2];
    // swizzle constants of this lane: rows 8*ti + g share swz(g); rows k0 + c4 share swz(c4)
    const int sg = swz(g), sc = swz(c4);
    // accumulator pair of tile (ti, tj): row 8 ti + g, logical columns 8 tj + 2 c4 + {0,1} -> physical tile
    // column tj ^ (sg>>1), 32-byte group (c4>>1) ^ (sg&1).  tj ^ 1 is tj + 1 for even and tj - 1 for odd tj:
    double* const dE = Cb + g * LDC + ((c4 >> 1) ^ (sg & 1)) * 4 + (c4 & 1) * 2 + (sg >> 1) * 8;   // even tj
    double* const dO = dE - (sg >> 1) * 16;                                                         // odd tj
#define DPAIR(ti, tjj) (((tjj) & 1 ? dO : dE) + 8 * (ti) * LDC + 8 * (tjj))

    mbar_wait(mbar, 0);

    for (int t = 0; t < T; ++t) {
        while (t >= next_sw) {   // rare: at most K1 - 1 times per filter; pointers are recomputed, not kept
            ++r_cur;
            s = p.run_states[static_cast<size_t>(pidx) * p.K1 + r_cur];
            next_sw = (r_cur + 1 < p.K1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + r_cur + 1] : 0x7fffffff;
        }
        if ((t & 31) == 0) vword = __ldg(vbits + (t >> 5));
        const bool is_valid = (vword >> (t & 31)) & 1u;
        double x0 = 0.0, x1 = 0.0;   // fetched early: consumed only after both products
        if (is_valid) {
            if (qv0) x0 = __ldg(xg + t * D + xc0);
            if (qv1) x1 = __ldg(xg + t * D + xc1);
        }
        const double* Bs = Bsm + s * MATB;
        const double* Gsrc = (t == 0) ? mp.C0m + static_cast<size_t>(MATG) * s : mp.Sigm + static_cast<size_t>(MATG) * s;
        Gsrc += g * NPm + 2 * c4;

        if (t > 0) {
            // ---------------- P1: T = B_s [C | M], in place, one block of CB tile columns at a time
#pragma unroll
            for (int pass = 0; pass < NPASS; ++pass) {
                constexpr int dummy = 0; (void)dummy;
                const int c_lo = pass * CB;
                const int nc = (GTC - c_lo) < CB ? (GTC - c_lo) : CB;
#pragma unroll
                for (int ti = 0; ti < GT; ++ti)
#pragma unroll
                    for (int tc = 0; tc < CB; ++tc)
                        if (tc < nc) acc[ti * CB + tc][0] = acc[ti * CB + tc][1] = 0.0;
                {
                    const double* Ap = Bs + g * LDB + c4;
                    const double* BpE = Cb + c4 * LDC + (g & 3) + ((g >> 2) ^ (sc & 1)) * 4 + (sc >> 1) * 8;
                    const double* BpO = BpE - (sc >> 1) * 16;
#pragma unroll 1
                    for (int k0 = 0; k0 < NK; k0 += 4) {
                        double a[GT], b[CB];
                        const int ao = ((k0 >> 2) ^ sg) << 2;
#pragma unroll
                        for (int ti = 0; ti < GT; ++ti) a[ti] = Ap[8 * ti * LDB + ao];
#pragma unroll
                        for (int tc = 0; tc < CB; ++tc)
                            if (tc < nc) b[tc] = ((c_lo + tc) & 1 ? BpO : BpE)[8 * (c_lo + tc)];
                        BpE += 4 * LDC;
                        BpO += 4 * LDC;
#pragma unroll
                        for (int ti = 0; ti < GT; ++ti)
#pragma unroll
                            for (int tc = 0; tc < CB; ++tc)
                                if (tc < nc) dmma884(acc[ti * CB + tc], a[ti], b[tc]);
                    }
                }
                __syncwarp();   // everybody finished reading C[:, J] (and M when J holds the mean columns)
#pragma unroll
                for (int ti = 0; ti < GT; ++ti)
#pragma unroll
                    for (int tc = 0; tc < CB; ++tc)
                        if (tc < nc)
                            *reinterpret_cast<double2*>(DPAIR(ti, c_lo + tc)) = make_double2(acc[ti * CB + tc][0], acc[ti * CB + tc][1]);
            }
        }
        // accumulators of the upper tiles start at Sig (t > 0) / hold the steady state C0 (t = 0); the global
        // loads are issued before the barrier so that their latency overlaps it
#pragma unroll
        for (int ti = 0; ti < GT; ++ti)
#pragma unroll
            for (int tjj = ti; tjj < GT; ++tjj) {
                const double2 v = __ldg(reinterpret_cast<const double2*>(Gsrc + 8 * ti * NPm + 8 * tjj));
                acc[UIDX(ti, tjj)][0] = v.x;
                acc[UIDX(ti, tjj)][1] = v.y;
            }
        if (t > 0) {
            __syncwarp();   // B: T (with M' in its mean columns) complete
            // ---------------- P2: C' = T B_s + Sig, upper tiles only
            const double* Ap = Cb + g * LDC + c4;
            const double* BpE = Bs + c4 * LDB + (g & 3) + ((g >> 2) ^ (sc & 1)) * 4 + (sc >> 1) * 8;
            const double* BpO = BpE - (sc >> 1) * 16;
#pragma unroll 1
            for (int k0 = 0; k0 < NK; k0 += 4) {
                double a[GT], b[GT];
                const int ao = ((k0 >> 2) ^ sg) << 2;
#pragma unroll
                for (int ti = 0; ti < GT; ++ti) a[ti] = Ap[8 * ti * LDC + ao];
#pragma unroll
                for (int tjj = 0; tjj < GT; ++tjj) b[tjj] = (tjj & 1 ? BpO : BpE)[8 * tjj];
                BpE += 4 * LDB;
                BpO += 4 * LDB;
#pragma unroll
                for (int ti = 0; ti < GT; ++ti)
#pragma unroll
                    for (int tjj = ti; tjj < GT; ++tjj) dmma884(acc[UIDX(ti, tjj)], a[ti], b[tjj]);
            }
        }

        // prior mean pair owned by this lane in tile row ti (rows 8*ti + g, mean columns q0, q1):
        // M0 at t = 0, afterwards M' sits in the buffer's mean columns (it was written with T)
        auto mean_prior = [&](int ti, double& m0, double& m1) {
            const int row = 8 * ti + g;
            if (t == 0) {
                m0 = (qv0 && row < N) ? __ldg(p.M0 + (s * N + row) * D + xc0) : 0.0;
                m1 = (qv1 && row < N) ? __ldg(p.M0 + (s * N + row) * D + xc1) : 0.0;
            } else {
                const double2 v = *reinterpret_cast<const double2*>(DPAIR(ti, TJM));
                m0 = qv0 ? v.x : 0.0;
                m1 = qv1 ? v.y : 0.0;
                if (p.hasG && row < N) {
                    if (qv0) m0 += __ldg(p.Gm + (s * N + row) * D + xc0);
                    if (qv1) m1 += __ldg(p.Gm + (s * N + row) * D + xc1);
                }
            }
        };

        if (is_valid) {
            // publish the two columns of C' that w touches.  Column j, tile column tjz = j >> 3: rows of tile
            // rows ti <= tjz come from the upper tile (ti, tjz) (lanes c4 == (j&7)>>1, element j&1); rows of
            // tile rows ti > tjz come, by symmetry, from row j of the upper tile (tjz, ti) (lanes g == j&7).
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int jz = z ? j1 : j0;
                const int tjz = jz >> 3, cj = jz & 7;
#pragma unroll
                for (int tjj = 0; tjj < GT; ++tjj)
                    if (tjj == tjz) {
                        if (c4 == (cj >> 1)) {
#pragma unroll
                            for (int ti = 0; ti <= tjj; ++ti)
                                colb[z * NPm + 8 * ti + g] = (cj & 1) ? acc[UIDX(ti, tjj)][1] : acc[UIDX(ti, tjj)][0];
                        }
                        if (g == cj) {
#pragma unroll
                            for (int ti = tjj + 1; ti < GT; ++ti)
                                *reinterpret_cast<double2*>(colb + z * NPm + 8 * ti + 2 * c4) =
                                    make_double2(acc[UIDX(tjj, ti)][0], acc[UIDX(tjj, ti)][1]);
                        }
                    }
            }
            if (t == 0) {   // at t = 0 the mean is not in the buffer yet: put it where w . M' reads it
#pragma unroll
                for (int ti = 0; ti < GT; ++ti) {
                    double m0, m1;
                    mean_prior(ti, m0, m1);
                    if (qv0) Cb[(8 * ti + g) * LDC + pcol(mp.MC0 + q0, sg)] = m0;
                    if (qv1) Cb[(8 * ti + g) * LDC + pcol(mp.MC0 + q1, sg)] = m1;
                }
            }
        }
        __syncwarp();   // C: T no longer needed; published columns (and M') visible
        double kr[GT];
        double xm0 = 0.0, xm1 = 0.0;
        if (is_valid) {
            // S = s2 + w^T C' w, from the 2x2 block of C' at the non-zeros (pyx:55-63)
            const double cw_j0 = fma(w1, colb[NPm + j0], w0 * colb[j0]);   // (C' w)[j0]
            const double cw_j1 = fma(w1, colb[NPm + j1], w0 * colb[j1]);   // (C' w)[j1]
            const double S = fma(w1, cw_j1, fma(w0, cw_j0, s2));
            const double Sinv = __drcp_rn(S);                               // 1/S, correctly rounded (pyx:63)
#pragma unroll
            for (int ti = 0; ti < GT; ++ti)
                kr[ti] = fma(w1, colb[NPm + 8 * ti + g], w0 * colb[8 * ti + g]) * Sinv;   // K = C' w / S (pyx:66-67)
#pragma unroll
            for (int tjj = 0; tjj < GT; ++tjj) {
                const double2 u = *reinterpret_cast<const double2*>(colb + 8 * tjj + 2 * c4);
                const double2 v = *reinterpret_cast<const double2*>(colb + NPm + 8 * tjj + 2 * c4);
                const double c0v = fma(w1, v.x, w0 * u.x), c1v = fma(w1, v.y, w0 * u.y);   // (C' w)[column pair]
#pragma unroll
                for (int ti = 0; ti <= tjj; ++ti) {
                    acc[UIDX(ti, tjj)][0] = fma(-kr[ti], c0v, acc[UIDX(ti, tjj)][0]);   // pyx:71-75
                    acc[UIDX(ti, tjj)][1] = fma(-kr[ti], c1v, acc[UIDX(ti, tjj)][1]);
                }
            }
            // innovation (pyx:79): x - w . M'  (G, if any, is not in the buffer: add w . G)
            const int sw0 = swz(j0), sw1 = swz(j1);
            if (qv0) {
                double ma = Cb[j0 * LDC + pcol(mp.MC0 + q0, sw0)], mb = Cb[j1 * LDC + pcol(mp.MC0 + q0, sw1)];
                if (p.hasG && t > 0) { ma += __ldg(p.Gm + (s * N + j0) * D + xc0); mb += __ldg(p.Gm + (s * N + j1) * D + xc0); }
                xm0 = x0 - fma(w1, mb, w0 * ma);
                if (g == 0) quad = fma(xm0 * xm0, Sinv, quad);
            }
            if (qv1) {
                double ma = Cb[j0 * LDC + pcol(mp.MC0 + q1, sw0)], mb = Cb[j1 * LDC + pcol(mp.MC0 + q1, sw1)];
                if (p.hasG && t > 0) { ma += __ldg(p.Gm + (s * N + j0) * D + xc1); mb += __ldg(p.Gm + (s * N + j1) * D + xc1); }
                xm1 = x1 - fma(w1, mb, w0 * ma);
                if (g == 0) quad = fma(xm1 * xm1, Sinv, quad);
            }
            // running product of Sinv with the exponent split off (no overflow over thousands of frames)
            if (lane == 0) {
                double lmant = lst[0] * Sinv;
                const int ex = ((__double2hiint(lmant) >> 20) & 0x7ff) - 1023;
                lmant = __hiloint2double(__double2hiint(lmant) - (ex << 20), __double2loint(lmant));
                lst[0] = lmant;
                reinterpret_cast<int*>(lst + 1)[0] += ex;
                reinterpret_cast<int*>(lst + 1)[1] += 1;
            }
            __syncwarp();   // everybody has read w . M' before M+ lands in the buffer
        }
        // ---------------- C+ (and M+ in its columns) becomes the operand of the next propagation:
        //                  upper tiles as accumulator pairs, strictly-upper tiles also mirrored
        if (t + 1 < T) {
#pragma unroll
            for (int ti = 0; ti < GT; ++ti) {
                double m0, m1;
                mean_prior(ti, m0, m1);
                if (is_valid) {
                    m0 = fma(kr[ti], xm0, m0);   // pyx:82-85
                    m1 = fma(kr[ti], xm1, m1);
                }
                if (!MX && ti > 0) {
                    // tile (ti, TJM) with ti <= TJM is upper and handled below; the mean pair of tile rows whose
                    // (ti, TJM) store happens in the loop below is merged there
                }
#pragma unroll
                for (int tjj = ti; tjj < GT; ++tjj) {
                    double v0 = acc[UIDX(ti, tjj)][0], v1 = acc[UIDX(ti, tjj)][1];
                    if (tjj > ti) {   // mirror: C[8 tjj + 2 c4 + e][8 ti + g] = C[8 ti + g][8 tjj + 2 c4 + e]
                        const int r0 = 8 * tjj + 2 * c4;
                        Cb[r0 * LDC + pcol(8 * ti + g, swz(r0))] = v0;
                        Cb[(r0 + 1) * LDC + pcol(8 * ti + g, swz(r0 + 1))] = v1;
                    }
                    if (!MX && tjj == TJM) {
                        if (qv0) v0 = m0;
                        if (qv1) v1 = m1;
                    }
                    *reinterpret_cast<double2*>(DPAIR(ti, tjj)) = make_double2(v0, v1);
                }
                if (MX) *reinterpret_cast<double2*>(DPAIR(ti, TJM)) = make_double2(m0, m1);
            }
        }
        __syncwarp();   // D
    }

    // logL = -1/2 [ sum xmm^2 Sinv - ncols * sum_t log Sinv_t + nvalid * ncols * log 2 pi ]   (pyx:88, 251-256)
    quad += __shfl_xor_sync(0xffffffffu, quad, 1);   // lanes 0..3 (g == 0) hold the per-dimension sums
    quad += __shfl_xor_sync(0xffffffffu, quad, 2);
    if (lane == 0) {
        const int lexp = reinterpret_cast<const int*>(lst + 1)[0], nvalid = reinterpret_cast<const int*>(lst + 1)[1];
        const double logdet = log(lst[0]) + lexp * 0.6931471805599453;
        p.out[static_cast<size_t>(e_sub) * p.P + pidx] = -0.5 * (quad - ncols * logdet + static_cast<double>(nvalid) * ncols * LOG_2PI);
    }
}

#undef DPAIR
#undef UIDX
#undef qv0
#undef qv1
#undef xc0
#undef xc1

// ================================================================================================
// k_mmac - FP64 tensor cores for polymers that do not fit one warp (8 < GT <= 14, N <= 112):
// ONE CTA PER FILTER, ONE WARP PER TILE COLUMN.
//
//   P1   warp(c):  T[:, c] = B_s C[:, c]   in place (it is the only reader of column c: __syncwarp only);
//                  GT DMMAs per k-step fed by GT fragments of B_s and one of C
//   ---- CTA barrier: T complete
//   P2   warp(c):  the c+1 upper tiles (ti <= c) of C'[:, c] = T B_s[:, c] + Sig
//   ---- publish the two columns of C' that w touches; CTA barrier
//   upd  rank-1 update of the warp's tiles; the owner of the last tile column also carries the mean
//   ---- C+ written back (upper pairs + mirrored), CTA barrier
// Symmetric output as in k_mma: (GT^2 + GT (GT+1)/2) DMMAs per k-step instead of 2 GT^2.
// The column -> warp map (mp.colmap) is chosen on the host so that the four warp schedulers of the SM
// carry equal numbers of DMMAs (P2 work grows with the column index).  The propagator of the current
// state is resident in shared memory (all states when they fit), re-staged by TMA at a state switch.
template <int GT, int NT, int LDA, int LDBB>
__device__ __forceinline__ void mmac_p2(double (&acc)[GT][2], const double* __restrict__ Ap, const double* __restrict__ Bp, int NK) {
#pragma unroll 1
    for (int k0 = 0; k0 < NK; k0 += 4) {
        double a[NT];
#pragma unroll
        for (int ti = 0; ti < NT; ++ti) a[ti] = Ap[8 * ti * LDA + k0];
        const double b = Bp[k0 * LDBB];
#pragma unroll
        for (int ti = 0; ti < NT; ++ti) dmma884(acc[ti], a[ti], b);
    }
}

struct CParams {
    MParams m;
    unsigned char colmap[16];   // warp -> tile column
    int b_all;                  // all S propagators resident
    int nhelp;                  // helper warps (GT % 4 == 1, no extra mean column): warps GT .. GT + nhelp - 1, see k_mmac
};

// P1 of one tile column for the tile rows [LO, HI): acc[ti] = sum_k B_s[ti][k] C[k][cc]
template <int GT, int LO, int HI, int LDA, int LDBB>
__device__ __forceinline__ void mmac_p1(double (&acc)[GT][2], const double* __restrict__ Ap, const double* __restrict__ Bp, int NK) {
#pragma unroll
    for (int ti = LO; ti < HI; ++ti) acc[ti][0] = acc[ti][1] = 0.0;
#pragma unroll 1
    for (int k0 = 0; k0 < NK; k0 += 4) {
        double a[HI - LO];
#pragma unroll
        for (int ti = LO; ti < HI; ++ti) a[ti - LO] = Ap[8 * ti * LDA + k0];
        const double b = Bp[k0 * LDBB];
#pragma unroll
        for (int ti = LO; ti < HI; ++ti) dmma884(acc[ti], a[ti - LO], b);
    }
}

template <int GT, bool MX>
// registers are allocated per scheduler (16 K each): ceil(GT / 4) warps share one
__global__ void __maxnreg__((16384 / (32 * ((GT + 3) / 4))) / 8 * 8 > 240 ? 240 : (16384 / (32 * ((GT + 3) / 4))) / 8 * 8) k_mmac(const __grid_constant__ CParams cp) {
    constexpr int GTC = GT + (MX ? 1 : 0);
    constexpr int TJM = MX ? GT : GT - 1;
    constexpr int NPm = 8 * GT, LDB = NPm + 4, LDC = 8 * GTC + 4;
    constexpr int MATB = NPm * LDB, MATG = NPm * NPm;
    const MParams& mp = cp.m;
    const KParams& p = mp.k;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    double* Bsm = reinterpret_cast<double*>(smem_raw + 16);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, c4 = lane & 3;
    const int e_sub = blockIdx.y;
    const int N = p.N, D = p.D, NK = mp.NK;
    // Warps 0 .. GT-1 own one tile column each.  With GT = 4k + 1 columns one scheduler (warps 0, 4, 8, ...) carries
    // k + 1 column warps and the others k, and P1 - the same GT tiles for every column, between two CTA barriers -
    // would run at the pace of that scheduler (13 columns: 52 vs 39 tiles, a fifth of the P1 phase idle).  Three helper
    // warps (GT, GT+1, GT+2 -> schedulers 1, 2, 3) therefore take the first HR = GT / 4 tile rows of P1 of the columns of
    // warps 0, 4, 8: 43 / 42 / 42 / 42 tiles per scheduler.  Helper and owner read the whole column before either
    // overwrites it (T is produced in place), so the pair meets at a named barrier between the product and the store.
    constexpr int HR = GT / 4;
    const bool helper = wid >= GT;                // helpers run their own compact frame loop (below) and own no column
    const bool helped = HR > 0 && !helper && (wid & 3) == 0 && (wid >> 2) < cp.nhelp;   // pair (wid >> 2) with warp GT + (wid >> 2)
    const int c = cp.colmap[helper ? 4 * (wid - GT) : wid];   // this warp's tile column (helper: the column it helps with)
    const bool mown = !helper && (c == GT - 1);   // ... which also carries the mean columns
    const bool mxw = MX && !helper && (c == 0);   // ... or computes the extra mean tile column of T

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

    double acc[GT][2];
    double* const myC = Cb + g * LDC + 2 * c4;   // accumulator pair of tile (ti, tj): myC + 8 ti LDC + 8 tj

    if (cp.b_all) mbar_wait(mbar, 0);

    if (HR > 0 && helper) {
        // ---------------- helper warp: tile rows [0, HR) of P1 of column c, and every CTA-wide synchronisation of
        //                  the main loop below (propagator re-staging wait, three barriers per frame)
        for (int t = 0; t < T; ++t) {
            while (t >= next_sw) {
                ++r_cur;
                s = p.run_states[static_cast<size_t>(pidx) * p.K1 + r_cur];
                next_sw = (r_cur + 1 < p.K1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + r_cur + 1] : 0x7fffffff;
            }
            if (!cp.b_all && t > 0 && s != s_loaded) {
                mbar_wait(mbar, bphase);
                bphase ^= 1;
                s_loaded = s;
            }
            if (t > 0) {
                const double* Bs = Bsm + (cp.b_all ? s * MATB : 0);
                mmac_p1<GT, 0, (HR > 0 ? HR : 1), LDB, LDC>(acc, Bs + g * LDB + c4, Cb + c4 * LDC + 8 * c + g, NK);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + wid - GT) : "memory");   // owner and helper have read column c
#pragma unroll
                for (int ti = 0; ti < HR; ++ti)
                    *reinterpret_cast<double2*>(myC + 8 * ti * LDC + 8 * c) = make_double2(acc[ti][0], acc[ti][1]);
                __syncthreads();   // T complete
            }
            __syncthreads();       // published columns visible
            __syncthreads();       // C+ complete
        }
        return;
    }

    for (int t = 0; t < T; ++t) {
        while (t >= next_sw) {
            ++r_cur;
            s = p.run_states[static_cast<size_t>(pidx) * p.K1 + r_cur];
            next_sw = (r_cur + 1 < p.K1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + r_cur + 1] : 0x7fffffff;
        }
        if ((t & 31) == 0) vword = __ldg(vbits + (t >> 5));
        const bool is_valid = (vword >> (t & 31)) & 1u;
        double x0 = 0.0, x1 = 0.0;
        if (is_valid && mown) {
            if (qv0) x0 = __ldg(xg + t * D + xc0);
            if (qv1) x1 = __ldg(xg + t * D + xc1);
        }
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

        if (t > 0) {
            // ---------------- P1: T[:, cc] = B_s Caug[:, cc], in place
            auto p1_column = [&](int cc) {
                const double* Ap = Bs + g * LDB + c4;
                const double* Bp = Cb + c4 * LDC + 8 * cc + g;
                mmac_p1<GT, 0, GT, LDB, LDC>(acc, Ap, Bp, NK);
                __syncwarp();   // this warp is the only reader of column cc
#pragma unroll
                for (int ti = 0; ti < GT; ++ti)
                    *reinterpret_cast<double2*>(myC + 8 * ti * LDC + 8 * cc) = make_double2(acc[ti][0], acc[ti][1]);
            };
            if (helped) {      // tile rows [HR, GT); the helper warp GT + (wid >> 2) does [0, HR)
                mmac_p1<GT, HR, GT, LDB, LDC>(acc, Bs + g * LDB + c4, Cb + c4 * LDC + 8 * c + g, NK);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + (wid >> 2)) : "memory");   // both warps have read column c
#pragma unroll
                for (int ti = HR; ti < GT; ++ti)
                    *reinterpret_cast<double2*>(myC + 8 * ti * LDC + 8 * c) = make_double2(acc[ti][0], acc[ti][1]);
            } else {
                p1_column(c);
            }
            if (mxw) p1_column(GT);
        }
        // upper tiles of this column start at Sig (t > 0) / hold C0 (t = 0)
#pragma unroll
        for (int ti = 0; ti < GT; ++ti)
            if (ti <= c) {
                const double2 v = __ldg(reinterpret_cast<const double2*>(Gsrc + 8 * ti * NPm + 8 * c));
                acc[ti][0] = v.x;
                acc[ti][1] = v.y;
            }
        if (t > 0) {
            __syncthreads();   // T complete
            // ---------------- P2: upper tiles of C'[:, c] = T B_s[:, c] + Sig.  The number of tiles (c + 1) is fixed
            // per warp; the loop is instantiated for every count so that the hot loop carries no predicates.
            const double* Ap = Cb + g * LDC + c4;
            const double* Bp = Bs + c4 * LDB + 8 * c + g;
            switch (c) {
#define P2_CASE(CC) case CC: if (CC < GT) mmac_p2<GT, (CC < GT ? CC + 1 : 1), LDC, LDB>(acc, Ap, Bp, NK); break;
                P2_CASE(0) P2_CASE(1) P2_CASE(2) P2_CASE(3) P2_CASE(4) P2_CASE(5) P2_CASE(6)
                P2_CASE(7) P2_CASE(8) P2_CASE(9) P2_CASE(10) P2_CASE(11) P2_CASE(12) P2_CASE(13)
#undef P2_CASE
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

        if (is_valid) {
            // publish column j of C' (j = j0, j1; tile column tjz = j >> 3): the warp of column tjz has its rows
            // 8 ti + g for ti <= tjz; warps of columns c > tjz hold, by symmetry, row j of tile (tjz, c)
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int jz = z ? j1 : j0;
                const int tjz = jz >> 3, cj = jz & 7;
                if (c == tjz) {
                    if (c4 == (cj >> 1)) {
#pragma unroll
                        for (int ti = 0; ti < GT; ++ti)
                            if (ti <= c) colb[z * NPm + 8 * ti + g] = (cj & 1) ? acc[ti][1] : acc[ti][0];
                    }
                } else if (c > tjz) {
                    if (g == cj) {
#pragma unroll
                        for (int ti = 0; ti < GT; ++ti)
                            if (ti == tjz) *reinterpret_cast<double2*>(colb + z * NPm + 8 * c + 2 * c4) = make_double2(acc[ti][0], acc[ti][1]);
                    }
                }
            }
            if (t == 0 && mown) {
#pragma unroll
                for (int ti = 0; ti < GT; ++ti) {
                    double m0, m1;
                    mean_prior(ti, m0, m1);
                    if (qv0) Cb[(8 * ti + g) * LDC + mp.MC0 + q0] = m0;
                    if (qv1) Cb[(8 * ti + g) * LDC + mp.MC0 + q1] = m1;
                }
            }
        }
        __syncthreads();   // T no longer needed; published columns (and M') visible
        double kr[GT];
        double xm0 = 0.0, xm1 = 0.0;
        if (is_valid) {
            const double cw_j0 = fma(w1, colb[NPm + j0], w0 * colb[j0]);
            const double cw_j1 = fma(w1, colb[NPm + j1], w0 * colb[j1]);
            const double S = fma(w1, cw_j1, fma(w0, cw_j0, s2));
            const double Sinv = __drcp_rn(S);                               // pyx:63
#pragma unroll
            for (int ti = 0; ti < GT; ++ti)
                if (ti <= c) kr[ti] = fma(w1, colb[NPm + 8 * ti + g], w0 * colb[8 * ti + g]) * Sinv;   // pyx:66-67
            {
                const double2 u = *reinterpret_cast<const double2*>(colb + 8 * c + 2 * c4);
                const double2 v = *reinterpret_cast<const double2*>(colb + NPm + 8 * c + 2 * c4);
                const double c0v = fma(w1, v.x, w0 * u.x), c1v = fma(w1, v.y, w0 * u.y);
#pragma unroll
                for (int ti = 0; ti < GT; ++ti)
                    if (ti <= c) {
                        acc[ti][0] = fma(-kr[ti], c0v, acc[ti][0]);   // pyx:71-75
                        acc[ti][1] = fma(-kr[ti], c1v, acc[ti][1]);
                    }
            }
            if (mown) {
                if (qv0) {
                    double ma = Cb[j0 * LDC + mp.MC0 + q0], mb = Cb[j1 * LDC + mp.MC0 + q0];
                    if (p.hasG && t > 0) { ma += __ldg(p.Gm + (s * N + j0) * D + xc0); mb += __ldg(p.Gm + (s * N + j1) * D + xc0); }
                    xm0 = x0 - fma(w1, mb, w0 * ma);                       // pyx:79
                    if (g == 0) quad = fma(xm0 * xm0, Sinv, quad);
                }
                if (qv1) {
                    double ma = Cb[j0 * LDC + mp.MC0 + q1], mb = Cb[j1 * LDC + mp.MC0 + q1];
                    if (p.hasG && t > 0) { ma += __ldg(p.Gm + (s * N + j0) * D + xc1); mb += __ldg(p.Gm + (s * N + j1) * D + xc1); }
                    xm1 = x1 - fma(w1, mb, w0 * ma);
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
                __syncwarp();   // the mean owner has read w . M' before M+ lands in the buffer
            }
        }
        // ---------------- C+ written back: upper pairs of this column, mirrored below the diagonal
        if (t + 1 < T) {
#pragma unroll
            for (int ti = 0; ti < GT; ++ti)
                if (ti <= c) {
                    double v0 = acc[ti][0], v1 = acc[ti][1];
                    if (ti < c) {   // mirror: C[8 c + 2 c4 + e][8 ti + g]
                        const int r0 = 8 * c + 2 * c4;
                        Cb[r0 * LDC + 8 * ti + g] = v0;
                        Cb[(r0 + 1) * LDC + 8 * ti + g] = v1;
                    }
                    if (mown) {
                        double m0, m1;
                        mean_prior(ti, m0, m1);
                        if (is_valid) {
                            m0 = fma(kr[ti], xm0, m0);   // pyx:82-85
                            m1 = fma(kr[ti], xm1, m1);
                        }
                        if (!MX) {
                            if (qv0) v0 = m0;
                            if (qv1) v1 = m1;
                        } else {
                            *reinterpret_cast<double2*>(myC + 8 * ti * LDC + 8 * TJM) = make_double2(m0, m1);
                        }
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

namespace bildk {

// ================================================================================================
// k_mmag - FP64 tensor cores for polymers whose covariance does not fit in shared memory
// (112 < N <= 256): ONE CTA PER FILTER, ONE WARP PER TILE COLUMN, covariance and intermediate in a per-CTA
// L2-resident workspace (two buffers, so no in-place constraint), fragments loaded with plain global
// loads (L1-cached; block barriers order them), tiles processed in row chunks of CH to stay within the
// 64-72 registers that 7-8 warps per scheduler leave.  Same symmetric scheme as k_mmac.
//   P1   warp(c): T[:, c] = B_s C[:, c]          (chunks of CH tile rows, stored to the T buffer)
//   ---- CTA barrier
//   P2   warp(c): prior C'[ti <= c, c] = T B_s[:, c] + Sig  (chunks; published columns; stored to the C buffer)
//   ---- CTA barrier
//   upd  read-modify-write of the warp's own tiles in the C buffer (rank-1 update), mirrored; mean by the
//        owner of the last tile column
//   ---- CTA barrier
struct GMParams {
    MParams m;
    unsigned char colmap[40];   // warp -> tile column
    double* work;               // [gridDim.x * gridDim.y][2 * NPm * LDC]
    const int* prof_traj;       // [P] trajectory of every profile (nullptr: single trajectory)
};

template <int CH, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_mmag(const __grid_constant__ GMParams gp, const int GT, const int MXi) {
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
    const int c = gp.colmap[wid];
    const bool mown = (c == GT - 1);
    const bool mxw = MX && (c == 0);

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
            // ---------------- P1: T[:, cc] = B_s Caug[:, cc]
            for (int pass = 0; pass < (mxw ? 2 : 1); ++pass) {
                const int cc = pass ? GT : c;
                for (int r0 = 0; r0 < GT; r0 += CH) {
                    double acc[CH][2];
#pragma unroll
                    for (int i = 0; i < CH; ++i) acc[i][0] = acc[i][1] = 0.0;
                    const double* Ap = Bs + static_cast<size_t>(8 * r0 + g) * LDB + c4;
                    const double* Bp = Cg + c4 * LDC + 8 * cc + g;
#pragma unroll 1
                    for (int k0 = 0; k0 < NK; k0 += 4) {
                        double a[CH];
#pragma unroll
                        for (int i = 0; i < CH; ++i) a[i] = (r0 + i < GT) ? Ap[8 * i * LDB + k0] : 0.0;
                        const double b = Bp[k0 * LDC];
#pragma unroll
                        for (int i = 0; i < CH; ++i) dmma884(acc[i], a[i], b);
                    }
#pragma unroll
                    for (int i = 0; i < CH; ++i)
                        if (r0 + i < GT)
                            *reinterpret_cast<double2*>(Tg + pairoff + 8 * (r0 + i) * LDC + 8 * cc) = make_double2(acc[i][0], acc[i][1]);
                }
            }
            __syncthreads();   // T complete
        }
        // ---------------- P2 (t > 0) / steady state (t = 0): prior C' tiles (ti <= c, c), published columns, stored to C
        for (int r0 = 0; r0 <= c; r0 += CH) {
            double acc[CH][2];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                acc[i][0] = acc[i][1] = 0.0;
                if (r0 + i <= c) {
                    const double2 v = __ldg(reinterpret_cast<const double2*>(Gsrc + 8 * (r0 + i) * NPm + 8 * c));
                    acc[i][0] = v.x;
                    acc[i][1] = v.y;
                }
            }
            if (t > 0) {
                const double* Ap = Tg + (8 * r0 + g) * LDC + c4;
                const double* Bp = Bs + c4 * LDB + 8 * c + g;
#pragma unroll 1
                for (int k0 = 0; k0 < NK; k0 += 4) {
                    double a[CH];
#pragma unroll
                    for (int i = 0; i < CH; ++i) a[i] = (r0 + i <= c) ? Ap[8 * i * LDC + k0] : 0.0;
                    const double b = Bp[k0 * LDB];
#pragma unroll
                    for (int i = 0; i < CH; ++i) dmma884(acc[i], a[i], b);
                }
            }
            if (is_valid) {
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    const int jz = z ? j1 : j0;
                    const int tjz = jz >> 3, cj = jz & 7;
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int ti = r0 + i;
                        if (ti <= c) {
                            if (c == tjz && c4 == (cj >> 1)) colb[z * NPm + 8 * ti + g] = (cj & 1) ? acc[i][1] : acc[i][0];
                            if (c > tjz && ti == tjz && g == cj)
                                *reinterpret_cast<double2*>(colb + z * NPm + 8 * c + 2 * c4) = make_double2(acc[i][0], acc[i][1]);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < CH; ++i)
                if (r0 + i <= c) *reinterpret_cast<double2*>(Cg + pairoff + 8 * (r0 + i) * LDC + 8 * c) = make_double2(acc[i][0], acc[i][1]);
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
                if (mown) {
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

namespace bildk {

// ================================================================================================
// k_mma2 - FP64 tensor cores, TWO WARPS PER FILTER (5 <= GT <= 7, i.e. 33 <= N <= 56).
//
// At these sizes shared memory holds only 4-6 filters per SM, so with one warp per filter (k_mma) every
// scheduler has a single warp and the FP64 pipe idles through that warp's update / write-back phases
// (ncu: DMMA pipe 70 % active, profiles/r01_ncu_n50_mma_v7_final.txt).  Here each filter is shared by a
// warp pair, so every scheduler holds two warps of different filters that cover each other's gaps:
//   role 0: P1 tile columns [0, CB)        P2 upper tiles of columns [0, CS)
//   role 1: P1 tile columns [CB, GTC)      P2 upper tiles of columns [CS, GT), the mean, the log-likelihood
// P1 is in place per column block, so the pair needs no barrier inside P1; three named barriers
// (bar.sync id, 64) per frame: T complete / published columns visible / posterior written back.
// CS balances the DMMA counts of the roles; roles alternate between filters so that each scheduler gets one
// warp of either role.
struct M2Params {
    MParams m;
    int FPC2;   // filters per CTA (2 warps each)
};

__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int GT, bool MX>
struct Mma2Cfg {
    static constexpr int GTC = GT + (MX ? 1 : 0);
    static constexpr int TJM = MX ? GT : GT - 1;
    static constexpr int NPm = 8 * GT, LDB = NPm + 4, LDC = 8 * GTC + 4;
    static constexpr int CB = (GT + 1) / 2;
    static constexpr int NU = GT * (GT + 1) / 2;
    static constexpr int pick_cs() {
        int best = 1, bestd = 1 << 30;
        for (int cs = 1; cs < GT; ++cs) {
            const int r0 = GT * CB + cs * (cs + 1) / 2, r1 = GT * (GTC - CB) + NU - cs * (cs + 1) / 2;
            const int d = r0 > r1 ? r0 - r1 : r1 - r0;
            if (d < bestd) { bestd = d; best = cs; }
        }
        return best;
    }
    static constexpr int CS = pick_cs();
};

template <int GT, bool MX, int ROLE>
__device__ __forceinline__ void mma2_run(const MParams& mp, double* Bsm, double* Cb, double* colb, double* lst,
                                         int pidx, int tj, int e_sub, int barid, int lane) {
    using C = Mma2Cfg<GT, MX>;
    constexpr int GTC = C::GTC, TJM = C::TJM, NPm = C::NPm, LDB = C::LDB, LDC = C::LDC, CB = C::CB, CS = C::CS;
    constexpr int MATB = NPm * LDB, MATG = NPm * NPm;
    constexpr int C1LO = ROLE == 0 ? 0 : CB, C1HI = ROLE == 0 ? CB : GTC, NC1 = C1HI - C1LO;       // P1 columns
    constexpr int C2LO = ROLE == 0 ? 0 : CS, C2HI = ROLE == 0 ? CS : GT;                            // P2 columns
    constexpr int RMAX = C2HI;                                                                       // tile rows touched in P2
    constexpr int NU2 = C2HI * (C2HI + 1) / 2 - C2LO * (C2LO + 1) / 2;
    constexpr int NACC = (GT * NC1 > NU2) ? GT * NC1 : NU2;
#define U2(ti, tjj) ((tjj) * ((tjj) + 1) / 2 - C2LO * (C2LO + 1) / 2 + (ti))
    const KParams& p = mp.k;
    const int g = lane >> 2, c4 = lane & 3;
    const int N = p.N, D = p.D, NK = mp.NK;
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
    if (ROLE == 1 && lane == 0) { lst[0] = 1.0; reinterpret_cast<int*>(lst + 1)[0] = 0; reinterpret_cast<int*>(lst + 1)[1] = 0; }

    int r_cur = 0;
    int s = p.run_states[static_cast<size_t>(pidx) * p.K1];
    int next_sw = (p.K1 > 1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + 1] : 0x7fffffff;

    double acc[NACC][2];
    double* const myC = Cb + g * LDC + 2 * c4;   // accumulator pair of tile (ti, tj): + 8 ti LDC + 8 tj

    for (int t = 0; t < T; ++t) {
        while (t >= next_sw) {
            ++r_cur;
            s = p.run_states[static_cast<size_t>(pidx) * p.K1 + r_cur];
            next_sw = (r_cur + 1 < p.K1) ? p.run_starts[static_cast<size_t>(pidx) * p.K1 + r_cur + 1] : 0x7fffffff;
        }
        if ((t & 31) == 0) vword = __ldg(vbits + (t >> 5));
        const bool is_valid = (vword >> (t & 31)) & 1u;
        double x0 = 0.0, x1 = 0.0;
        if (ROLE == 1 && is_valid) {
            if (qv0) x0 = __ldg(xg + t * D + xc0);
            if (qv1) x1 = __ldg(xg + t * D + xc1);
        }
        const double* Bs = Bsm + s * MATB;
        const double* Gsrc = ((t == 0) ? mp.C0m : mp.Sigm) + static_cast<size_t>(MATG) * s + g * NPm + 2 * c4;

        if (t > 0) {
            // ---------------- P1: T[:, J] = B_s Caug[:, J] for this role's column block, in place
#pragma unroll
            for (int i = 0; i < GT * NC1; ++i) acc[i][0] = acc[i][1] = 0.0;
            {
                const double* Ap = Bs + g * LDB + c4;
                const double* Bp = Cb + c4 * LDC + 8 * C1LO + g;
#pragma unroll 1
                for (int k0 = 0; k0 < NK; k0 += 4) {
                    double a[GT], b[NC1];
#pragma unroll
                    for (int ti = 0; ti < GT; ++ti) a[ti] = Ap[8 * ti * LDB + k0];
#pragma unroll
                    for (int tc = 0; tc < NC1; ++tc) b[tc] = Bp[k0 * LDC + 8 * tc];
#pragma unroll
                    for (int ti = 0; ti < GT; ++ti)
#pragma unroll
                        for (int tc = 0; tc < NC1; ++tc) dmma884(acc[ti * NC1 + tc], a[ti], b[tc]);
                }
            }
            __syncwarp();   // this warp is the only reader of its column block
#pragma unroll
            for (int ti = 0; ti < GT; ++ti)
#pragma unroll
                for (int tc = 0; tc < NC1; ++tc)
                    *reinterpret_cast<double2*>(myC + 8 * ti * LDC + 8 * (C1LO + tc)) = make_double2(acc[ti * NC1 + tc][0], acc[ti * NC1 + tc][1]);
        }
        // upper tiles of this role's P2 columns start at Sig (t > 0) / hold C0 (t = 0)
#pragma unroll
        for (int tjj = C2LO; tjj < C2HI; ++tjj)
#pragma unroll
            for (int ti = 0; ti <= tjj; ++ti) {
                const double2 v = __ldg(reinterpret_cast<const double2*>(Gsrc + 8 * ti * NPm + 8 * tjj));
                acc[U2(ti, tjj)][0] = v.x;
                acc[U2(ti, tjj)][1] = v.y;
            }
        if (t > 0) {
            pair_sync(barid);   // T complete (both column blocks)
            // ---------------- P2: C' = T B_s + Sig on this role's upper tiles
            const double* Ap = Cb + g * LDC + c4;
            const double* Bp = Bs + c4 * LDB + g;
#pragma unroll 1
            for (int k0 = 0; k0 < NK; k0 += 4) {
                double a[RMAX], b[C2HI - C2LO];
#pragma unroll
                for (int ti = 0; ti < RMAX; ++ti) a[ti] = Ap[8 * ti * LDC + k0];
#pragma unroll
                for (int tjj = C2LO; tjj < C2HI; ++tjj) b[tjj - C2LO] = Bp[k0 * LDB + 8 * tjj];
#pragma unroll
                for (int tjj = C2LO; tjj < C2HI; ++tjj)
#pragma unroll
                    for (int ti = 0; ti <= tjj; ++ti) dmma884(acc[U2(ti, tjj)], a[ti], b[tjj - C2LO]);
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

        if (is_valid) {
            // publish column j of C' (j = j0, j1): rows of tile rows ti <= tjz from the upper tile (ti, tjz) if this
            // role owns column tjz; rows of owned columns tjj > tjz from row j of the upper tile (tjz, tjj)
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int jz = z ? j1 : j0;
                const int tjz = jz >> 3, cj = jz & 7;
#pragma unroll
                for (int tjj = C2LO; tjj < C2HI; ++tjj) {
                    if (tjj == tjz) {
                        if (c4 == (cj >> 1)) {
#pragma unroll
                            for (int ti = 0; ti <= tjj; ++ti) colb[z * NPm + 8 * ti + g] = (cj & 1) ? acc[U2(ti, tjj)][1] : acc[U2(ti, tjj)][0];
                        }
                    } else if (tjj > tjz) {
                        if (g == cj) {
#pragma unroll
                            for (int ti = 0; ti < tjj; ++ti)
                                if (ti == tjz)
                                    *reinterpret_cast<double2*>(colb + z * NPm + 8 * tjj + 2 * c4) = make_double2(acc[U2(ti, tjj)][0], acc[U2(ti, tjj)][1]);
                        }
                    }
                }
            }
            if (ROLE == 1 && t == 0) {
#pragma unroll
                for (int ti = 0; ti < GT; ++ti) {
                    double m0, m1;
                    mean_prior(ti, m0, m1);
                    if (qv0) Cb[(8 * ti + g) * LDC + mp.MC0 + q0] = m0;
                    if (qv1) Cb[(8 * ti + g) * LDC + mp.MC0 + q1] = m1;
                }
            }
        }
        pair_sync(barid);   // T no longer needed by either warp; published columns (and M') visible
        double kr[RMAX];
        double xm0 = 0.0, xm1 = 0.0;
        if (is_valid) {
            const double cw_j0 = fma(w1, colb[NPm + j0], w0 * colb[j0]);
            const double cw_j1 = fma(w1, colb[NPm + j1], w0 * colb[j1]);
            const double S = fma(w1, cw_j1, fma(w0, cw_j0, s2));
            const double Sinv = __drcp_rn(S);                               // pyx:63
#pragma unroll
            for (int ti = 0; ti < RMAX; ++ti) kr[ti] = fma(w1, colb[NPm + 8 * ti + g], w0 * colb[8 * ti + g]) * Sinv;   // pyx:66-67
#pragma unroll
            for (int tjj = C2LO; tjj < C2HI; ++tjj) {
                const double2 u = *reinterpret_cast<const double2*>(colb + 8 * tjj + 2 * c4);
                const double2 v = *reinterpret_cast<const double2*>(colb + NPm + 8 * tjj + 2 * c4);
                const double c0v = fma(w1, v.x, w0 * u.x), c1v = fma(w1, v.y, w0 * u.y);
#pragma unroll
                for (int ti = 0; ti <= tjj; ++ti) {
                    acc[U2(ti, tjj)][0] = fma(-kr[ti], c0v, acc[U2(ti, tjj)][0]);   // pyx:71-75
                    acc[U2(ti, tjj)][1] = fma(-kr[ti], c1v, acc[U2(ti, tjj)][1]);
                }
            }
            if (ROLE == 1) {
                if (qv0) {
                    double ma = Cb[j0 * LDC + mp.MC0 + q0], mb = Cb[j1 * LDC + mp.MC0 + q0];
                    if (p.hasG && t > 0) { ma += __ldg(p.Gm + (s * N + j0) * D + xc0); mb += __ldg(p.Gm + (s * N + j1) * D + xc0); }
                    xm0 = x0 - fma(w1, mb, w0 * ma);                       // pyx:79
                    if (g == 0) quad = fma(xm0 * xm0, Sinv, quad);
                }
                if (qv1) {
                    double ma = Cb[j0 * LDC + mp.MC0 + q1], mb = Cb[j1 * LDC + mp.MC0 + q1];
                    if (p.hasG && t > 0) { ma += __ldg(p.Gm + (s * N + j0) * D + xc1); mb += __ldg(p.Gm + (s * N + j1) * D + xc1); }
                    xm1 = x1 - fma(w1, mb, w0 * ma);
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
                __syncwarp();   // this warp has read w . M' before M+ lands in the buffer
            }
        }
        // ---------------- C+ written back: this role's upper pairs, mirrored below the diagonal; role 1 merges the mean
        if (t + 1 < T) {
#pragma unroll
            for (int tjj = C2LO; tjj < C2HI; ++tjj)
#pragma unroll
                for (int ti = 0; ti <= tjj; ++ti) {
                    double v0 = acc[U2(ti, tjj)][0], v1 = acc[U2(ti, tjj)][1];
                    if (ti < tjj) {
                        const int r0 = 8 * tjj + 2 * c4;
                        Cb[r0 * LDC + 8 * ti + g] = v0;
                        Cb[(r0 + 1) * LDC + 8 * ti + g] = v1;
                    }
                    if (ROLE == 1 && tjj == GT - 1) {   // the last tile column sees every tile row: carry the mean here
                        double m0, m1;
                        mean_prior(ti, m0, m1);
                        if (is_valid) {
                            m0 = fma(kr[ti], xm0, m0);   // pyx:82-85
                            m1 = fma(kr[ti], xm1, m1);
                        }
                        if (!MX) {
                            if (qv0) v0 = m0;
                            if (qv1) v1 = m1;
                        } else {
                            *reinterpret_cast<double2*>(myC + 8 * ti * LDC + 8 * TJM) = make_double2(m0, m1);
                        }
                    }
                    *reinterpret_cast<double2*>(myC + 8 * ti * LDC + 8 * tjj) = make_double2(v0, v1);
                }
        }
        pair_sync(barid);   // C+ complete
    }

    if (ROLE == 1) {
        quad += __shfl_xor_sync(0xffffffffu, quad, 1);
        quad += __shfl_xor_sync(0xffffffffu, quad, 2);
        if (lane == 0) {
            const int lexp = reinterpret_cast<const int*>(lst + 1)[0], nvalid = reinterpret_cast<const int*>(lst + 1)[1];
            const double logdet = log(lst[0]) + lexp * 0.6931471805599453;
            p.out[static_cast<size_t>(e_sub) * p.P + pidx] = -0.5 * (quad - ncols * logdet + static_cast<double>(nvalid) * ncols * LOG_2PI);
        }
    }
#undef U2
}

template <int GT, bool MX>
__global__ void __launch_bounds__(384, 1) k_mma2(const __grid_constant__ M2Params mp2) {
    using C = Mma2Cfg<GT, MX>;
    constexpr int MATB = C::NPm * C::LDB;
    const MParams& mp = mp2.m;
    const KParams& p = mp.k;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    double* Bsm = reinterpret_cast<double*>(smem_raw + 16);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int fl = wid >> 1;                          // filter within the CTA
    const int role = (wid & 1) ^ ((fl >> 1) & 1);     // roles alternate so that every scheduler holds one warp of each
    const int tj = p.cta_traj ? p.cta_traj[blockIdx.x] : 0;
    const int first = p.cta_first ? p.cta_first[blockIdx.x] : blockIdx.x * mp2.FPC2;
    const int pend = p.traj_first[tj + 1];
    const int pidx = first + fl;
    const bool alive = (fl < mp2.FPC2) && (pidx < pend);

    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(mbar, static_cast<uint32_t>(MATB * p.S * sizeof(double)));
        for (int st = 0; st < p.S; ++st) {
            constexpr uint32_t CH = 32768;
            constexpr uint32_t bytes = MATB * sizeof(double);
            for (uint32_t off = 0; off < bytes; off += CH)
                tma_load_1d(reinterpret_cast<char*>(Bsm + st * MATB) + off, reinterpret_cast<const char*>(mp.Bm + static_cast<size_t>(MATB) * st) + off,
                            bytes - off < CH ? bytes - off : CH, mbar);
        }
    }
    if (!alive) return;   // both warps of a pair leave together; the named barriers below are per pair
    double* Cb = Bsm + MATB * p.S + fl * mp.fstride_m;
    double* colb = Cb + C::NPm * C::LDC;
    double* lst = colb + 2 * C::NPm;
    mbar_wait(mbar, 0);
    if (role == 0) mma2_run<GT, MX, 0>(mp, Bsm, Cb, colb, lst, pidx, tj, blockIdx.y, 1 + fl, lane);
    else mma2_run<GT, MX, 1>(mp, Bsm, Cb, colb, lst, pidx, tj, blockIdx.y, 1 + fl, lane);
}

}  // namespace bildk
