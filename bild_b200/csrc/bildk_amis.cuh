// Device-resident AMIS ensemble: the per-iteration bookkeeping of FixedkSampler.step on the GPU.
//
// The reference recomputes, on the host, every step (bild/amis.py):
//   :824-827  log-density of ALL old samples under the proposal that just joined the mixture, logaddexp into their
//             mixture denominators logdelta                                   (Dirichlet.logpdf :83-108, CFC.logpmf :258-281)
//   :836-839  log-density of the new samples under EVERY proposal so far, log-sum-exp -> their logdelta
//   :843-845  log_w = logL - logdelta + log(n_steps)        :878-900  evidence, its standard error, KL
//   :137-151  Dirichlet method of moments: weighted mean / variance of the interval lengths
//   :300-303  CFC method of marginals: per (state, slot) log-sum-exp of the weights
// i.e. O(ensemble) work per step through scipy objects, plus (round 1 of this engine) an upload of three ensemble-sized
// vectors per step for the weight kernel.  Here the ensemble (log s, theta, logL, logdelta, log q_cur, log w) stays in
// HBM; one launch per step does all of the above and returns only the statistics the host needs for the (tiny)
// proposal refit: 4 evidence sums, 2 K1 moments, S K1 marginals.
//
// One thread-block cluster of AMIS_CLUSTER CTAs per ensemble; CTA r owns a contiguous chunk of the samples in every pass, partial
// sums cross CTAs through distributed shared memory in rank order: the result is a function of the ensemble alone
// (bitwise reproducible, independent of what else runs), as the dataset driver's "identical to one-by-one runs"
// contract requires.
#pragma once
#include "bildk_kernels.cuh"

namespace bildk {

constexpr int AMIS_CLUSTER = 8;
constexpr int AMIS_THREADS = 512;
constexpr int AMIS_MAXK1 = 32;    // slots (k + 1) handled on the device
constexpr int AMIS_MAXS = 4;      // states handled on the device

struct AmisParams {
    int n_old, n_new, K1, S, n_par;
    // ensemble arrays, capacity >= n_old + n_new
    const double* logs;        // [n][K1] log of the interval lengths (-inf for s == 0)
    const uint8_t* thetas;     // [n][K1]
    const uint8_t* flags;      // [n] bit 0: sample rejected by scipy's dirichlet (outside [0,1] / sum off) -> +inf
    const double* ss;          // [n][K1] interval lengths (moments)
    const double* logL;        // [n]
    double* per;               // [n][3] per sample: log weight | mixture denominator logdelta (updated in place) | density
                               //        under the current proposal (one contiguous block for the copy back to the host)
    // proposals 0 .. n_par-1 (the last one is the current one)
    const double* A;           // [n_par][K1]
    const double* lognorm;     // [n_par] lgamma(sum a) - sum lgamma(a)
    const double* logp;        // [n_par][S][K1]
    const double* reach;       // [n_par][S][K1] log-sum-exp of logp[:, c] over the states reachable from m
    const double* norm0;       // [n_par]
    double log_nsteps;
    double* out;               // [4 + 2 K1 + S K1] stats | m | v | log marginals
};

// log q(s, theta) under proposal j (amis.py:697-715)
__device__ __forceinline__ double amis_log_proposal(const AmisParams& p, int j, const double* __restrict__ ls,
                                                    const uint8_t* __restrict__ th, bool bad0) {
    const int K1 = p.K1;
    const double* __restrict__ a = p.A + static_cast<size_t>(j) * K1;
    const double* __restrict__ lp = p.logp + static_cast<size_t>(j) * p.S * K1;
    const double* __restrict__ rc = p.reach + static_cast<size_t>(j) * p.S * K1;
    double dir = p.lognorm[j];
    bool bad = bad0;
    for (int c = 0; c < K1; ++c) {
        const double ac = a[c], l = ls[c];
        if (ac != 1.0) dir = fma(ac - 1.0, l, dir);          // xlogy(a - 1, s): zero when a == 1, even at s == 0
        if (l == -INFINITY && ac < 1.0) bad = true;          // scipy rejects s_i == 0 with a_i < 1 -> +inf (amis.py:98-108)
    }
    double cat = lp[th[0] * K1] - p.norm0[j];
    for (int c = 1; c < K1; ++c) cat += lp[th[c] * K1 + c] - rc[th[c - 1] * K1 + c];
    return (bad ? INFINITY : dir) + cat;
}

__device__ __forceinline__ double amis_logaddexp(double a, double b) {   // np.logaddexp
    if (a == b) return a + 0.6931471805599453;                            // also +-inf == +-inf
    const double hi = fmax(a, b), lo = fmin(a, b);
    if (hi == INFINITY || lo == -INFINITY) return hi;
    return hi + log1p(exp(lo - hi));
}

// One step of one ensemble as the kernels see it.  A launch handles MANY ensembles (blockIdx.y = job): the dataset driver
// enqueues the steps of all trajectories of a fused likelihood launch behind it with two launches in total.
struct AmisJob {
    AmisParams p;
    const double* stage;       // this step's inputs: ss_new (n_new K1) | A (K1) lognorm norm0 logp (S K1) reach (S K1) | thetas bytes (n_new K1)
    const double* logL_new;    // (n_new) likelihoods of the new samples (e.g. inside the filter kernel's output)
};

// New samples join the ensemble: log of the interval lengths (xlogy needs it for every proposal, every later step) and
// the sample-level rejection flag of scipy's dirichlet (entries outside [0, 1], sum off by more than 1e-9; amis.py:98-108);
// the proposal they were drawn from joins the list of proposals.
__global__ void k_amis_append(const AmisJob* __restrict__ jobs) {
    const AmisJob& job = jobs[blockIdx.y];
    const int n_new = job.p.n_new, K1 = job.p.K1, S = job.p.S;
    if (blockIdx.x * blockDim.x >= n_new) return;
    const size_t nk = static_cast<size_t>(n_new) * K1, n_old = job.p.n_old;
    const int sk = S * K1;
    const double* __restrict__ ss_new = job.stage;
    const double* __restrict__ prop_new = job.stage + nk;
    const uint8_t* __restrict__ th_new = reinterpret_cast<const uint8_t*>(prop_new + K1 + 2 + 2 * sk);
    if (blockIdx.x == 0) {   // the joining proposal, staged as A (K1) | lognorm | norm0 | logp (S K1) | reach (S K1)
        const size_t jp = job.p.n_par - 1;
        double* A = const_cast<double*>(job.p.A) + jp * K1;
        double* logp = const_cast<double*>(job.p.logp) + jp * sk;
        double* reach = const_cast<double*>(job.p.reach) + jp * sk;
        for (int e = threadIdx.x; e < K1 + 2 + 2 * sk; e += blockDim.x) {
            const double v = prop_new[e];
            if (e < K1) A[e] = v;
            else if (e == K1) const_cast<double*>(job.p.lognorm)[jp] = v;
            else if (e == K1 + 1) const_cast<double*>(job.p.norm0)[jp] = v;
            else if (e < K1 + 2 + sk) logp[e - K1 - 2] = v;
            else reach[e - K1 - 2 - sk] = v;
        }
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_new) return;
    double* ss = const_cast<double*>(job.p.ss) + n_old * K1;
    double* logs = const_cast<double*>(job.p.logs) + n_old * K1;
    uint8_t* thetas = const_cast<uint8_t*>(job.p.thetas) + n_old * K1;
    double sum = 0.0;
    bool bad = false;
    for (int c = 0; c < K1; ++c) {
        const double v = ss_new[static_cast<size_t>(i) * K1 + c];
        sum += v;
        if (!(v >= 0.0) || v > 1.0) bad = true;
        ss[static_cast<size_t>(i) * K1 + c] = v;
        logs[static_cast<size_t>(i) * K1 + c] = v > 0.0 ? log(v) : -INFINITY;
        thetas[static_cast<size_t>(i) * K1 + c] = th_new[static_cast<size_t>(i) * K1 + c];
    }
    if (!(fabs(sum - 1.0) <= 1e-9)) bad = true;
    const_cast<uint8_t*>(job.p.flags)[n_old + i] = bad ? 1 : 0;
    const_cast<double*>(job.p.logL)[n_old + i] = job.logL_new[i];
}

// CPAD: columns per row group (16 or 32, >= K1); a warp covers 32 / CPAD samples per pass-2/3 iteration.
template <int CPAD>
__global__ void __launch_bounds__(AMIS_THREADS) k_amis_step(const AmisJob* __restrict__ jobs) {
    namespace cg = cooperative_groups;
    __shared__ AmisParams p;                       // one cluster (gridDim.x CTAs) per job
    static_assert(sizeof(AmisParams) % 8 == 0, "copied by 64-bit words");
    if (threadIdx.x < sizeof(AmisParams) / 8)
        reinterpret_cast<unsigned long long*>(&p)[threadIdx.x] = reinterpret_cast<const unsigned long long*>(&jobs[blockIdx.y].p)[threadIdx.x];
    __syncthreads();
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = static_cast<int>(cluster.block_rank()), ncta = static_cast<int>(cluster.num_blocks());
    constexpr int NW = AMIS_THREADS / 32;
    constexpr int RPW = 32 / CPAD;                 // samples per warp and iteration in the column passes
    constexpr int NQ = 2 + AMIS_MAXS;              // per-column quantities of one pass (max)
    __shared__ double red[NW][4];
    __shared__ double colred[NW][CPAD][NQ];
    __shared__ double part[4];                     // this CTA's scalar partials: max, sum w, ssd, s3 (read by the peers)
    __shared__ double colpart[CPAD][NQ];           // this CTA's per-column partials of the current pass (read by the peers)
    __shared__ double bc[4];
    __shared__ double colbc[CPAD][NQ];             // cluster-wide per-column results of a pass (mmax / m)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = p.n_old + p.n_new, K1 = p.K1, S = p.S;
    const int chunk = (n + ncta - 1) / ncta;
    const int lo = crank * chunk, hi = min(n, lo + chunk);
    const int col = lane % CPAD, rsub = lane / CPAD;
    const bool colok = col < K1;

    // ---------------- pass 1 (thread per sample): densities, mixture denominators, log weights, their maximum
    double mx = -INFINITY;
    for (int i = lo + tid; i < hi; i += AMIS_THREADS) {
        const double* ls = p.logs + static_cast<size_t>(i) * K1;
        const uint8_t* th = p.thetas + static_cast<size_t>(i) * K1;
        const bool bad0 = p.flags[i] & 1;
        double ld, cl;
        if (i < p.n_old) {                         // amis.py:824-827
            cl = amis_log_proposal(p, p.n_par - 1, ls, th, bad0);
            ld = amis_logaddexp(p.per[3 * static_cast<size_t>(i) + 1], cl);
        } else {                                   // amis.py:836-839: streaming log-sum-exp over all proposals
            double m = -INFINITY, s = 0.0;
            cl = 0.0;
            for (int j = 0; j < p.n_par; ++j) {
                cl = amis_log_proposal(p, j, ls, th, bad0);
                if (cl > m) { s = (m == -INFINITY) ? 1.0 : fma(s, exp(m - cl), 1.0); m = cl; }
                else if (cl > -INFINITY) s += exp(cl - m);
            }
            ld = (m == INFINITY || m == -INFINITY) ? m : m + log(s);
        }
        p.per[3 * static_cast<size_t>(i) + 1] = ld;
        p.per[3 * static_cast<size_t>(i) + 2] = cl;
        const double lw = p.logL[i] - ld + p.log_nsteps;       // amis.py:843-845
        p.per[3 * static_cast<size_t>(i)] = lw;
        mx = fmax(mx, lw);
    }
    mx = warp_max(mx);
    if (lane == 0) red[wid][0] = mx;
    __syncthreads();                               // also: this CTA's logw / curlp are visible to all its threads
    if (wid == 0) {
        mx = warp_max((lane < NW) ? red[lane][0] : -INFINITY);
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

    // fixed-order reduction of per-thread column quantities q[0..nq): lanes sharing a column inside the warp (shuffle),
    // warps (shared memory, warp order), CTAs (distributed shared memory, rank order) -> colbc[col][k] in every CTA
    auto col_reduce = [&](double (&q)[NQ], int nq, bool is_max) {
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            if (k < nq) {
#pragma unroll
                for (int o = 16; o >= CPAD; o >>= 1) {
                    const double other = __shfl_xor_sync(0xffffffffu, q[k], o);
                    q[k] = is_max ? fmax(q[k], other) : q[k] + other;
                }
                if (rsub == 0) colred[wid][col][k] = q[k];
            }
        }
        __syncthreads();
        for (int e = tid; e < CPAD * nq; e += AMIS_THREADS) {
            const int c = e / nq, k = e % nq;
            double v = is_max ? -INFINITY : 0.0;
            for (int w = 0; w < NW; ++w) v = is_max ? fmax(v, colred[w][c][k]) : v + colred[w][c][k];
            colpart[c][k] = v;
        }
        cluster.sync();
        for (int e = tid; e < CPAD * nq; e += AMIS_THREADS) {
            const int c = e / nq, k = e % nq;
            double v = is_max ? -INFINITY : 0.0;
            for (int r = 0; r < ncta; ++r) {
                const double o = *cluster.map_shared_rank(&colpart[c][k], r);
                v = is_max ? fmax(v, o) : v + o;
            }
            colbc[c][k] = v;
        }
        cluster.sync();                            // peers are done reading colpart before the next pass overwrites it
    };

    // ---------------- pass 1b (thread per (sample, slot)): masked maxima of the weights per (state, slot)  (amis.py:300-303)
    double q[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) q[k] = -INFINITY;
    for (int i = lo + wid * RPW + rsub; i < hi; i += NW * RPW) {
        if (colok) {
            const double lw = p.per[3 * static_cast<size_t>(i)];
            const int st = p.thetas[static_cast<size_t>(i) * K1 + col];
#pragma unroll
            for (int k = 0; k < AMIS_MAXS; ++k) if (st == k) q[k] = fmax(q[k], lw);
        }
    }
    col_reduce(q, S, true);
    double mmax[AMIS_MAXS];
#pragma unroll
    for (int k = 0; k < AMIS_MAXS; ++k) mmax[k] = (k < S) ? colbc[col][k] : -INFINITY;
    __syncthreads();

    // ---------------- pass 2: sum of the shifted weights, unnormalised first moments, marginal sums
    double s1 = 0.0;
#pragma unroll
    for (int k = 0; k < NQ; ++k) q[k] = 0.0;
    for (int i = lo + tid; i < hi; i += AMIS_THREADS) s1 += exp(p.per[3 * static_cast<size_t>(i)] - mx);
    for (int i = lo + wid * RPW + rsub; i < hi; i += NW * RPW) {
        if (colok) {
            const double lw = p.per[3 * static_cast<size_t>(i)];
            const double wo = exp(lw - mx);
            q[0] = fma(wo, p.ss[static_cast<size_t>(i) * K1 + col], q[0]);
            const int st = p.thetas[static_cast<size_t>(i) * K1 + col];
#pragma unroll
            for (int k = 0; k < AMIS_MAXS; ++k)
                if (st == k && mmax[k] > -INFINITY && mmax[k] < INFINITY) q[1 + k] += exp(lw - mmax[k]);
        }
    }
    s1 = warp_sum(s1);
    if (lane == 0) red[wid][0] = s1;
    __syncthreads();
    if (wid == 0) {
        s1 = warp_sum((lane < NW) ? red[lane][0] : 0.0);
        if (lane == 0) part[1] = s1;
    }
    col_reduce(q, 1 + S, false);                   // its cluster.sync also publishes part[1]
    if (tid == 0) {
        double v = 0.0;
        for (int r = 0; r < ncta; ++r) v += *cluster.map_shared_rank(&part[1], r);
        bc[1] = v;
    }
    __syncthreads();
    s1 = bc[1];
    const double mean = s1 / n;
    const double mcol = colok ? colbc[col][0] / s1 : 0.0;      // weighted mean of slot `col` (amis.py:141-143)
    if (crank == 0 && tid < CPAD && tid < K1) {
        p.out[4 + tid] = colbc[tid][0] / s1;
        for (int k = 0; k < S; ++k) {                          // log marginal of (state k, slot tid), unnormalised over states
            const double m = (k == 0) ? mmax[0] : (k == 1) ? mmax[1] : (k == 2) ? mmax[2] : mmax[3];
            p.out[4 + 2 * K1 + k * K1 + tid] = (m > -INFINITY && m < INFINITY) ? log(colbc[tid][1 + k]) + m : m;
        }
    }
    __syncthreads();

    // ---------------- pass 3: centred second moment of the weights, KL numerator, weighted variances of the slots
    double ssd = 0.0, s3 = 0.0;
    for (int i = lo + tid; i < hi; i += AMIS_THREADS) {
        const double wo = exp(p.per[3 * static_cast<size_t>(i)] - mx);
        const double dv = wo - mean;
        ssd = fma(dv, dv, ssd);
        const double term = wo * (p.logL[i] - p.per[3 * static_cast<size_t>(i) + 2]);
        if (term == term) s3 += term;                          // nansum (amis.py:885-895)
    }
#pragma unroll
    for (int k = 0; k < NQ; ++k) q[k] = 0.0;
    for (int i = lo + wid * RPW + rsub; i < hi; i += NW * RPW) {
        if (colok) {
            const double wo = exp(p.per[3 * static_cast<size_t>(i)] - mx);
            const double dv = p.ss[static_cast<size_t>(i) * K1 + col] - mcol;
            q[0] = fma(wo, dv * dv, q[0]);
        }
    }
    ssd = warp_sum(ssd);
    s3 = warp_sum(s3);
    if (lane == 0) { red[wid][0] = ssd; red[wid][1] = s3; }
    __syncthreads();
    if (wid == 0) {
        ssd = warp_sum((lane < NW) ? red[lane][0] : 0.0);
        s3 = warp_sum((lane < NW) ? red[lane][1] : 0.0);
        if (lane == 0) { part[2] = ssd; part[3] = s3; }
    }
    col_reduce(q, 1, false);
    if (crank == 0) {
        if (tid == 0) {
            double a = 0.0, b = 0.0;
            for (int r = 0; r < ncta; ++r) { a += *cluster.map_shared_rank(&part[2], r); b += *cluster.map_shared_rank(&part[3], r); }
            p.out[0] = mx; p.out[1] = s1; p.out[2] = a; p.out[3] = b;
        }
        if (tid < CPAD && tid < K1) p.out[4 + K1 + tid] = colbc[tid][0] / s1;     // weighted variance (amis.py:144)
    }
    cluster.sync();   // no CTA may exit while rank 0 still reads its shared memory
}

}  // namespace bildk
