/*
 * bild_b200 - C ABI of the B200-native BILD likelihood engine (libbild_b200.so).
 *
 * The reference has exactly one native plugin slot for this path: the CPython extension
 * ``bild.bin.MSRouse_logL`` exporting ``MSRouse_logL(model, profile, traj) -> float``
 * (/root/reference/bild/cython_imports.py:3-7, built by /root/reference/setup.py:41-46, source
 * /root/reference/bild/src/MSRouse_logL.pyx:95-256), called once per profile from the serial map in
 * ``FixedkSampler.logL`` (/root/reference/bild/amis.py:717-739).  This header is what a binding for
 * that slot would call instead (see INTEGRATION.md for the ctypes / Cython stub): the per-call
 * Python->C setup of the .pyx (pyx:143-199) is hoisted into a model handle and a trajectory handle,
 * and the per-profile call becomes one batched call.
 *
 * Conventions: plain C, no torch / CUDA types in any signature.  All matrices are C-ordered
 * (row-major) float64 exactly as the .pyx declares them (``FLOAT_t[:, :, ::1]``, pyx:114-122).
 * Every function returns 0 on success or a negative BILDK_E* code; bildk_last_error() gives the
 * message for the calling thread.  There is NO CPU fallback: without a CUDA device every compute
 * entry point fails with BILDK_ECUDA.
 *
 * Thread safety: handles may be shared between host threads.  The host-pointer entry points
 * (bildk_logl_runs / _st / _states / _runs_multi) serialise on a per-model lock from staging to copy-back;
 * bildk_amis_weights and bildk_marginal_posterior use per-thread scratch and a private stream.  The
 * *_device variants only enqueue work: calls on one model that may overlap in time must use ONE stream
 * (per-model scratch is reused in stream order).  Creating / destroying a handle must not race with calls
 * that use it.  Every entry point that launches work opens an NVTX range named after itself.
 */
#ifndef BILD_B200_H
#define BILD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BILDK_OK        0
#define BILDK_EINVAL   -1   /* bad argument (shape, state index out of range, non-finite error, ...) */
#define BILDK_ECUDA    -2   /* CUDA runtime error or no device */
#define BILDK_ENOMEM   -3
#define BILDK_EUNSUP   -4   /* configuration not supported by any kernel */

typedef struct bildk_model *bildk_model_t;
typedef struct bildk_traj  *bildk_traj_t;

/* ABI version (major*1000 + minor) and the last error message of the calling thread. */
int         bildk_version(void);
const char *bildk_last_error(void);

/* Number of CUDA devices visible (0 if none / no driver); never fails. */
int bildk_device_count(void);

/*
 * Model handle = what pyx:150-160 gathers on EVERY call, done once:
 *   w    (N)       model.measurement                       pyx:150
 *   B    (S,N,N)   [m._dynamics['B']   for m in models]    pyx:155   (symmetric; lower triangle is read,
 *   G    (S,N,d)   [m._dynamics['G']   ...]                pyx:156    as dsymv("u") on C-order does)
 *   Sig  (S,N,N)   [m._dynamics['Sig'] ...]                pyx:157
 *   M0   (S,N,d), C0 (S,N,N)   m.steady_state() of every state; profile[0] selects one   pyx:160
 * The arrays are copied to `device` (padded into the kernels' tile layout); the caller keeps
 * ownership of the host arrays.
 */
int bildk_model_create(int N, int d, int S,
                       const double *B, const double *G, const double *Sig,
                       const double *M0, const double *C0, const double *w,
                       int device, bildk_model_t *out);
int bildk_model_destroy(bildk_model_t model);

/*
 * Trajectory handle = pyx:144-147 and pyx:174-178, done once per trajectory:
 *   x     (T,d)   traj[:]; a frame is missing iff ANY component is NaN        pyx:178
 *   s2    (dstar) squared distinct localisation errors  (np.unique(noise)**2) pyx:145-146
 *   Cind  (d)     index into s2 for every spatial dimension                   pyx:147
 */
int bildk_traj_create(bildk_model_t model, int T, const double *x,
                      int dstar, const double *s2, const uint32_t *Cind,
                      bildk_traj_t *out);
int bildk_traj_destroy(bildk_traj_t traj);

/*
 * Batched replacement of the serial map amis.py:735-739 -> models.py:278 -> pyx:95.
 * Profiles are run-length coded the way FixedkSampler.st2profile builds them (amis.py:685-693):
 * profile p consists of K1 runs; run r has state run_states[p*K1+r] and covers frames
 * [run_starts[p*K1+r], run_starts[p*K1+r+1])  (the last run extends to T; run_starts[p*K1] must be 0;
 * empty runs are allowed and vanish, as the empty numpy slices do).  out[p] = logL.
 * HOST pointers; the call copies in, launches, copies out and returns when out[] is valid.
 */
int bildk_logl_runs(bildk_traj_t traj, int P, int K1,
                    const int32_t *run_starts, const uint8_t *run_states, double *out);

/*
 * Same batch, straight from the AMIS parametrisation (ss (P,K1) float64 interval lengths on the simplex,
 * thetas (P,K1) int64 states): the conversion FixedkSampler.st2profile does per profile (amis.py:685-693:
 * switches = floor(cumsum(s)[:-1] * (T-1)) + 1) is done here with the same IEEE operations in the same
 * order, so the discrete profiles are the very same ones.
 */
int bildk_logl_st(bildk_traj_t traj, int P, int K1, const double *ss, const int64_t *thetas, double *out);

/* Same, from per-frame state arrays states[p*T + t] (the Loopingprofile format, util.py:15-23). */
int bildk_logl_states(bildk_traj_t traj, int P, const int32_t *states, double *out);

/*
 * Device-resident variant: all three pointers are DEVICE pointers on the model's device, the launch
 * is asynchronous on `stream` (a cudaStream_t passed as void*, NULL = default stream).
 */
int bildk_logl_runs_device(bildk_traj_t traj, int P, int K1,
                           const int32_t *d_run_starts, const uint8_t *d_run_states,
                           double *d_out, void *stream);

/*
 * Many trajectories in one launch (dataset runs): profiles [offsets[i], offsets[i+1]) belong to
 * trajs[i]; all trajectories must share one model.  Host pointers, synchronous.
 */
int bildk_logl_runs_multi(int n_traj, const bildk_traj_t *trajs, const int32_t *offsets, int K1,
                          const int32_t *run_starts, const uint8_t *run_states, double *out);

/*
 * One fused AMIS step of a trajectory's batch (see bildk_amis_step below for the meaning of the fields): with `amis`
 * given, bildk_logl_runs_multi_submit enqueues, behind the filter kernel and in the same stream, the bookkeeping of every
 * trajectory whose entry has a non-NULL `ens` - the likelihoods go from the kernel's output into the ensemble on the
 * device, the statistics come back with the likelihoods at bildk_logl_wait.  All steps of a batch share two launches
 * (append; one cluster per ensemble).  An ensemble may appear once per batch; every array of a request must stay valid
 * until bildk_logl_wait.  The result of a step is a function of its ensemble alone (same bits as bildk_amis_step).
 */
typedef struct bildk_amis *bildk_amis_t;
typedef struct bildk_amis_req {
    bildk_amis_t ens;              /* NULL: no bookkeeping for this trajectory */
    const double *ss;              /* (n_i, K1 of the ensemble) interval lengths of the batch, n_i = offsets[i+1] - offsets[i] */
    const int64_t *thetas;         /* (n_i, K1) state traces */
    const double *A_cur;           /* (K1) */
    const double *logp_cur;        /* (S, K1) */
    double *head;                  /* out: 4 + 2 K1 + S K1 statistics */
    double *per_sample;            /* out: (n_total, 3), or NULL */
} bildk_amis_req;

/*
 * Asynchronous form of bildk_logl_runs_multi: `submit` stages the batch in pinned memory, enqueues upload, kernel and
 * download on the model's private stream WITHOUT waiting for anything, and returns a ticket; `out` must stay valid until
 * bildk_logl_wait(ticket) has returned (it fills out[]).  A model has two batch slots, i.e. at most two tickets in flight
 * (BILDK_EINVAL beyond that): the host code of one group of trajectories overlaps the kernel of the other
 * (bild_b200/dataset.py).  An empty batch returns a NULL ticket, which bildk_logl_wait accepts.
 */
int bildk_logl_runs_multi_submit(int n_traj, const bildk_traj_t *trajs, const int32_t *offsets, int K1,
                                 const int32_t *run_starts, const uint8_t *run_states, double *out,
                                 const bildk_amis_req *amis /* n_traj entries, or NULL */, void **ticket);
int bildk_logl_wait(void *ticket);
/* Non-blocking: 1 if the batch behind `ticket` has finished (bildk_logl_wait will not block), 0 if it is still running,
 * < 0 on error.  The two batches a model may have in flight run concurrently on the device when they share no scratch. */
int bildk_logl_ready(void *ticket);

/*
 * AMIS weight normalisation (amis.py:843-845, 878-900), deterministic fixed-order reduction:
 *   log_w[i] = logL[i] - logdelta[i] + log_nsteps
 *   stats[0] = max_i log_w        stats[1] = sum_i wo_i, wo_i = exp(log_w_i - max)
 *   stats[2] = sum_i (wo_i - mean(wo))^2   (the centred second moment scipy.stats.sem needs, amis.py:884)
 *   stats[3] = sum_i exp(log_w - max) * (logL[i] - cur_log_proposal[i])   (NaN terms skipped, as nansum)
 * Host pointers; log_w may be NULL.  n >= 1.
 */
int bildk_amis_weights(int n, const double *logL, const double *logdelta,
                       const double *cur_log_proposal, double log_nsteps,
                       double *log_w, double stats[4], int device);

/*
 * Marginal posterior of the state at every frame from a weighted ensemble of run-length profiles
 * (FixedkSampler.log_marginal_posterior, amis.py:942-972; SamplingResults.log_marginal_posterior, core.py:345-372):
 *   out[s][t] = log sum_{i : state_i(t) = s} exp(log_w[i])  -  log sum_i exp(log_w[i])
 * run_starts (n, K1) / run_states (n, K1) as for bildk_logl_runs (first run starts at 0, empty runs allowed,
 * padding runs start at T); log_w (n) the AMIS log-weights (or log-likelihoods for an exhaustive sample);
 * out (S, T) row-major.  Host pointers.  The reference materialises an (n, S, T) boolean tensor on the host for this.
 */
int bildk_marginal_posterior(int n, int K1, int T, int S, const int32_t *run_starts, const uint8_t *run_states,
                             const double *log_w, double *out, int device);

/*
 * AMIS proposal densities (amis.py:83-108 `Dirichlet.logpdf`, amis.py:258-281 `CFC.logpmf`, combined as
 * `FixedkSampler.log_proposal`, amis.py:697-715) of n samples under n_par proposals at once:
 *   out[j][i] = log Dirichlet(ss[i]; A[j]) + log CFC(thetas[i]; logp[j])
 * A (n_par, K1) concentrations; logp (n_par, S, K1) CFC log-weights; transitions (S, S), non-zero = allowed;
 * ss (n, K1) interval lengths; thetas (n, K1) state traces; out (n_par, n).  Host arrays, row-major.
 * Conventions of the reference: a sample with an s_i == 0 whose a_i < 1, entries outside [0, 1] or a sum off by
 * more than 1e-9 gets +inf (its importance weight vanishes, amis.py:98-108).  Host-side bookkeeping of the AMIS
 * loop - it replaces O(steps) scipy distribution constructions per step; it does
 * not touch the GPU and consumes no random numbers.
 */
int bildk_amis_log_proposal(int n_par, int n, int K1, int S, const double *A, const double *logp,
                            const uint8_t *transitions, const double *ss, const int64_t *thetas, double *out);

/*
 * ChoiceSampler arithmetic (reference bild/choicesampler.py:112-175), host side, no GPU, no random numbers: the selection
 * rule "first k within dE of the row maximum" on the Monte-Carlo draws scaled_rvs (samplesize, kmax) + mu (kmax).
 *   bildk_choice_pick: picks (samplesize) for one mean vector; NaN entries of mu mark omitted k (choicesampler.py:128-133)
 *   bildk_choice_dn  : dn (kmax, kmax), dn[k1][k2] = change of the count of k2 when muhat[k1] moves from -Dmu[k1]/2 to
 *                      +Dmu[k1]/2 (choicesampler.py:135-160) - 2 kmax evaluations of the rule in one pass over the draws.
 * Same double additions and comparisons as the numpy statement, hence the same integers (tests compare them).
 */
int bildk_choice_pick(int samplesize, int kmax, const double *scaled_rvs, const double *mu, double dE, int64_t *picks);
int bildk_choice_dn(int samplesize, int kmax, const double *scaled_rvs, const double *muhat, const double *Dmu, double dE,
                    int64_t *dn);

/*
 * Device-resident AMIS ensemble: the whole per-iteration bookkeeping of FixedkSampler.step (amis.py:824-845 mixture
 * denominators and weights, :878-900 evidence statistics, :137-151 Dirichlet moments, :300-303 CFC marginals) in ONE call.
 * The ensemble (interval lengths, state traces, likelihoods, mixture denominators, weights) and all past proposals stay
 * on the device; per step only the new batch goes up and the statistics of the refit come back:
 *   ss (n_new, K1), thetas (n_new, K1), logL (n_new)   the batch just evaluated (host arrays)
 *   A_cur (K1), logp_cur (S, K1)                        the proposal it was drawn from (it joins the mixture now)
 *   head[0..3]              max log_w | sum w | sum (w - mean w)^2 | nansum w (logL - log q_cur)    (as bildk_amis_weights)
 *   head[4 .. 4+K1)         weighted mean of the interval lengths       head[4+K1 .. 4+2K1)  their weighted variance
 *   head[4+2K1 + s*K1 + c]  log sum of the weights of the samples with theta[c] == s  (not normalised over s)
 *   per_sample (n_total, 3) log_w | logdelta | log q_cur of EVERY sample so far, or NULL
 * Limits: K1 <= 32, S <= 4 (BILDK_EUNSUP otherwise: the caller keeps the bookkeeping on the host).
 */
int bildk_amis_create(int K1, int S, const uint8_t *transitions, int device, bildk_amis_t *out);
int bildk_amis_destroy(bildk_amis_t h);
int bildk_amis_size(bildk_amis_t h, int *n_samples, int *n_proposals);
int bildk_amis_step(bildk_amis_t h, int n_new, const double *ss, const int64_t *thetas, const double *logL,
                    const double *A_cur, const double *logp_cur, double *head, double *per_sample);

/* Device-resident variant of bildk_amis_weights: device pointers (d_log_w may be NULL, d_stats has
 * room for 4 doubles), asynchronous on `stream`. */
int bildk_amis_weights_device(int n, const double *d_logL, const double *d_logdelta,
                              const double *d_cur_log_proposal, double log_nsteps,
                              double *d_log_w, double *d_stats, void *stream);

/* Diagnostics: kernels launched by this library since load; description of the kernel variant the
 * next bildk_logl_* call on this trajectory would use for a batch of P ("tile TS=5 G=4 warp ..."). */
long long   bildk_launch_count(void);
const char *bildk_describe_plan(bildk_traj_t traj, int P);

/* Diagnostics (host only, no GPU needed): the work-distribution tables two of the kernels are launched with, so that
 * their invariants can be tested on a CPU.
 *   kernel 0 (k_mmar, GT <= 4 tile columns, r = N - 8 (GT - 1) in 1..4 covariance rows in the last row block, ncols mean
 *     columns): out[0..7] = buffer row read by B-fragment lane g for the last tile column, out[8..11] = buffer row of
 *     mean column q (unused entries repeat the zero row).
 *   kernel 1 (k_mmact, GT in 8..14): for each of 16 warps w: out[18 w] = number of slots, out[18 w + 1] = slots in the
 *     first segment, out[18 w + 2 + 2 i] / out[18 w + 3 + 2 i] = tile row / tile column of slot i (i < 8).
 * Returns the number of bytes written (12 or 288), or a negative error code. */
int bildk_debug_tables(int kernel, int GT, int r, int ncols, unsigned char *out);

#ifdef __cplusplus
}
#endif
#endif /* BILD_B200_H */
