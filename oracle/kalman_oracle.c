/*
 * ORACLE - TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
 * (bild_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it, and only as the checker.
 *
 * Plain-C restatement of the reference's multi-state-Rouse Kalman-filter log-likelihood:
 *   /root/reference/bild/src/MSRouse_logL.pyx:19-90    Kalman_update
 *   /root/reference/bild/src/MSRouse_logL.pyx:143-256  MSRouse_logL (setup, frame loop, summation)
 * The BLAS-1/2 calls of the .pyx (dsymv/ddot/dscal/daxpy/dger/dcopy from scipy.linalg.cython_blas)
 * are written out as loops.  dsymv("u", ...) on a C-ordered array reads the LOWER triangle of the
 * row-major matrix (pyx:55, 210, 227, 235); symv_lower() below does the same.
 *
 * Pinned against: the reference .pyx itself, compiled from /root/reference into oracle/_ref/
 * (oracle/Makefile), and the reference's pure-Python twin (MSRouse_logL_py.py) imported from
 * /root/reference by oracle/make_golden.py -> tests/golden/ *.npz.  See tests/test_oracle.py.
 *
 * Build: gcc -O2 -fPIC -shared -o oracle/_build/libbild_oracle.so oracle/kalman_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* y <- alpha * sym(A) x + beta * y ; sym(A)[i][j] = A[max(i,j)][min(i,j)] (row-major lower triangle) */
static void symv_lower(int n, double alpha, const double *A, const double *x, int incx,
                       double beta, double *y, int incy)
{
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int j = 0; j <= i; ++j) acc += A[(size_t)i * n + j] * x[(size_t)j * incx];
        for (int j = i + 1; j < n; ++j) acc += A[(size_t)j * n + i] * x[(size_t)j * incx];
        y[(size_t)i * incy] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y[(size_t)i * incy]);
    }
}

/* pyx:19-90 */
static void kalman_update(int N, int D, int Dstar, const double *w, const double *x,
                          double *M /*N,D*/, double *C /*Dstar,N,N*/, const double *s2,
                          const uint32_t *Cind, double *logL /*D*/,
                          double *xmm, double *K, double *Sinv, double *Cw)
{
    static const double LOG_2PI = 1.8378770664093453; /* np.log(2*np.pi), pyx:14 */
    for (int d = 0; d < Dstar; ++d) {
        double *Cd = C + (size_t)d * N * N;
        double *Cwd = Cw + (size_t)d * N, *Kd = K + (size_t)d * N;
        symv_lower(N, 1.0, Cd, w, 1, 0.0, Cwd, 1);                 /* pyx:55-59  Cw = C w */
        double dot = 0.0;
        for (int i = 0; i < N; ++i) dot += Cwd[i] * w[i];
        Sinv[d] = 1.0 / (s2[d] + dot);                             /* pyx:63 */
        for (int i = 0; i < N; ++i) Kd[i] = Sinv[d] * Cwd[i];      /* pyx:66-67 */
        for (int i = 0; i < N; ++i)                                /* pyx:71-75  C -= K Cw^T */
            for (int j = 0; j < N; ++j) Cd[(size_t)i * N + j] -= Kd[i] * Cwd[j];
    }
    for (int d = 0; d < D; ++d) {
        double dot = 0.0;
        for (int i = 0; i < N; ++i) dot += w[i] * M[(size_t)i * D + d];
        xmm[d] = x[d] - dot;                                       /* pyx:79 */
        const double *Kd = K + (size_t)Cind[d] * N;
        for (int i = 0; i < N; ++i) M[(size_t)i * D + d] += xmm[d] * Kd[i]; /* pyx:82-85 */
        double si = Sinv[Cind[d]];
        logL[d] = -0.5 * (xmm[d] * xmm[d] * si - log(si) + LOG_2PI); /* pyx:88 */
    }
}

/*
 * One filter.  All arrays C-ordered float64.
 *   Bs (S,N,N) Gs (S,N,D) Sigs (S,N,N)      pyx:155-157
 *   M0 (S,N,D) C0 (S,N,N)                   steady state of every state model; profile[0] picks (pyx:160)
 *   w (N)  x (T,D; NaN = missing)  s2 (Dstar)  Cind (D)  states (T)
 * Returns 0 and writes *out; -1 on bad arguments.  A trajectory without any valid frame gives 0.0
 * (what MSRouse_logL_py.py:121 returns; the .pyx reads out of bounds there, pyx:186).
 */
int bild_oracle_logl(int N, int D, int Dstar, int S, int T,
                     const double *Bs, const double *Gs, const double *Sigs,
                     const double *M0, const double *C0, const double *w,
                     const double *x, const double *s2, const uint32_t *Cind,
                     const int32_t *states, double *out)
{
    if (N <= 0 || D <= 0 || Dstar <= 0 || S <= 0 || T <= 0) return -1;
    for (int t = 0; t < T; ++t) if (states[t] < 0 || states[t] >= S) return -1;
    size_t NN = (size_t)N * N, ND = (size_t)N * D;
    double *M = malloc(ND * 8), *Mp = malloc(ND * 8);
    double *C = malloc(Dstar * NN * 8), *Cp = malloc(NN * 8), *BC = malloc(N * 8);
    double *Cw = malloc((size_t)Dstar * N * 8), *K = malloc((size_t)Dstar * N * 8);
    double *Sinv = malloc(Dstar * 8), *xmm = malloc(D * 8), *ll = malloc(D * 8);
    memcpy(M, M0 + (size_t)states[0] * ND, ND * 8);                /* pyx:160-162 */
    for (int d = 0; d < Dstar; ++d) memcpy(C + d * NN, C0 + (size_t)states[0] * NN, NN * 8); /* pyx:163 */

    double total = 0.0;                                            /* pyx:251-254: sequential sum */
    for (int t = 0; t < T; ++t) {
        if (t > 0) {
            const double *B = Bs + (size_t)states[t] * NN;
            const double *G = Gs + (size_t)states[t] * ND;
            const double *Sg = Sigs + (size_t)states[t] * NN;
            for (int d = 0; d < D; ++d) {                          /* pyx:206-216  M <- B M + G */
                for (int i = 0; i < N; ++i) Mp[(size_t)i * D + d] = G[(size_t)i * D + d];
                symv_lower(N, 1.0, B, M + d, D, 1.0, Mp + d, D);
            }
            memcpy(M, Mp, ND * 8);
            for (int d = 0; d < Dstar; ++d) {                      /* pyx:220-241  C <- B C B + Sig */
                double *Cd = C + d * NN;
                for (int n0 = 0; n0 < N; ++n0) {
                    memcpy(Cp + (size_t)n0 * N, Sg + (size_t)n0 * N, N * 8);
                    symv_lower(N, 1.0, Cd, B + (size_t)n0 * N, 1, 0.0, BC, 1);
                    symv_lower(N, 1.0, B, BC, 1, 1.0, Cp + (size_t)n0 * N, 1);
                }
                memcpy(Cd, Cp, NN * 8);
            }
        }
        int valid = 1;                                             /* pyx:178 */
        for (int d = 0; d < D; ++d) if (isnan(x[(size_t)t * D + d])) valid = 0;
        if (valid) {                                               /* pyx:186-190, 244-248 */
            kalman_update(N, D, Dstar, w, x + (size_t)t * D, M, C, s2, Cind, ll, xmm, K, Sinv, Cw);
            for (int d = 0; d < D; ++d) total += ll[d];
        }
    }
    *out = total;
    free(M); free(Mp); free(C); free(Cp); free(BC); free(Cw); free(K); free(Sinv); free(xmm); free(ll);
    return 0;
}

/* Batch of P profiles given per frame: states (P,T).  Serial loop (amis.py:735-739 is a serial map). */
int bild_oracle_logl_batch(int N, int D, int Dstar, int S, int T, int P,
                           const double *Bs, const double *Gs, const double *Sigs,
                           const double *M0, const double *C0, const double *w,
                           const double *x, const double *s2, const uint32_t *Cind,
                           const int32_t *states, double *out)
{
    for (int p = 0; p < P; ++p) {
        int rc = bild_oracle_logl(N, D, Dstar, S, T, Bs, Gs, Sigs, M0, C0, w, x, s2, Cind,
                                  states + (size_t)p * T, out + p);
        if (rc) return rc;
    }
    return 0;
}
