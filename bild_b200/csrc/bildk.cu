// Host side of libbild_b200.so: the C ABI declared in include/bild_b200.h.
//
// Replaces the per-call Python->C setup of /root/reference/bild/src/MSRouse_logL.pyx:143-199 by two
// handles (model, trajectory) that live on the GPU, and the serial per-profile map of
// /root/reference/bild/amis.py:735-739 by one batched kernel launch.
#include "../../include/bild_b200.h"
#include "bildk_kernels.cuh"
#include "bildk_mma.cuh"
#include "bildk_mmar.cuh"
#include "bildk_mmar2.cuh"
#include "bildk_mmarb.cuh"
#include "bildk_mmag2.cuh"
#include "bildk_mmact.cuh"
#include "bildk_amis.cuh"
#include "bildk_launch.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h>

using namespace bildk;

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(x)                                                                                        \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) return fail(BILDK_ECUDA, "%s: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// NVTX range around every ABI entry point that launches work (shows up in nsys / ncu --nvtx timelines)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: remember the largest value
// configured for every (device, kernel) pair (a process-wide `static` would skip the call on a second device).
cudaError_t ensure_dyn_smem(const void* kernel, size_t smem) {   // declared in bildk_launch.h
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> configured;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = configured[std::make_pair(dev, kernel)];
    if (smem > cur) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        cur = smem;
    }
    return cudaSuccess;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

template <typename T>
struct DevBuf {   // growable device buffer
    T* p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return BILDK_OK;
        size_t want = std::max(n, cap * 2);
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) return fail(BILDK_ENOMEM, "cudaMalloc(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e));
        cap = want;
        return BILDK_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// ------------------------------------------------------------------------------------------------
struct bildk_model {
    int N, D, S, device;
    // tile layout (fixed per model)
    int TS, G, BS, LD, NP;
    bool tile_ok;          // the tile kernel can hold this model on chip
    bool hasG;
    int nnz;
    int wz_idx[NZMAX];
    double wz_val[NZMAX];
    // device arrays
    double *dB = nullptr, *dSig = nullptr, *dC0 = nullptr;            // unpadded [S][N][N]
    double *dBpad = nullptr, *dSigpad = nullptr, *dC0pad = nullptr;   // padded   [S][NP][LD]
    double *dG = nullptr, *dM0 = nullptr, *dw = nullptr;
    uint16_t* d_lane_ab = nullptr;   // [G*G] lane -> tile map (2x2 tile blocks per lane quad)
    // tensor-core (DMMA) layout: 8x8 tiles, row strides == 4 (mod 8)
    bool mma_ok = false, mma_mx = false, mmac_ok = false, mmag_ok = false;
    int GT = 0, NPm = 0, LDBm = 0, LDCm = 0, MC0 = 0, NK = 0;
    double *dBm = nullptr, *dSigm = nullptr, *dC0m = nullptr;
    // register-chained tensor-core kernel (k_mmar): GT <= 4 and N mod 8 in 1..4
    bool mmar_ok = false;
    bool mmar2_ok = false;   // the same with two warps per filter (k_mmar2): GT 5..7
    bool mmar_mx = false;    // N mod 8 in {0, 5, 6, 7}: M^T in an extra row block of the filter buffer
    bool mmarb_ok = false;   // N mod 8 in {1, 2}, GT 2..4: border rows / columns in DFMAs (k_mmarb)
    int r_last = 0, LDr = 0, fstride_r = 0;
    double* dBr = nullptr;
    // per-model scratch of the launcher: partial logL of the d* sub-filters; covariance workspace of the N > 112 kernels
    DevBuf<double> part, work;
    int max_smem_optin = 0;
    int n_sm = 0;
    // the scratch buffers above are shared by every call on this model: host-pointer entry points hold this lock while
    // they stage and enqueue (ctypes releases the GIL, so two Python threads may call in concurrently)
    std::mutex mu;
    // host-pointer batches go through two SLOTS (double buffering): a pinned host block and its device mirror that hold one
    // batch's inputs, launch metadata and outputs, on a private stream.  Nothing in the submit path synchronises the
    // stream (all copies are pinned <-> device), so the host code of one group of trajectories overlaps the kernel of
    // another (bild_b200/dataset.py); bildk_logl_runs_multi = submit + wait.
    struct Slot {
        bildk_model* owner = nullptr;
        char* pin = nullptr;
        char* dev = nullptr;
        size_t cap = 0, cap_dev = 0;
        cudaEvent_t done = nullptr;
        struct PendingAmis { double* head; size_t off_head, n_head; double* per; size_t off_per, n_per; struct bildk_amis* ens; };
        std::vector<PendingAmis> amis;     // fused AMIS steps of the batch in flight: where their results go
        double* user_out = nullptr;
        size_t off_out = 0;
        int P = 0;
        bool busy = false;
        cudaStream_t st = nullptr;         // stream of the batch in flight (the model's, or this slot's own)
    } slots[2];
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr;            // second stream: batches in different slots overlap on the device when they share no scratch
};

struct bildk_traj {
    bildk_model* m;
    int T, dstar;
    double* dx = nullptr;        // (T,D)
    uint8_t* dvalid = nullptr;   // (T)
    double s2[DMAX];
    int ncols[DMAX];
    int cols[DMAX][DMAX];
    // single-trajectory launch metadata, resident
    const double** d_xptr = nullptr;
    const uint8_t** d_vptr = nullptr;
    int* d_T = nullptr;
    int* d_first = nullptr;      // [2] = {0, P}; P patched per call
    int n_valid;
    std::string plan;
};

// ------------------------------------------------------------------------------------------------
static const int TS_CAND[] = {5, 4, 6, 8};
static double ts_penalty(int TS) {   // relative cost per FMA, from tools/fp64_peak.cu on B200
    switch (TS) {
        case 4: return 1.12;
        case 5: return 1.00;
        case 6: return 0.97;
        default: return 1.08;
    }
}
static int maxt_for(int TS, bool ws) { return ws ? 128 : (TS <= 6 ? 512 : 256); }

// shared memory per filter, in doubles; == 8 (mod 16) so that the two filters sharing a warp (and the
// propagators of two states) sit 64 bytes apart modulo the 128-byte bank row
static size_t stride_8mod16(size_t n) { return (n + 15) / 16 * 16 + 8; }
static size_t filter_doubles(int NP, int LD, bool densew) {
    return stride_8mod16(static_cast<size_t>(NP) * LD + 2 * static_cast<size_t>(NP) * MSTRIDE + (densew ? 2 : NZMAX) * static_cast<size_t>(NP) + DMAX);
}

// Pick the register-tile edge for a model: least padded arithmetic among the variants that fit.
static void choose_tile(bildk_model* m) {
    const int forced = env_int("BILDK_TS", 0);
    double best = 1e300;
    m->tile_ok = false;
    for (int TS : TS_CAND) {
        if (forced && TS != forced) continue;
        const int G = (m->N + TS - 1) / TS;
        const int BS = TS + (TS & 1);
        const int LD = G * BS, NP = G * TS;
        if (G < m->D) continue;                       // mean columns ride on thread columns b < d
        const bool ws = G * G <= 32;
        if (G * G > maxt_for(TS, ws)) continue;
        const size_t matb = static_cast<size_t>(NP) * LD * 8;
        const size_t need = 16 + matb /*one B*/ + filter_doubles(NP, LD, m->nnz > NZMAX) * 8;
        if (need > static_cast<size_t>(m->max_smem_optin)) continue;
        int tpfs = G * G;
        if (ws) { tpfs = 1; while (tpfs < G * G) tpfs *= 2; }
        const double cost = static_cast<double>(NP) * NP * ts_penalty(TS) * tpfs / (G * G);
        if (cost < best) {
            best = cost;
            m->TS = TS; m->G = G; m->BS = BS; m->LD = LD; m->NP = NP;
            m->tile_ok = true;
        }
    }
}

static void pad_matrix(const double* src, int N, int TS, int BS, int LD, int NP, double* dst, bool symmetrize_lower) {
    std::fill(dst, dst + static_cast<size_t>(NP) * LD, 0.0);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            // dsymv("u") on C-ordered memory reads the lower triangle of the row-major array (pyx:55, 210, 227, 235)
            const double v = symmetrize_lower ? src[static_cast<size_t>(std::max(i, j)) * N + std::min(i, j)] : src[static_cast<size_t>(i) * N + j];
            dst[static_cast<size_t>(i) * LD + (j / TS) * BS + (j % TS)] = v;
        }
}

extern "C" int bildk_version(void) { return 1000; }
extern "C" const char* bildk_last_error(void) { return g_err.c_str(); }
extern "C" long long bildk_launch_count(void) { return g_launches.load(); }

extern "C" int bildk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int bildk_model_destroy(bildk_model_t m) {
    if (!m) return BILDK_OK;
    cudaSetDevice(m->device);
    for (double* p : {m->dB, m->dSig, m->dC0, m->dBpad, m->dSigpad, m->dC0pad, m->dG, m->dM0, m->dw, m->dBm, m->dSigm, m->dC0m, m->dBr})
        if (p) cudaFree(p);
    if (m->d_lane_ab) cudaFree(m->d_lane_ab);
    m->part.release(); m->work.release();
    for (auto& sl : m->slots) {
        if (sl.pin) cudaFreeHost(sl.pin);
        if (sl.dev) cudaFree(sl.dev);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    if (m->st) cudaStreamDestroy(m->st);
    if (m->st2) cudaStreamDestroy(m->st2);
    delete m;
    return BILDK_OK;
}

extern "C" int bildk_model_create(int N, int d, int S, const double* B, const double* G, const double* Sig,
                                  const double* M0, const double* C0, const double* w, int device,
                                  bildk_model_t* out) {
    if (!out) return fail(BILDK_EINVAL, "out is NULL");
    *out = nullptr;
    if (N < 1 || d < 1 || d > DMAX || S < 1 || S > 255)
        return fail(BILDK_EINVAL, "need N >= 1, 1 <= d <= %d, 1 <= S <= 255 (got N=%d d=%d S=%d)", DMAX, N, d, S);
    if (!B || !G || !Sig || !M0 || !C0 || !w) return fail(BILDK_EINVAL, "NULL array argument");
    const size_t NN = static_cast<size_t>(N) * N, ND = static_cast<size_t>(N) * d;
    for (size_t i = 0; i < S * NN; ++i)
        if (!std::isfinite(B[i]) || !std::isfinite(Sig[i]) || !std::isfinite(C0[i])) return fail(BILDK_EINVAL, "non-finite entry in B, Sig or C0");
    for (size_t i = 0; i < S * ND; ++i)
        if (!std::isfinite(G[i]) || !std::isfinite(M0[i])) return fail(BILDK_EINVAL, "non-finite entry in G or M0");
    int ndev = bildk_device_count();
    if (ndev == 0) return fail(BILDK_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(BILDK_EINVAL, "device %d out of range [0,%d)", device, ndev);
    CU(cudaSetDevice(device));

    bildk_model* m = new bildk_model();
    struct Guard {
        bildk_model* m;
        ~Guard() { if (m) bildk_model_destroy(m); }
    } guard{m};
    m->N = N; m->D = d; m->S = S; m->device = device;
    CU(cudaDeviceGetAttribute(&m->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CU(cudaDeviceGetAttribute(&m->n_sm, cudaDevAttrMultiProcessorCount, device));
    CU(cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&m->st2, cudaStreamNonBlocking));
    for (auto& sl : m->slots) {
        sl.owner = m;
        CU(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    m->hasG = false;
    for (size_t i = 0; i < S * ND; ++i) if (G[i] != 0.0) m->hasG = true;
    m->nnz = 0;
    for (int i = 0; i < N; ++i) {
        if (!std::isfinite(w[i])) return fail(BILDK_EINVAL, "non-finite measurement vector");
        if (w[i] != 0.0) {
            if (m->nnz < NZMAX) { m->wz_idx[m->nnz] = i; m->wz_val[m->nnz] = w[i]; }
            ++m->nnz;
        }
    }
    choose_tile(m);

    auto upload = [&](double** dst, const double* src, size_t n) -> int {
        CU(cudaMalloc(dst, n * sizeof(double)));
        CU(cudaMemcpy(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice));
        return BILDK_OK;
    };
    int rc = BILDK_OK;
    // unpadded copies (catch-all kernel); B symmetrised from its lower triangle like dsymv("u") reads it
    std::vector<double> Bs(S * NN);
    for (int s = 0; s < S; ++s)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) Bs[s * NN + static_cast<size_t>(i) * N + j] = B[s * NN + static_cast<size_t>(std::max(i, j)) * N + std::min(i, j)];
    if ((rc = upload(&m->dB, Bs.data(), S * NN)) || (rc = upload(&m->dSig, Sig, S * NN)) || (rc = upload(&m->dC0, C0, S * NN)) ||
        (rc = upload(&m->dG, G, S * ND)) || (rc = upload(&m->dM0, M0, S * ND)) || (rc = upload(&m->dw, w, N)))
        return rc;
    if (m->tile_ok) {
        const size_t matd = static_cast<size_t>(m->NP) * m->LD;
        std::vector<double> pad(S * matd);
        for (int s = 0; s < S; ++s) pad_matrix(B + s * NN, N, m->TS, m->BS, m->LD, m->NP, pad.data() + s * matd, true);
        if ((rc = upload(&m->dBpad, pad.data(), S * matd))) return rc;
        for (int s = 0; s < S; ++s) pad_matrix(Sig + s * NN, N, m->TS, m->BS, m->LD, m->NP, pad.data() + s * matd, false);
        if ((rc = upload(&m->dSigpad, pad.data(), S * matd))) return rc;
        for (int s = 0; s < S; ++s) pad_matrix(C0 + s * NN, N, m->TS, m->BS, m->LD, m->NP, pad.data() + s * matd, false);
        if ((rc = upload(&m->dC0pad, pad.data(), S * matd))) return rc;
        // lane -> (a, b): walk the tile grid in 2x2 blocks so that an aligned quad of lanes owns one block
        std::vector<uint16_t> lab;
        const int Gt = m->G;
        for (int ba = 0; ba < Gt; ba += 2)
            for (int bb = 0; bb < Gt; bb += 2)
                for (int ia = 0; ia < 2; ++ia)
                    for (int ib = 0; ib < 2; ++ib)
                        if (ba + ia < Gt && bb + ib < Gt) lab.push_back(static_cast<uint16_t>(((ba + ia) << 8) | (bb + ib)));
        CU(cudaMalloc(&m->d_lane_ab, lab.size() * sizeof(uint16_t)));
        CU(cudaMemcpy(m->d_lane_ab, lab.data(), lab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    // tensor-core layout (one warp per filter): N <= 56, sparse measurement vector
    {
        const int GT = (N + 7) / 8;
        if (GT <= 32 && m->nnz == 2) {
            m->GT = GT; m->NPm = 8 * GT;
            m->mma_mx = (m->NPm - N) < d;
            m->MC0 = m->mma_mx ? m->NPm : N;
            const bool swz = BILDK_MMA_SWZ && GT <= 4;   // must match k_mma's SWZ
            m->LDBm = swz ? (m->NPm + 15) / 16 * 16 : m->NPm + 4;
            m->LDCm = swz ? (8 * (GT + (m->mma_mx ? 1 : 0)) + 15) / 16 * 16 : 8 * (GT + (m->mma_mx ? 1 : 0)) + 4;
            m->NK = (N + 3) / 4 * 4;
            const size_t matb = static_cast<size_t>(m->NPm) * m->LDBm;
            const size_t matg = static_cast<size_t>(m->NPm) * m->NPm;
            // B: the shared-memory image (32-byte column groups XOR-swizzled by the row), TMA-copied verbatim
            {
                std::vector<double> pad(S * matb, 0.0);
                for (int s = 0; s < S; ++s)
                    for (int i = 0; i < N; ++i) {
                        const int sw = swz ? (((i & 1) << 1) | ((i >> 1) & 1)) : 0;
                        for (int j = 0; j < N; ++j)
                            pad[s * matb + static_cast<size_t>(i) * m->LDBm + ((((j >> 2) ^ sw) << 2) | (j & 3))] =
                                B[s * NN + static_cast<size_t>(std::max(i, j)) * N + std::min(i, j)];
                    }
                if ((rc = upload(&m->dBm, pad.data(), S * matb))) return rc;
            }
            auto fill = [&](const double* src, std::vector<double>& pad) {
                std::fill(pad.begin(), pad.end(), 0.0);
                for (int s = 0; s < S; ++s)
                    for (int i = 0; i < N; ++i)
                        for (int j = 0; j < N; ++j) pad[s * matg + static_cast<size_t>(i) * m->NPm + j] = src[s * NN + static_cast<size_t>(i) * N + j];
            };
            std::vector<double> padg(S * matg);
            fill(Sig, padg);
            if ((rc = upload(&m->dSigm, padg.data(), S * matg))) return rc;
            fill(C0, padg);
            if ((rc = upload(&m->dC0m, padg.data(), S * matg))) return rc;
            const size_t fbytes = (static_cast<size_t>(m->NPm) * m->LDCm + static_cast<size_t>(2) * m->NPm + 2) * 8;
            // one CTA per filter, one warp per tile column (k_mmac); at least one propagator resident
            if (GT >= 5 && GT <= 14) m->mmac_ok = 16 + matb * 8 + fbytes <= static_cast<size_t>(m->max_smem_optin);
            if (GT > 7) m->mmag_ok = true;   // covariance in an L2 workspace: any GT <= 32
            if (GT > 7) m->mma_ok = false;
            else
            m->mma_ok = 16 + matb * 8 * S + fbytes <= static_cast<size_t>(m->max_smem_optin);
            // register-chained kernels: M^T rides in the spare rows of the last tile-row block (rl <= 4: room for d <= 4 mean
            // rows and a zero row) or, for rl >= 5, in an extra row block of the filter buffer ("MX", bildk_mmar.cuh)
            const int rl = N - 8 * (GT - 1);
            if (GT <= 13 && rl >= 1 && rl <= 8 && d <= 4 && m->wz_idx[0] == 0 && m->wz_idx[1] == N - 1 && N >= 2) {   // end-to-end measurement only
                const int R = 8 * GT;
                const int LDr = (R % 16 == 8) ? R : R + 8;
                const size_t matr = static_cast<size_t>(R) * LDr;
                std::vector<double> pad(S * matr, 0.0);
                for (int s = 0; s < S; ++s)
                    for (int i = 0; i < N; ++i)
                        for (int j = 0; j < N; ++j)
                            pad[s * matr + static_cast<size_t>(i) * LDr + (j ^ (4 * ((i >> 1) & 1)))] =
                                B[s * NN + static_cast<size_t>(std::max(i, j)) * N + std::min(i, j)];
                if ((rc = upload(&m->dBr, pad.data(), S * matr))) return rc;
                m->r_last = rl; m->LDr = LDr;
                m->mmar_mx = rl > 4;
                m->fstride_r = static_cast<int>(matr) + (m->mmar_mx ? 8 * LDr : 0) + 2 * R + 8 + (GT >= 10 ? 8 * R : GT == 9 ? 5 * R : GT >= 8 ? 4 * R : GT >= 5 ? 2 * R : 0);   // k_mmar2 / k_mmar8: one C' w vector per warp
                // GT >= 10 (k_mmar8): one filter per CTA next to ONE resident propagator
                const bool fits = GT >= 10 ? 16 + matr * 8 + static_cast<size_t>(m->fstride_r) * 8 <= static_cast<size_t>(m->max_smem_optin)
                                           : 16 + matr * 8 * S + static_cast<size_t>(m->fstride_r) * 8 * (GT >= 8 ? 2 : 4) <= static_cast<size_t>(m->max_smem_optin);
                m->mmar_ok = fits && GT <= 4;
                m->mmar2_ok = fits && GT >= 5;
                m->mmarb_ok = m->mmar_ok && rl <= 2 && GT >= 2 && N + d <= 32;   // k_mmarb: + [2][R] doubles per filter for t = C b
            }
        }
    }
    guard.m = nullptr;
    *out = m;
    return BILDK_OK;
}

extern "C" int bildk_traj_destroy(bildk_traj_t t) {
    if (!t) return BILDK_OK;
    cudaSetDevice(t->m->device);
    if (t->dx) cudaFree(t->dx);
    if (t->dvalid) cudaFree(t->dvalid);
    if (t->d_xptr) cudaFree(t->d_xptr);
    if (t->d_vptr) cudaFree(t->d_vptr);
    if (t->d_T) cudaFree(t->d_T);
    if (t->d_first) cudaFree(t->d_first);
    delete t;
    return BILDK_OK;
}

extern "C" int bildk_traj_create(bildk_model_t m, int T, const double* x, int dstar, const double* s2,
                                 const uint32_t* Cind, bildk_traj_t* out) {
    if (!out) return fail(BILDK_EINVAL, "out is NULL");
    *out = nullptr;
    if (!m || !x || !s2 || !Cind) return fail(BILDK_EINVAL, "NULL argument");
    if (T < 1) return fail(BILDK_EINVAL, "trajectory needs at least one frame");
    if (dstar < 1 || dstar > m->D) return fail(BILDK_EINVAL, "dstar=%d must be in [1, d=%d]", dstar, m->D);
    for (int e = 0; e < dstar; ++e)
        if (!std::isfinite(s2[e]) || s2[e] < 0) return fail(BILDK_EINVAL, "localisation error must be finite and non-negative");
    for (int j = 0; j < m->D; ++j)
        if (Cind[j] >= static_cast<uint32_t>(dstar)) return fail(BILDK_EINVAL, "Cind[%d]=%u out of range", j, Cind[j]);
    CU(cudaSetDevice(m->device));
    bildk_traj* t = new bildk_traj();
    struct Guard {   // an early return through CU() must not leak the handle and its device allocations
        bildk_traj* t;
        ~Guard() { if (t) bildk_traj_destroy(t); }
    } guard{t};
    t->m = m; t->T = T; t->dstar = dstar;
    for (int e = 0; e < DMAX; ++e) { t->s2[e] = 0; t->ncols[e] = 0; for (int c = 0; c < DMAX; ++c) t->cols[e][c] = 0; }
    for (int e = 0; e < dstar; ++e) t->s2[e] = s2[e];
    for (int j = 0; j < m->D; ++j) { int e = Cind[j]; t->cols[e][t->ncols[e]++] = j; }
    std::vector<uint8_t> valid(T);
    t->n_valid = 0;
    for (int i = 0; i < T; ++i) {   // pyx:178: a frame is valid iff no component is NaN
        bool ok = true;
        for (int j = 0; j < m->D; ++j) if (std::isnan(x[static_cast<size_t>(i) * m->D + j])) ok = false;
        valid[i] = ok;
        t->n_valid += ok;
    }
    const size_t nx = static_cast<size_t>(T) * m->D;
    CU(cudaMalloc(&t->dx, nx * sizeof(double)));
    CU(cudaMemcpy(t->dx, x, nx * sizeof(double), cudaMemcpyHostToDevice));
    // layout: T byte flags | pad to 4 bytes | ceil(T/32) packed words (bit i of word w = frame 32 w + i)
    {
        const size_t boff = (static_cast<size_t>(T) + 3) / 4 * 4;
        const size_t nwords = (static_cast<size_t>(T) + 31) / 32;
        std::vector<uint8_t> buf(boff + 4 * nwords, 0);
        std::memcpy(buf.data(), valid.data(), T);
        for (int i = 0; i < T; ++i)
            if (valid[i]) buf[boff + 4 * (i / 32) + (i % 32) / 8] |= static_cast<uint8_t>(1u << (i % 8));   // little endian
        CU(cudaMalloc(&t->dvalid, buf.size()));
        CU(cudaMemcpy(t->dvalid, buf.data(), buf.size(), cudaMemcpyHostToDevice));
    }
    CU(cudaMalloc(&t->d_xptr, sizeof(double*)));
    CU(cudaMalloc(&t->d_vptr, sizeof(uint8_t*)));
    CU(cudaMalloc(&t->d_T, sizeof(int)));
    CU(cudaMalloc(&t->d_first, 2 * sizeof(int)));
    const double* xp = t->dx;
    const uint8_t* vp = t->dvalid;
    CU(cudaMemcpy(t->d_xptr, &xp, sizeof xp, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(t->d_vptr, &vp, sizeof vp, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(t->d_T, &T, sizeof T, cudaMemcpyHostToDevice));
    guard.t = nullptr;
    *out = t;
    return BILDK_OK;
}

// ------------------------------------------------------------------------------------------------
struct Plan {
    bool mma = false;      // tensor-core kernel, one warp per filter
    bool mmac = false;     // tensor-core kernel, one CTA per filter, one warp per tile column
    bool mmag = false;     // same, covariance in an L2 workspace (N > 112)
    int mmag_nc = 1;       // ... tile columns per warp (2, 4: k_mmag2)
    int nhelp = 0;         // mmac: helper warps that balance P1 across the schedulers
    bool mmact = false;    // ... and P2 / update / write-back dealt out as tile slots (k_mmact)
    bool mma2 = false;     // tensor-core kernel, two warps per filter (GT 5..7)
    bool mmar = false;     // tensor-core kernel, one warp per filter, T chained through registers (GT <= 4)
    bool mmarb = false;    // k_mmar with the N mod 8 in {1, 2} border rows / columns in DFMAs (k_mmarb; set together with mmar)
    bool mmar2 = false;    // the same with two warps per filter splitting the tile rows (GT 5..7)
    int maxf = 4;          // k_mmar2: filters per CTA the launched instantiation is compiled for
    bool mmar8 = false;    // k_mmar8 (GT = 13): one filter per CTA, eight warps (set together with mmar2)
    int nb = 0;            // ... its variant: resident 4-warp CTAs per SM it is compiled for
    unsigned char colmap[40] = {0};
    int WPC = 0;
    bool tile;
    bool ws, densew, b_all;
    int TS, FPC, TPFS, threads, maxt;
    size_t smem;
    int fstride, bstride;
};

template <int GT, bool MX>
static cudaError_t mma_launch(const MParams& mp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&k_mma<GT, MX>), smem);
        if (e != cudaSuccess) return e;
    }
    k_mma<GT, MX><<<grid, threads, smem, st>>>(mp);
    return cudaGetLastError();
}
template <int GT, bool MX>
static int mma_regs() {
    cudaFuncAttributes a{};
    if (cudaFuncGetAttributes(&a, k_mma<GT, MX>) != cudaSuccess) { cudaGetLastError(); return 255; }
    return a.numRegs;
}
#define MMA_DISPATCH(GTV, MXV, CALL)                                          \
    switch ((GTV) * 2 + ((MXV) ? 1 : 0)) {                                     \
        case 2: return CALL(1, false); case 3: return CALL(1, true);           \
        case 4: return CALL(2, false); case 5: return CALL(2, true);           \
        case 6: return CALL(3, false); case 7: return CALL(3, true);           \
        case 8: return CALL(4, false); case 9: return CALL(4, true);           \
        case 10: return CALL(5, false); case 11: return CALL(5, true);         \
        case 12: return CALL(6, false); case 13: return CALL(6, true);         \
        case 14: return CALL(7, false); case 15: return CALL(7, true);         \
    }
static int mma_regs_for(int GT, bool MX) {
#define CALL_REGS(G_, M_) mma_regs<G_, M_>()
    MMA_DISPATCH(GT, MX, CALL_REGS)
#undef CALL_REGS
    return 255;
}
static cudaError_t mma_launch_for(int GT, bool MX, const MParams& mp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
#define CALL_LAUNCH(G_, M_) mma_launch<G_, M_>(mp, grid, threads, smem, st)
    MMA_DISPATCH(GT, MX, CALL_LAUNCH)
#undef CALL_LAUNCH
    return cudaErrorInvalidValue;
}

// The launchers of the register-chained kernels (k_mmar / k_mmarb / k_mmar2 / k_mmar8) live in translation units of their own
// (bildk_tu_*.cu, declared in bildk_launch.h): NVVM spends four of the five minutes of a single-file build on their unrolled
// template instantiations, and the translation units compile in parallel (bild_b200/build.py).
constexpr int MMAR2_MAXF = 4;
static int mmar2_maxf(int GT, bool MX) {
    // GT = 5: 12 warps at 168 registers (spill-free without MX, 76 bytes with) beat 8 warps at 194 / 210 by 3-4 %
    // (profiles/r02_mx_variants.txt); GT = 6 spills at 168 registers and loses 18 %
    // GT = 8: four warps per filter, 2 filters = 8 warps at 255 registers (spill-free; 3 filters at 168 registers spill 400-600
    // bytes per thread and are 5-12 % slower, profiles/r02_gt8_variants.txt)
    const int dflt = GT == 5 ? 6 : (GT >= 8 ? 2 : MMAR2_MAXF);
    const int want = env_int("BILDK_MMAR2_MAXF", dflt);
    return mmar2_has(GT, want, MX) ? want : dflt;
}

// Rows of the last tile-row block that carry M^T (and one all-zero row), placed so that the B-fragment loads of the
// permuted last tile column are bank-conflict free: a 128-bit load is served per quarter warp (lanes g = 2i, 2i+1:
// the two rows must differ in parity), a 64-bit load per half warp (lanes g = 4h..4h+3: rows distinct mod 4).
static void mmar_tables(int GT, int r, int ncols, unsigned char* lastrow, unsigned char* mrow) {
    if (r > 4) {
        // MX: M^T row q = buffer row 8 GT + 2 q + 1, zero row = 8 GT.  The permuted (mean) tile column is only ever read with
        // 128-bit loads, served per quarter warp (lanes g = 2i, 2i+1): the zero row is even, the mean rows are odd, and the
        // row stride is == 8 (mod 16) doubles, so the two rows of a quarter warp fall into opposite halves of the banks
        const int Z = 8 * GT;
        for (int i = 0; i < 4; ++i) {
            lastrow[2 * i] = static_cast<unsigned char>(Z);
            lastrow[2 * i + 1] = static_cast<unsigned char>(i < ncols ? Z + 2 * i + 1 : Z);
            mrow[i] = static_cast<unsigned char>(i < ncols ? Z + 2 * i + 1 : Z);
        }
        return;
    }
    const int base = 8 * (GT - 1);
    std::vector<int> avail;
    for (int i = r; i < 8; ++i) avail.push_back(base + i);
    const bool need_zero = (r < 4) || (ncols < 4);
    const int want = ncols + (need_zero ? 1 : 0);
    std::vector<int> idx(avail.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = static_cast<int>(i);
    int best = 1 << 30;
    std::vector<int> pick(want, base + r);
    std::sort(idx.begin(), idx.end());
    do {   // permutations of the spare rows; the first `want` entries are (m_0 .. m_{ncols-1}, zero)
        int rows[8];
        const int Z = need_zero ? avail[idx[ncols]] : -1;
        for (int i = 0; i < 4; ++i) {
            rows[2 * i] = i < r ? base + i : Z;
            rows[2 * i + 1] = i < ncols ? avail[idx[i]] : Z;
        }
        int cost = 0;
        for (int i = 0; i < 4; ++i)
            if (rows[2 * i] != rows[2 * i + 1] && ((rows[2 * i] - rows[2 * i + 1]) & 1) == 0) cost += 4;
        for (int h = 0; h < 2; ++h)
            for (int a = 0; a < 4; ++a)
                for (int b = a + 1; b < 4; ++b)
                    if (rows[4 * h + a] != rows[4 * h + b] && ((rows[4 * h + a] - rows[4 * h + b]) & 3) == 0) cost += 1;
        if (cost < best) {
            best = cost;
            for (int i = 0; i < want; ++i) pick[i] = avail[idx[i]];
        }
        std::reverse(idx.begin() + std::min<size_t>(want, idx.size()), idx.end());   // skip permutations of the unused tail
    } while (std::next_permutation(idx.begin(), idx.end()));
    const int Z = need_zero ? pick[ncols] : base;
    for (int i = 0; i < 4; ++i) {
        lastrow[2 * i] = static_cast<unsigned char>(i < r ? base + i : Z);
        lastrow[2 * i + 1] = static_cast<unsigned char>(i < ncols ? pick[i] : Z);
        mrow[i] = static_cast<unsigned char>(i < ncols ? pick[i] : Z);
    }
}

template <int GT, bool MX>
static cudaError_t mmac_launch(const CParams& cp, dim3 grid, size_t smem, cudaStream_t st) {
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&k_mmac<GT, MX>), smem);
        if (e != cudaSuccess) return e;
    }
    k_mmac<GT, MX><<<grid, 32 * (GT + cp.nhelp), smem, st>>>(cp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFuncAttributes a{};
        cudaFuncGetAttributes(&a, k_mmac<GT, MX>);
        fprintf(stderr, "k_mmac<%d,%d> launch failed: regs=%d maxThreads=%d static_smem=%zu local=%zu dyn_smem=%zu threads=%d max_dyn=%d\n",
                GT, (int)MX, a.numRegs, a.maxThreadsPerBlock, a.sharedSizeBytes, a.localSizeBytes, smem, 32 * GT, a.maxDynamicSharedSizeBytes);
    }
    return e;
}
template <int GT, int NH, int NW>
static cudaError_t mmact_launch(const CTParams& ct, dim3 grid, size_t smem, cudaStream_t st) {
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&k_mmact<GT, NH, NW>), smem);
        if (e != cudaSuccess) return e;
    }
    k_mmact<GT, NH, NW><<<grid, 32 * NW, smem, st>>>(ct);
    return cudaGetLastError();
}
// warps of k_mmact per GT: GT column warps (+ 3 P1 helpers when GT = 4k + 1), topped up to a multiple of four so that
// the upper tiles of P2 can be dealt out evenly over the four schedulers
static int mmact_warps(int GT) { return GT <= 8 ? 8 : (GT <= 9 ? 12 : 16); }
static int mmact_helpers(int GT) { return (GT % 4 == 1 && GT >= 9) ? 3 : 0; }
static cudaError_t mmact_launch_for(int GT, const CTParams& ct, dim3 grid, size_t smem, cudaStream_t st) {
    switch (GT) {
        case 8: return mmact_launch<8, 0, 8>(ct, grid, smem, st);
        case 9: return mmact_launch<9, 3, 12>(ct, grid, smem, st);
        case 10: return mmact_launch<10, 0, 16>(ct, grid, smem, st);
        case 11: return mmact_launch<11, 0, 16>(ct, grid, smem, st);
        case 12: return mmact_launch<12, 0, 16>(ct, grid, smem, st);
        case 13: return mmact_launch<13, 3, 16>(ct, grid, smem, st);
        case 14: return mmact_launch<14, 0, 16>(ct, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}
// slot table of k_mmact: the upper tiles (ti <= c), last column first, rows ascending, dealt out to the GT + 3 warps in
// order as SEGMENTS (consecutive tile rows of one column), at most two per warp; the first (total mod warps) warps may
// take one slot more.  Pieces that would have been a third segment go to the lightest single-segment warps.
// Warp w runs on scheduler w % 4.
static bool mmact_slots(int GT, CTParams& ct) {
    const int nw = mmact_warps(GT), total = GT * (GT + 1) / 2, base = total / nw, rem = total % nw;
    struct Seg { int c, lo, n; };
    std::vector<std::vector<Seg>> warps(1);
    std::vector<Seg> extra;
    auto cap = [&](int w) { return base + (w < rem ? 1 : 0); };
    auto load = [](const std::vector<Seg>& v) { int x = 0; for (const Seg& sg : v) x += sg.n; return x; };
    for (int c = GT - 1; c >= 0; --c) {
        int ti = 0;
        while (ti <= c) {
            const int w = static_cast<int>(warps.size()) - 1;
            if (w >= nw) { extra.push_back({c, ti, c + 1 - ti}); break; }
            const int room = cap(w) - load(warps.back());
            if (room == 0 || warps.back().size() == 2) { warps.emplace_back(); continue; }
            const int n = std::min(room, c + 1 - ti);
            warps.back().push_back({c, ti, n});
            ti += n;
        }
    }
    while (static_cast<int>(warps.size()) > nw) {   // pieces of warps beyond the last one
        for (const Seg& sg : warps.back()) extra.push_back(sg);
        warps.pop_back();
    }
    warps.resize(nw);
    for (const Seg& sg : extra) {
        int best = -1;
        for (int w = 0; w < nw; ++w)
            if (warps[w].size() < 2 && load(warps[w]) + sg.n <= MMACT_MAXS && (best < 0 || load(warps[w]) < load(warps[best]))) best = w;
        if (best < 0) return false;
        warps[best].push_back(sg);
    }
    for (int w = 0; w < 16; ++w) { ct.nslot[w] = 0; ct.nsegA[w] = 0; }
    for (int w = 0; w < nw; ++w) {
        int k = 0;
        for (size_t sgi = 0; sgi < warps[w].size(); ++sgi) {
            const Seg& sg = warps[w][sgi];
            for (int i = 0; i < sg.n; ++i, ++k) {
                ct.slot_ti[w][k] = static_cast<unsigned char>(sg.lo + i);
                ct.slot_c[w][k] = static_cast<unsigned char>(sg.c);
            }
            if (sgi == 0) ct.nsegA[w] = static_cast<unsigned char>(sg.n);
        }
        ct.nslot[w] = static_cast<unsigned char>(k);
        if (k == 0 || k > MMACT_MAXS) return false;
    }
    return ct.slot_ti[0][0] == 0 && ct.slot_c[0][0] == GT - 1;
}

static cudaError_t mmac_launch_for(int GT, bool MX, const CParams& cp, dim3 grid, size_t smem, cudaStream_t st) {
    switch (GT * 2 + (MX ? 1 : 0)) {
        case 10: return mmac_launch<5, false>(cp, grid, smem, st);   case 11: return mmac_launch<5, true>(cp, grid, smem, st);
        case 12: return mmac_launch<6, false>(cp, grid, smem, st);   case 13: return mmac_launch<6, true>(cp, grid, smem, st);
        case 14: return mmac_launch<7, false>(cp, grid, smem, st);   case 15: return mmac_launch<7, true>(cp, grid, smem, st);
        case 16: return mmac_launch<8, false>(cp, grid, smem, st);   case 17: return mmac_launch<8, true>(cp, grid, smem, st);
        case 18: return mmac_launch<9, false>(cp, grid, smem, st);   case 19: return mmac_launch<9, true>(cp, grid, smem, st);
        case 20: return mmac_launch<10, false>(cp, grid, smem, st);  case 21: return mmac_launch<10, true>(cp, grid, smem, st);
        case 22: return mmac_launch<11, false>(cp, grid, smem, st);  case 23: return mmac_launch<11, true>(cp, grid, smem, st);
        case 24: return mmac_launch<12, false>(cp, grid, smem, st);  case 25: return mmac_launch<12, true>(cp, grid, smem, st);
        case 26: return mmac_launch<13, false>(cp, grid, smem, st);  case 27: return mmac_launch<13, true>(cp, grid, smem, st);
        case 28: return mmac_launch<14, false>(cp, grid, smem, st);  case 29: return mmac_launch<14, true>(cp, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

template <int GT, bool MX>
static cudaError_t mma2_launch(const M2Params& mp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&k_mma2<GT, MX>), smem);
        if (e != cudaSuccess) return e;
    }
    k_mma2<GT, MX><<<grid, threads, smem, st>>>(mp);
    return cudaGetLastError();
}
static cudaError_t mma2_launch_for(int GT, bool MX, const M2Params& mp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    switch (GT * 2 + (MX ? 1 : 0)) {
        case 10: return mma2_launch<5, false>(mp, grid, threads, smem, st);  case 11: return mma2_launch<5, true>(mp, grid, threads, smem, st);
        case 12: return mma2_launch<6, false>(mp, grid, threads, smem, st);  case 13: return mma2_launch<6, true>(mp, grid, threads, smem, st);
        case 14: return mma2_launch<7, false>(mp, grid, threads, smem, st);  case 15: return mma2_launch<7, true>(mp, grid, threads, smem, st);
    }
    return cudaErrorInvalidValue;
}

static Plan make_plan(const bildk_model* m, int P_per_traj_hint) {
    Plan pl{};
    {
        const char* force0 = getenv("BILDK_KERNEL");
        if (m->mmar2_ok && m->GT >= 10 && !force0 && !env_int("BILDK_FORCE_GENERIC", 0) && env_int("BILDK_MMAR8", 1)) {
            // k_mmar8: one filter per CTA, eight warps, one resident propagator
            const size_t matb = static_cast<size_t>(8 * m->GT) * m->LDr * 8;
            const size_t fbytes = static_cast<size_t>(m->fstride_r) * 8;
            pl.mmar2 = true;
            pl.mmar8 = true;
            pl.maxf = 1;
            pl.tile = false;
            pl.FPC = 1;
            pl.WPC = 1;
            pl.threads = 256;
            pl.smem = 16 + matb + fbytes;
            pl.fstride = m->fstride_r;
            pl.bstride = static_cast<int>(matb / 8);
            return pl;
        }
        if (m->mmar2_ok && m->GT <= 9 && !force0 && !env_int("BILDK_FORCE_GENERIC", 0) && env_int("BILDK_MMAR2", 1) && (!m->mmar_mx || env_int("BILDK_MMAR_MX", 1))) {
            const size_t matb = static_cast<size_t>(8 * m->GT) * m->LDr * 8;
            const size_t fbytes = static_cast<size_t>(m->fstride_r) * 8;
            const int maxf = mmar2_maxf(m->GT, m->mmar_mx);
            int f = env_int("BILDK_FPC2", 0);
            if (f <= 0) {
                // filters per CTA by  waves x time per wave.  Measured at N = 50 (T = 200, P = 16384, profiles/r02_mmar2_variants.txt):
                // 4 filters in one CTA per SM 52.3 ms, 2 filters x 2 CTAs per SM 53.1, 3 filters (one scheduler pair carries
                // two warps) 69.2, 1 filter x 3 CTAs per SM 59.3 - relative time of one full wave (all resident slots busy):
                // relative time of one full wave (all resident slots busy), per tile-grid size: GT = 7 as above; GT = 6
                // (N = 48, profiles/r02_mx_variants.txt) 2 filters x 2 CTAs per SM 0.977 of 4 x 1; GT = 5 (N = 40, 36; compiled for
                // 6 filters) 2 x 3 CTAs 0.983, 3 x 2 CTAs 0.985 of 6 x 1.  Counts that do not divide the compiled maximum leave
                // register file unused and are not candidates for GT < 7.
                static const double wt7[7] = {0.0, 0.85, 1.016, 0.99, 1.0, 1.0, 1.0};
                static const double wt6[7] = {0.0, 1.0, 0.977, 1.0, 1.0, 1.0, 1.0};
                static const double wt5[7] = {0.0, 1.0, 0.983, 0.985, 1.0, 1.0, 1.0};
                static const double wt8[7] = {0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
                const double* wave_time = m->GT == 7 ? wt7 : (m->GT == 6 ? wt6 : (m->GT >= 8 ? wt8 : wt5));
                const long long P = std::max(1, P_per_traj_hint);
                double best = 1e300;
                for (int c = maxf; c >= 1; --c) {
                    if (m->GT != 7 && maxf % c && c != 2) continue;
                    const size_t smem_c = 16 + matb * m->S + fbytes * c;
                    if (smem_c > static_cast<size_t>(m->max_smem_optin)) continue;   // e.g. three propagators next to three GT = 8 filters
                    const int per_sm = std::max<int>(1, std::min<int>(static_cast<int>((228 * 1024) / (smem_c + 1024)), maxf / c));
                    const long long n_cta = (P + c - 1) / c, slots = static_cast<long long>(m->n_sm) * per_sm;
                    const double cost = static_cast<double>((n_cta + slots - 1) / slots) * wave_time[c];
                    if (cost < best - 1e-9) { best = cost; f = c; }
                }
            }
            f = std::max(1, std::min(f, maxf));
            while (f > 1 && 16 + matb * m->S + fbytes * f > static_cast<size_t>(m->max_smem_optin)) --f;
            pl.mmar2 = true;
            pl.maxf = maxf;
            pl.tile = false;
            pl.FPC = f;
            pl.WPC = f;
            pl.threads = 32 * mmar2_nw(m->GT) * f;
            pl.smem = 16 + matb * m->S + fbytes * f;
            pl.fstride = m->fstride_r;
            pl.bstride = static_cast<int>(matb / 8);
            return pl;
        }
        if (m->mma_ok && m->GT >= 5 && m->GT <= 7 && !(force0 && (!strcmp(force0, "tile") || !strcmp(force0, "mma1") || !strcmp(force0, "mmag"))) &&
            !env_int("BILDK_FORCE_GENERIC", 0) && m->GT < env_int("BILDK_MMAC_MIN_GT", 8)) {
            const size_t matb = static_cast<size_t>(m->NPm) * m->LDBm * 8;
            const size_t fbytes = (static_cast<size_t>(m->NPm) * m->LDCm + static_cast<size_t>(2) * m->NPm + 2) * 8;
            const size_t cap = static_cast<size_t>(m->max_smem_optin);
            // filters per CTA (one CTA per SM: the propagators take 2 x 26 KB of it).  Measured time of ONE full wave at
            // N = 50 (T = 200, 148 f filters, profiles/r02_mma2_fpc_table.txt), relative: a single warp pair is latency bound,
            // odd counts leave one scheduler pair with an unmatched role.  Pick the count that minimises waves x time per
            // wave for this batch: P = 16384 -> 6 (19 waves), P = 1024 -> 4 (2 waves of 148 + 108 CTAs instead of 148 + 23).
            static const double wave_time[7] = {0.0, 1.70, 1.715, 2.33, 2.43, 3.33, 3.19};
            int fmax = 6;
            while (fmax > 1 && 16 + matb * m->S + fbytes * fmax > cap) --fmax;
            int f = env_int("BILDK_FPC2", 0);
            if (f <= 0) {
                double best = 1e300;
                f = fmax;
                for (int c = fmax; c >= 1; --c) {
                    const long long n_cta = (static_cast<long long>(std::max(1, P_per_traj_hint)) + c - 1) / c;
                    const double cost = static_cast<double>((n_cta + m->n_sm - 1) / m->n_sm) * wave_time[c];
                    if (cost < best - 1e-9) { best = cost; f = c; }
                }
            }
            f = std::min(f, fmax);
            if (16 + matb * m->S + fbytes * f <= cap) {
                pl.mma2 = true;
                pl.tile = false;
                pl.FPC = f;
                pl.WPC = f;
                pl.threads = 64 * f;
                pl.smem = 16 + matb * m->S + fbytes * f;
                pl.fstride = static_cast<int>(fbytes / 8);
                pl.bstride = static_cast<int>(matb / 8);
                return pl;
            }
        }
        const bool no_tc = (force0 && !strcmp(force0, "tile")) || env_int("BILDK_FORCE_GENERIC", 0);
        if (m->mmag_ok && !no_tc && (!m->mmac_ok || (force0 && !strcmp(force0, "mmag")))) {
            const int GT = m->GT;
            pl.mmag = true;
            pl.tile = false;
            pl.smem = (static_cast<size_t>(2) * m->NPm + 2) * 8;
            pl.threads = 32 * GT;
            pl.FPC = 1;
            const int ncw_env = env_int("BILDK_MMAG2", 3);   // tile columns per warp: 3 (default, measured best at N = 150, 200), 2, 4, or 0/1 = k_mmag
            if (ncw_env >= 2) {
                // NC adjacent tile columns per warp (k_mmag2): group j = columns [NC j, NC j + NC) of the GTC columns of
                // [C | M]; groups -> warps longest-processing-time first onto the four schedulers (warp i runs on scheduler i % 4)
                const int NC = ncw_env >= 4 ? 4 : ncw_env;
                const int GTC = GT + (m->mma_mx ? 1 : 0);
                const int ngrp = (GTC + NC - 1) / NC;
                pl.mmag_nc = NC;
                pl.threads = 32 * ngrp;
                double load2[4] = {0, 0, 0, 0};
                int slots2[4], used2[4] = {0, 0, 0, 0};
                for (int k = 0; k < 4; ++k) slots2[k] = (ngrp - k + 3) / 4;
                for (int j = ngrp - 1; j >= 0; --j) {   // P2 work grows with the group index
                    int best = -1;
                    for (int k = 0; k < 4; ++k)
                        if (used2[k] < slots2[k] && (best < 0 || load2[k] < load2[best])) best = k;
                    pl.colmap[best + 4 * used2[best]] = static_cast<unsigned char>(j);
                    const int cA = NC * j, ncw = std::min(NC, GTC - cA), ncu = std::max(0, std::min(NC, GT - cA));
                    load2[best] += ncw * GT + ncu * (cA + ncu);
                    ++used2[best];
                }
                return pl;
            }
            double load[4] = {0, 0, 0, 0};
            int slots[4], used[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; ++k) slots[k] = (GT - k + 3) / 4;
            for (int c = GT - 1; c >= 0; --c) {
                int best = -1;
                for (int k = 0; k < 4; ++k)
                    if (used[k] < slots[k] && (best < 0 || load[k] < load[best])) best = k;
                pl.colmap[best + 4 * used[best]] = static_cast<unsigned char>(c);
                load[best] += GT + c + 1 + ((c == 0 && m->mma_mx) ? GT : 0);
                ++used[best];
            }
            return pl;
        }
        if (m->mmac_ok && (m->GT >= env_int("BILDK_MMAC_MIN_GT", 8)) && !(force0 && !strcmp(force0, "tile")) && !env_int("BILDK_FORCE_GENERIC", 0)) {
            const int GT = m->GT;
            const size_t matb = static_cast<size_t>(m->NPm) * m->LDBm * 8;
            const size_t fbytes = (static_cast<size_t>(m->NPm) * m->LDCm + static_cast<size_t>(2) * m->NPm + 2) * 8;
            pl.mmac = true;
            pl.tile = false;
            pl.b_all = 16 + matb * m->S + fbytes <= static_cast<size_t>(m->max_smem_optin);
            pl.smem = 16 + matb * (pl.b_all ? m->S : 1) + fbytes;
            pl.threads = 32 * GT;
            pl.FPC = 1;
            // GT = 4k + 1 columns: three helper warps balance P1 across the schedulers (see k_mmac), so the column map
            // only has to balance P2 (weight c + 1); otherwise it balances P1 + P2 (weight GT + c + 1)
            pl.nhelp = (GT % 4 == 1 && GT >= 9 && !m->mma_mx && env_int("BILDK_MMAC_HELPERS", 1)) ? 3 : 0;
            pl.threads = 32 * (GT + pl.nhelp);
            // slots need the mean in the padding columns; measured against k_mmac (T=300, 2960 filters): GT=13 +7 %, GT=14 +13 %,
            // GT=8/10/11/12 -3..-14 % (their column warps are already balanced four-by-four)
            pl.mmact = !m->mma_mx && (GT == 9 || GT == 13 || GT == 14) && env_int("BILDK_MMACT", 1);
            if (pl.mmact) { pl.nhelp = mmact_helpers(GT); pl.threads = 32 * mmact_warps(GT); }
            // column -> warp: longest-processing-time first onto the four schedulers (warp i runs on scheduler i % 4)
            double load[4] = {0, 0, 0, 0};
            int slots[4], used[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; ++k) slots[k] = (GT - k + 3) / 4;
            for (int c = GT - 1; c >= 0; --c) {   // the weight decreases with c
                int best = -1;
                for (int k = 0; k < 4; ++k)
                    if (used[k] < slots[k] && (best < 0 || load[k] < load[best])) best = k;
                pl.colmap[best + 4 * used[best]] = static_cast<unsigned char>(c);
                load[best] += (pl.nhelp ? 0 : GT) + c + 1 + ((c == 0 && m->mma_mx) ? GT : 0);
                ++used[best];
            }
            return pl;
        }
    }
    {
        const char* force = getenv("BILDK_KERNEL");
        const bool want_mmar = m->mmar_ok && !(force && strcmp(force, "mmar")) && !env_int("BILDK_FORCE_GENERIC", 0) && env_int("BILDK_MMAR", 1) &&
                               (!m->mmar_mx || env_int("BILDK_MMAR_MX", 1));
        if (want_mmar) {
            // measured on B200 (profiles/r01_mmar_variants.txt): 4 warps per scheduler with ~128 registers beat 7 warps
            // with 72 (spills) for every GT; GT = 4 needs 3 per scheduler to stay spill-free
            int nb = env_int("BILDK_MMAR_NB", m->GT <= 3 ? 4 : 3);
            // border kernel: selected for N = 17, 25 (r = 1: +20 %, +23 %) and N = 26 (r = 2, GT = 4: +8 %); N = 18 and GT = 2 lose
            // against k_mmar (profiles/r02_border_variants.txt) - compiled and tested (BILDK_MMARB=2), not selected
            pl.mmarb = m->mmarb_ok && env_int("BILDK_MMARB", 1) >= (((m->GT >= 3 && m->r_last == 1) || (m->GT == 4 && m->r_last == 2)) ? 1 : 2);
            if (pl.mmarb ? !mmarb_has(m->GT, nb, m->r_last) : !mmar_has(m->GT, nb, m->mmar_mx)) nb = m->GT <= 3 ? 4 : 3;
            const size_t matb = static_cast<size_t>(8 * m->GT) * m->LDr * 8;
            const size_t fbytes = (static_cast<size_t>(m->fstride_r) + (pl.mmarb ? 16 * m->GT : 0)) * 8;
            pl.mmar = true;
            pl.nb = nb;
            pl.WPC = 4;
            pl.threads = 128;
            pl.smem = 16 + matb * m->S + fbytes * 4;
            pl.fstride = static_cast<int>(fbytes / 8);
            pl.bstride = static_cast<int>(matb / 8);
            pl.tile = false;
            return pl;
        }
    }
    {
        const char* force = getenv("BILDK_KERNEL");
        const bool want_mma = m->mma_ok && !(force && !strcmp(force, "tile")) && !env_int("BILDK_FORCE_GENERIC", 0);
        if (want_mma) {
            const size_t matb = static_cast<size_t>(m->NPm) * m->LDBm * 8;
            const size_t fbytes = (static_cast<size_t>(m->NPm) * m->LDCm + static_cast<size_t>(2) * m->NPm + 2) * 8;
            const size_t cap = static_cast<size_t>(m->max_smem_optin);
            const size_t sm_total = 228 * 1024;
            const int regs = mma_regs_for(m->GT, m->mma_mx);
            const int P = std::max(1, P_per_traj_hint);
            double best = -1;
            int best_w = 1;
            const int forced = env_int("BILDK_WPC", 0);
            const int wmax = m->GT <= 3 ? 28 : 8;   // __launch_bounds__ of k_mma
            for (int w = 1; w <= wmax; ++w) {
                if (forced && w != forced) continue;
                const size_t smem = 16 + matb * m->S + fbytes * w;
                if (smem > cap) break;
                int per_sm = static_cast<int>(sm_total / (smem + 1024));
                per_sm = std::min(per_sm, 65536 / (regs * 32 * w));
                per_sm = std::min(per_sm, 64 / w);
                if (per_sm < 1) continue;
                const long long slots = static_cast<long long>(m->n_sm) * per_sm * w;
                const long long waves = (P + slots - 1) / slots;
                // warp i of a CTA runs on scheduler i % 4: the busiest scheduler sets the pace
                const int busiest = per_sm * ((w + 3) / 4);
                const double balance = (per_sm * w / 4.0) / busiest;
                // tail efficiency of the last wave x scheduler balance x how well the resident warps can keep the
                // FP64 pipe fed (>= 2 warps per scheduler hide each other's store / update / barrier phases)
                const double eff = static_cast<double>(P) / (waves * slots) * balance * std::min(1.0, 0.6 + 0.4 * per_sm * w / 8.0) + 1e-4 * per_sm * w;
                if (eff > best) { best = eff; best_w = w; }
            }
            if (best > 0) {
                pl.mma = true;
                pl.WPC = best_w;
                pl.threads = 32 * best_w;
                pl.smem = 16 + matb * m->S + fbytes * best_w;
                pl.fstride = static_cast<int>(fbytes / 8);
                pl.bstride = static_cast<int>(matb / 8);
                pl.tile = false;
                return pl;
            }
        }
    }
    pl.tile = m->tile_ok && !env_int("BILDK_FORCE_GENERIC", 0);
    if (!pl.tile) return pl;
    const int G = m->G, TPF = G * G;
    pl.TS = m->TS;
    pl.densew = m->nnz > NZMAX;
    pl.ws = TPF <= 32;
    pl.maxt = maxt_for(m->TS, pl.ws);
    const size_t matb = static_cast<size_t>(m->NP) * m->LD * 8;
    const size_t fbytes = filter_doubles(m->NP, m->LD, pl.densew) * 8;
    pl.fstride = static_cast<int>(fbytes / 8);
    const size_t cap = static_cast<size_t>(m->max_smem_optin);
    pl.bstride = static_cast<int>(stride_8mod16(matb / 8));
    pl.b_all = 16 + static_cast<size_t>(pl.bstride) * 8 * m->S + fbytes <= cap;
    const size_t bbytes = pl.b_all ? static_cast<size_t>(pl.bstride) * 8 * m->S : matb;
    int fpc_smem = static_cast<int>((cap - 16 - bbytes) / fbytes);
    if (!pl.b_all) fpc_smem = 1;
    if (pl.ws) {
        pl.TPFS = 1;
        while (pl.TPFS < TPF) pl.TPFS *= 2;
        const int per_warp = 32 / pl.TPFS;
        int warps = env_int("BILDK_WARPS", 2);
        warps = std::max(1, std::min(warps, pl.maxt / 32));
        pl.FPC = std::min(warps * per_warp, std::max(per_warp, (fpc_smem / per_warp) * per_warp));
        pl.threads = (pl.FPC / per_warp) * 32;
    } else {
        pl.TPFS = TPF;
        int fpc = std::max(1, std::min(pl.maxt / TPF, fpc_smem));
        // keep enough CTAs to cover the SMs
        while (fpc > 1 && (P_per_traj_hint + fpc - 1) / fpc < 2 * m->n_sm) --fpc;
        const int forced = env_int("BILDK_FPC", 0);
        if (forced > 0) fpc = std::max(1, std::min(forced, std::min(pl.maxt / TPF, fpc_smem)));
        pl.FPC = fpc;
        pl.threads = ((fpc * TPF + 31) / 32) * 32;
    }
    pl.smem = 16 + bbytes + static_cast<size_t>(pl.FPC) * fbytes;
    return pl;
}

template <int TS, bool WS, bool DW, int MAXT>
static cudaError_t launch_one(const KParams& kp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&k_tile<TS, WS, DW, MAXT>), smem);
        if (e != cudaSuccess) return e;
    }
    k_tile<TS, WS, DW, MAXT><<<grid, threads, smem, st>>>(kp);
    return cudaGetLastError();
}

template <int TS>
static cudaError_t launch_ts(const Plan& pl, const KParams& kp, dim3 grid, cudaStream_t st) {
    constexpr int MT = (TS <= 6 ? 512 : 256);
    if (pl.ws) {
        if (pl.densew) return launch_one<TS, true, true, 128>(kp, grid, pl.threads, pl.smem, st);
        return launch_one<TS, true, false, 128>(kp, grid, pl.threads, pl.smem, st);
    }
    if (pl.densew) return launch_one<TS, false, true, MT>(kp, grid, pl.threads, pl.smem, st);
    return launch_one<TS, false, false, MT>(kp, grid, pl.threads, pl.smem, st);
}

static cudaError_t launch_tile(const Plan& pl, const KParams& kp, dim3 grid, cudaStream_t st) {
    switch (pl.TS) {
        case 4: return launch_ts<4>(pl, kp, grid, st);
        case 5: return launch_ts<5>(pl, kp, grid, st);
        case 6: return launch_ts<6>(pl, kp, grid, st);
        case 8: return launch_ts<8>(pl, kp, grid, st);
    }
    return cudaErrorInvalidValue;
}

static std::string plan_string(const bildk_model* m, const Plan& pl) {
    char buf[256];
    if (pl.mma2)
        snprintf(buf, sizeof buf, "mma2 (DMMA m8n8k4) GT=%d %s two-warps-per-filter FPC=%d threads=%d smem=%zu", m->GT,
                 m->mma_mx ? "mean-in-extra-tile" : "mean-in-padding", pl.FPC, pl.threads, pl.smem);
    else if (pl.mmag)
        snprintf(buf, sizeof buf, "mmag (DMMA m8n8k4) GT=%d %s cta-per-filter %s covariance-in-L2-workspace threads=%d", m->GT,
                 m->mma_mx ? "mean-in-extra-tile" : "mean-in-padding", pl.mmag_nc == 4 ? "warp-per-four-tile-columns" : pl.mmag_nc == 3 ? "warp-per-three-tile-columns" : pl.mmag_nc == 2 ? "warp-per-two-tile-columns" : "warp-per-tile-column", pl.threads);
    else if (pl.mmac)
        snprintf(buf, sizeof buf, "%s (DMMA m8n8k4) GT=%d %s cta-per-filter %s%s B=%s threads=%d smem=%zu", pl.mmact ? "mmact" : "mmac", m->GT,
                 m->mma_mx ? "mean-in-extra-tile" : "mean-in-padding", pl.mmact ? "tile-slots-per-warp" : "warp-per-tile-column",
                 pl.nhelp ? " +3-P1-helper-warps" : "", pl.b_all ? "all" : "one", pl.threads, pl.smem);
    else if (pl.mmar8)
        snprintf(buf, sizeof buf, "mmar8 (DMMA m8n8k4) GT=%d%s register-chained cta-per-filter eight-warps (tile rows split) B=one threads=%d smem=%zu", m->GT,
                 m->mmar_mx ? " mean-in-extra-rows" : "", pl.threads, pl.smem);
    else if (pl.mmar2)
        snprintf(buf, sizeof buf, "mmar2 (DMMA m8n8k4) GT=%d%s register-chained %s-warps-per-filter (tile rows split) FPC=%d of %d threads=%d smem=%zu", m->GT,
                 m->mmar_mx ? " mean-in-extra-rows" : "", m->GT == 9 ? "five" : (m->GT >= 8 ? "four" : "two"), pl.FPC, pl.maxf, pl.threads, pl.smem);
    else if (pl.mmar)
        snprintf(buf, sizeof buf, "mmar (DMMA m8n8k4) GT=%d%s register-chained warp-per-filter WPC=%d CTAs/SM=%d threads=%d smem=%zu", m->GT,
                 m->mmar_mx ? " mean-in-extra-rows" : (pl.mmarb ? " border-in-DFMA" : ""), pl.WPC, pl.nb, pl.threads, pl.smem);
    else if (pl.mma)
        snprintf(buf, sizeof buf, "mma (DMMA m8n8k4) GT=%d %s warp-per-filter WPC=%d threads=%d smem=%zu", m->GT,
                 m->mma_mx ? "mean-in-extra-tile" : "mean-in-padding", pl.WPC, pl.threads, pl.smem);
    else if (!pl.tile)
        snprintf(buf, sizeof buf, "generic N=%d (covariance in L2 workspace)", m->N);
    else
        snprintf(buf, sizeof buf, "tile TS=%d G=%d %s-scope %s-w B=%s FPC=%d threads=%d smem=%zu", pl.TS, m->G,
                 pl.ws ? "warp" : "cta", pl.densew ? "dense" : "sparse", pl.b_all ? "all" : "one", pl.FPC, pl.threads, pl.smem);
    return buf;
}

extern "C" const char* bildk_describe_plan(bildk_traj_t t, int P) {
    if (!t) return "";
    Plan pl = make_plan(t->m, P);
    t->plan = plan_string(t->m, pl);
    return t->plan.c_str();
}

extern "C" int bildk_debug_tables(int kernel, int GT, int r, int ncols, unsigned char* out) {
    if (!out) return fail(BILDK_EINVAL, "out is NULL");
    if (kernel == 0) {
        if (GT < 1 || GT > 13 || r < 1 || r > 8 || ncols < 1 || ncols > 4) return fail(BILDK_EINVAL, "k_mmar tables need GT in 1..13, r in 1..8, ncols in 1..4");
        mmar_tables(GT, r, ncols, out, out + 8);
        return 12;
    }
    if (kernel == 1) {
        if (GT < 8 || GT > 14) return fail(BILDK_EINVAL, "k_mmact tables need GT in 8..14");
        CTParams ct{};
        if (!mmact_slots(GT, ct)) return fail(BILDK_EINVAL, "internal: slot table for GT=%d", GT);
        for (int w = 0; w < 16; ++w) {
            out[18 * w] = ct.nslot[w];
            out[18 * w + 1] = ct.nsegA[w];
            for (int i = 0; i < 8; ++i) { out[18 * w + 2 + 2 * i] = ct.slot_ti[w][i]; out[18 * w + 3 + 2 * i] = ct.slot_c[w][i]; }
        }
        return 288;
    }
    return fail(BILDK_EINVAL, "unknown kernel %d", kernel);
}

// Launch maps of a multi-trajectory batch (CTA -> trajectory / first filter; filter -> trajectory): written into pinned host
// memory and copied asynchronously, so that enqueuing a batch never waits for the stream.
struct MapArena {
    int* h_cta = nullptr;   // pinned, capacity 2 (P + n_traj)
    int* d_cta = nullptr;
    int* h_pt = nullptr;    // pinned, capacity P
    int* d_pt = nullptr;
};

// Core launcher on device-resident profile arrays.  Trajectory metadata arrays are device pointers.
static int launch_device(bildk_model* m, const bildk_traj* t0, int n_traj, const double* const* d_x,
                         const uint8_t* const* d_valid, const int* d_T, const int* d_first,
                         const std::vector<int>& h_first, int P, int K1, const int32_t* d_starts,
                         const uint8_t* d_states, double* d_out, cudaStream_t st, const MapArena* arena = nullptr) {
    if (P == 0) return BILDK_OK;
    const int dstar = t0->dstar;
    int max_per_traj = 0;
    for (int i = 0; i < n_traj; ++i) max_per_traj = std::max(max_per_traj, h_first[i + 1] - h_first[i]);
    Plan pl = make_plan(m, n_traj == 1 ? P : P);
    double* d_part = d_out;
    if (dstar > 1) {
        int rc = m->part.reserve(static_cast<size_t>(dstar) * P);
        if (rc) return rc;
        d_part = m->part.p;
    }
    if (pl.mma || pl.mmar || pl.mmar2 || pl.mma2 || pl.mmac || pl.mmag || pl.tile) {
        KParams kp{};
        kp.N = m->N; kp.D = m->D; kp.S = m->S; kp.G = m->G; kp.LD = m->LD; kp.NP = m->NP;
        kp.Bpad = m->dBpad; kp.Sigpad = m->dSigpad; kp.C0pad = m->dC0pad; kp.Gm = m->dG; kp.M0 = m->dM0; kp.w = m->dw;
        kp.hasG = m->hasG; kp.nnz = m->nnz;
        for (int z = 0; z < NZMAX; ++z) { kp.wz_idx[z] = m->wz_idx[z]; kp.wz_val[z] = m->wz_val[z]; }
        kp.n_traj = n_traj; kp.x = d_x; kp.valid = d_valid; kp.T = d_T; kp.traj_first = d_first;
        kp.dstar = dstar;
        for (int e = 0; e < DMAX; ++e) {
            kp.s2[e] = t0->s2[e]; kp.ncols[e] = t0->ncols[e];
            for (int c = 0; c < DMAX; ++c) kp.cols[e][c] = t0->cols[e][c];
        }
        kp.P = P; kp.K1 = K1; kp.run_starts = d_starts; kp.run_states = d_states; kp.out = d_part;
        if (pl.mma || pl.mmar) pl.FPC = pl.WPC;   // CTA -> first filter maps use FPC (mma2 sets FPC itself)
        if (pl.mmac || pl.mmag) pl.FPC = 1;
        kp.FPC = pl.FPC; kp.TPFS = pl.TPFS; kp.b_all = pl.b_all; kp.fstride = pl.fstride; kp.bstride = pl.bstride; kp.lane_ab = m->d_lane_ab;
        int n_cta = 0;
        if (n_traj == 1) {
            n_cta = (P + pl.FPC - 1) / pl.FPC;
            kp.cta_traj = nullptr; kp.cta_first = nullptr;
        } else {
            if (!arena) return fail(BILDK_EINVAL, "internal: multi-trajectory launch without a map arena");
            for (int i = 0; i < n_traj; ++i)
                for (int f = h_first[i]; f < h_first[i + 1]; f += pl.FPC) ++n_cta;
            int k = 0;
            for (int i = 0; i < n_traj; ++i)
                for (int f = h_first[i]; f < h_first[i + 1]; f += pl.FPC, ++k) { arena->h_cta[k] = i; arena->h_cta[n_cta + k] = f; }
            CU(cudaMemcpyAsync(arena->d_cta, arena->h_cta, static_cast<size_t>(2) * n_cta * sizeof(int), cudaMemcpyHostToDevice, st));
            kp.cta_traj = arena->d_cta; kp.cta_first = arena->d_cta + n_cta;
        }
        dim3 grid(n_cta, dstar);
        if (pl.mmag) {
            GMParams gp{};
            MParams& mp = gp.m;
            mp.k = kp;
            mp.k.cta_traj = nullptr; mp.k.cta_first = nullptr;
            mp.NPm = m->NPm; mp.LDB = m->LDBm; mp.LDC = m->LDCm; mp.MC0 = m->MC0; mp.NK = m->NK;
            mp.Bm = m->dBm; mp.Sigm = m->dSigm; mp.C0m = m->dC0m;
            for (int i = 0; i < 40; ++i) gp.colmap[i] = pl.colmap[i];
            gp.prof_traj = nullptr;
            if (n_traj > 1) {
                for (int i = 0; i < n_traj; ++i) for (int f = h_first[i]; f < h_first[i + 1]; ++f) arena->h_pt[f] = i;
                CU(cudaMemcpyAsync(arena->d_pt, arena->h_pt, P * sizeof(int), cudaMemcpyHostToDevice, st));
                gp.prof_traj = arena->d_pt;
            }
            const int nc = std::min(P, 2 * m->n_sm);
            const size_t wsz = static_cast<size_t>(2) * m->NPm * m->LDCm;
            int rc = m->work.reserve(wsz * nc * dstar);
            if (rc) return rc;
            gp.work = m->work.p;
            // row-chunk size by the registers the warp count leaves (registers are allocated per scheduler)
            auto waste = [&](int chv) { return ((m->GT + chv - 1) / chv) * chv; };   // padded tile rows per column
            const int ch = env_int("BILDK_MMAG_CH", m->GT <= 16 ? 8 : ((m->GT <= 28 && waste(6) <= waste(4)) ? 6 : 4));
            if (pl.mmag_nc == 4) {        // <= 8 warps per CTA
                k_mmag2<4, 4, 256, 2><<<dim3(nc, dstar), pl.threads, pl.smem, st>>>(gp, m->GT, m->mma_mx ? 1 : 0);
            } else if (pl.mmag_nc == 3) {   // <= 11 warps per CTA
                if (pl.threads <= 288) k_mmag2<4, 3, 288, 2><<<dim3(nc, dstar), pl.threads, pl.smem, st>>>(gp, m->GT, m->mma_mx ? 1 : 0);
                else k_mmag2<4, 3, 352, 2><<<dim3(nc, dstar), pl.threads, pl.smem, st>>>(gp, m->GT, m->mma_mx ? 1 : 0);
            } else if (pl.mmag_nc == 2) {
                if (pl.threads <= 448) k_mmag2<4, 2, 448, 2><<<dim3(nc, dstar), pl.threads, pl.smem, st>>>(gp, m->GT, m->mma_mx ? 1 : 0);
                else k_mmag2<4, 2, 512, 2><<<dim3(nc, dstar), pl.threads, pl.smem, st>>>(gp, m->GT, m->mma_mx ? 1 : 0);
            } else
            if (ch >= 8 && m->GT <= 16) k_mmag<8, 512><<<dim3(nc, dstar), pl.threads, pl.smem, st>>>(gp, m->GT, m->mma_mx ? 1 : 0);
            else if (ch >= 6 && m->GT <= 28) k_mmag<6, 896><<<dim3(nc, dstar), pl.threads, pl.smem, st>>>(gp, m->GT, m->mma_mx ? 1 : 0);
            else k_mmag<4, 1024><<<dim3(nc, dstar), pl.threads, pl.smem, st>>>(gp, m->GT, m->mma_mx ? 1 : 0);
            CU(cudaGetLastError());
        } else if (pl.mmar) {
            RParams rp{};
            rp.k = kp;
            rp.Br = m->dBr; rp.Sigm = m->dSigm; rp.C0m = m->dC0m;
            rp.WPC = pl.WPC; rp.fstride = pl.fstride; rp.r = m->r_last;
            rp.ww00 = m->wz_val[0] * m->wz_val[0]; rp.ww11 = m->wz_val[1] * m->wz_val[1]; rp.ww01 = 2.0 * m->wz_val[0] * m->wz_val[1];
            for (int e = 0; e < dstar; ++e) mmar_tables(m->GT, m->r_last, t0->ncols[e], rp.lastrow[e], rp.mrow[e]);
            if (pl.mmarb) CU(mmarb_launch_for(m->GT, pl.nb, m->r_last, rp, grid, pl.threads, pl.smem, st));
            else CU(mmar_launch_for(m->GT, pl.nb, m->mmar_mx, rp, grid, pl.threads, pl.smem, st));
        } else if (pl.mmar2) {
            R2Params r2{};
            RParams& rp = r2.r;
            rp.k = kp;
            rp.Br = m->dBr; rp.Sigm = m->dSigm; rp.C0m = m->dC0m;
            rp.WPC = pl.FPC; rp.fstride = pl.fstride; rp.r = m->r_last;
            rp.ww00 = m->wz_val[0] * m->wz_val[0]; rp.ww11 = m->wz_val[1] * m->wz_val[1]; rp.ww01 = 2.0 * m->wz_val[0] * m->wz_val[1];
            for (int e = 0; e < dstar; ++e) mmar_tables(m->GT, m->r_last, t0->ncols[e], rp.lastrow[e], rp.mrow[e]);
            r2.FPC2 = pl.FPC;
            if (pl.mmar8) {
                const cudaError_t e8 = mmar8_launch_for(m->GT, m->mmar_mx, r2, grid, pl.smem, st);
                CU(e8);
            } else {
                CU(mmar2_launch_for(m->GT, pl.maxf, m->mmar_mx, r2, grid, pl.threads, pl.smem, st));
            }
        } else if (pl.mma2) {
            M2Params m2{};
            MParams& mp = m2.m;
            mp.k = kp;
            mp.NPm = m->NPm; mp.LDB = m->LDBm; mp.LDC = m->LDCm; mp.MC0 = m->MC0; mp.NK = m->NK;
            mp.Bm = m->dBm; mp.Sigm = m->dSigm; mp.C0m = m->dC0m;
            mp.WPC = pl.FPC; mp.fstride_m = pl.fstride; mp.bstride_m = pl.bstride;
            m2.FPC2 = pl.FPC;
            CU(mma2_launch_for(m->GT, m->mma_mx, m2, grid, pl.threads, pl.smem, st));
        } else if (pl.mma || pl.mmac) {
            MParams mp{};
            mp.k = kp;
            mp.NPm = m->NPm; mp.LDB = m->LDBm; mp.LDC = m->LDCm; mp.MC0 = m->MC0; mp.NK = m->NK;
            mp.Bm = m->dBm; mp.Sigm = m->dSigm; mp.C0m = m->dC0m;
            mp.WPC = pl.WPC; mp.fstride_m = pl.fstride; mp.bstride_m = pl.bstride;
            if (pl.mmac) {
                CParams cp{};
                cp.m = mp;
                cp.b_all = pl.b_all;
                cp.nhelp = pl.nhelp;
                if (pl.mmact) {   // every phase balanced over GT + 3 warps
                    CTParams ct{};
                    ct.m = mp;
                    ct.b_all = pl.b_all;
                    if (!mmact_slots(m->GT, ct)) return fail(BILDK_EINVAL, "internal: slot table for GT=%d", m->GT);
                    CU(mmact_launch_for(m->GT, ct, grid, pl.smem, st));
                } else {
                    for (int i = 0; i < 16; ++i) cp.colmap[i] = pl.colmap[i];   // GT <= 14
                    CU(mmac_launch_for(m->GT, m->mma_mx, cp, grid, pl.smem, st));
                }
            } else
            CU(mma_launch_for(m->GT, m->mma_mx, mp, grid, pl.threads, pl.smem, st));
        } else {
            CU(launch_tile(pl, kp, grid, st));
        }
        g_launches++;
    } else {
        GParams gp{};
        gp.N = m->N; gp.D = m->D; gp.S = m->S;
        gp.B = m->dB; gp.Sig = m->dSig; gp.C0 = m->dC0; gp.Gm = m->dG; gp.M0 = m->dM0; gp.w = m->dw;
        gp.n_traj = n_traj; gp.x = d_x; gp.valid = d_valid; gp.T = d_T; gp.traj_first = d_first;
        gp.dstar = dstar;
        for (int e = 0; e < DMAX; ++e) {
            gp.s2[e] = t0->s2[e]; gp.ncols[e] = t0->ncols[e];
            for (int c = 0; c < DMAX; ++c) gp.cols[e][c] = t0->cols[e][c];
        }
        gp.P = P; gp.K1 = K1; gp.run_starts = d_starts; gp.run_states = d_states; gp.out = d_part;
        gp.prof_traj = nullptr;
        if (n_traj > 1) {
            if (!arena) return fail(BILDK_EINVAL, "internal: multi-trajectory launch without a map arena");
            for (int i = 0; i < n_traj; ++i) for (int f = h_first[i]; f < h_first[i + 1]; ++f) arena->h_pt[f] = i;
            CU(cudaMemcpyAsync(arena->d_pt, arena->h_pt, P * sizeof(int), cudaMemcpyHostToDevice, st));
            gp.prof_traj = arena->d_pt;
        }
        const int n_cta = std::min(P, 4 * m->n_sm);
        const size_t wsz = 2 * static_cast<size_t>(m->N) * m->N + 3 * static_cast<size_t>(m->N) * m->D + 2 * m->N;
        int rc = m->work.reserve(wsz * n_cta * dstar);
        if (rc) return rc;
        gp.work = m->work.p;
        k_generic<<<dim3(n_cta, dstar), 256, 0, st>>>(gp);
        CU(cudaGetLastError());
        g_launches++;
    }
    if (dstar > 1) {
        k_sum_parts<<<(P + 255) / 256, 256, 0, st>>>(d_part, d_out, P, dstar);
        CU(cudaGetLastError());
        g_launches++;
    }
    return BILDK_OK;
}

static int validate_runs(const bildk_model* m, int T, int P, int K1, const int32_t* starts, const uint8_t* states, int pbase) {
    for (int p = 0; p < P; ++p) {
        const int32_t* rs = starts + static_cast<size_t>(p) * K1;
        const uint8_t* rt = states + static_cast<size_t>(p) * K1;
        if (rs[0] != 0) return fail(BILDK_EINVAL, "profile %d: first run must start at frame 0 (got %d)", pbase + p, rs[0]);
        for (int r = 0; r < K1; ++r) {
            if (rt[r] >= m->S) return fail(BILDK_EINVAL, "profile %d: state %d out of range [0,%d)", pbase + p, rt[r], m->S);
            if (r && rs[r] < rs[r - 1]) return fail(BILDK_EINVAL, "profile %d: run starts must be non-decreasing", pbase + p);
            if (rs[r] < 0 || rs[r] > T) return fail(BILDK_EINVAL, "profile %d: run start %d outside [0,%d]", pbase + p, rs[r], T);
        }
    }
    return BILDK_OK;
}

extern "C" int bildk_logl_runs_device(bildk_traj_t t, int P, int K1, const int32_t* d_starts,
                                      const uint8_t* d_states, double* d_out, void* stream) {
    if (!t) return fail(BILDK_EINVAL, "NULL trajectory");
    if (P < 0 || K1 < 1) return fail(BILDK_EINVAL, "need P >= 0 and K1 >= 1");
    if (P == 0) return BILDK_OK;
    if (!d_starts || !d_states || !d_out) return fail(BILDK_EINVAL, "NULL device pointer");
    bildk_model* m = t->m;
    NvtxRange nvtx("bildk_logl_runs_device");
    std::lock_guard<std::mutex> lock(m->mu);   // covers the launch bookkeeping only: the work itself is asynchronous
    CU(cudaSetDevice(m->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int first[2] = {0, P};
    CU(cudaMemcpyAsync(t->d_first, first, sizeof first, cudaMemcpyHostToDevice, st));
    std::vector<int> hf = {0, P};
    return launch_device(m, t, 1, t->d_xptr, t->d_vptr, t->d_T, t->d_first, hf, P, K1, d_starts, d_states, d_out, st);
}

// ------------------------------------------------------------------------------------------------
// Device-resident AMIS ensemble (bildk_amis.cuh): one call per AMIS iteration.
//
// A dataset run creates thousands of ensembles (one per FixedkSampler: ~10 per trajectory), so creating one must be
// cheap: the stream, the pinned staging buffer and the device staging block are shared PER DEVICE, and an ensemble is
// two stream-ordered allocations (cudaMallocAsync from the device's memory pool: microseconds once the pool is warm) -
// one block for the samples, one for the proposals - that grow by doubling.
struct AmisDevice {
    std::mutex mu;                 // one AMIS step at a time per device (the calls are synchronous anyway)
    cudaStream_t st = nullptr;
    double* pinned = nullptr;      // host staging (inputs) / copy-back (head + 3 n)
    size_t cap_pinned = 0;
    double* stage = nullptr;       // device staging of one step's inputs
    size_t cap_stage = 0;
};
static AmisDevice* amis_device(int device) {
    static std::mutex mu;
    static std::map<int, AmisDevice*> devs;
    std::lock_guard<std::mutex> lock(mu);
    AmisDevice*& d = devs[device];
    if (!d) d = new AmisDevice();       // lives until process exit
    return d;
}

struct bildk_amis {
    int device = 0, K1 = 0, S = 0;
    int n = 0, n_par = 0;
    std::vector<uint8_t> transitions;
    AmisDevice* dev = nullptr;
    // samples block: ss [cap][K1] | logs [cap][K1] | logL [cap] | per [cap][3] | thetas [cap][K1] bytes | flags [cap] bytes
    char* sblock = nullptr;
    size_t cap = 0;
    // proposals block: A [capp][K1] | lognorm [capp] | norm0 [capp] | logp [capp][S K1] | reach [capp][S K1]
    char* pblock = nullptr;
    size_t capp = 0;
    double* head = nullptr;        // [4 + 2 K1 + S K1]
    double* ss() const { return reinterpret_cast<double*>(sblock); }
    double* logs() const { return ss() + cap * K1; }
    double* logL() const { return logs() + cap * K1; }
    double* per() const { return logL() + cap; }
    uint8_t* thetas() const { return reinterpret_cast<uint8_t*>(per() + 3 * cap); }
    uint8_t* flags() const { return thetas() + cap * K1; }
    static size_t sbytes(size_t cap, int K1) { return (cap * (2 * static_cast<size_t>(K1) + 4)) * 8 + (cap * (static_cast<size_t>(K1) + 1) + 15) / 16 * 16; }
    double* A() const { return reinterpret_cast<double*>(pblock); }
    double* lognorm() const { return A() + capp * K1; }
    double* norm0() const { return lognorm() + capp; }
    double* logp() const { return norm0() + capp; }
    double* reach() const { return logp() + capp * S * K1; }
    static size_t pbytes(size_t capp, int K1, int S) { return capp * (static_cast<size_t>(K1) + 2 + 2 * static_cast<size_t>(S) * K1) * 8; }
};

extern "C" int bildk_amis_destroy(bildk_amis_t h) {
    if (!h) return BILDK_OK;
    cudaSetDevice(h->device);
    cudaStream_t st = h->dev ? h->dev->st : nullptr;
    if (h->sblock) cudaFreeAsync(h->sblock, st);
    if (h->pblock) cudaFreeAsync(h->pblock, st);
    if (h->head) cudaFreeAsync(h->head, st);
    delete h;
    return BILDK_OK;
}

extern "C" int bildk_amis_create(int K1, int S, const uint8_t* transitions, int device, bildk_amis_t* out) {
    if (!out) return fail(BILDK_EINVAL, "out is NULL");
    *out = nullptr;
    if (K1 < 1 || S < 1 || !transitions) return fail(BILDK_EINVAL, "bad argument");
    if (K1 > AMIS_MAXK1 || S > AMIS_MAXS)
        return fail(BILDK_EUNSUP, "the device AMIS ensemble handles up to %d slots and %d states (got %d, %d)", AMIS_MAXK1, AMIS_MAXS, K1, S);
    int ndev = bildk_device_count();
    if (ndev == 0) return fail(BILDK_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(BILDK_EINVAL, "device %d out of range", device);
    CU(cudaSetDevice(device));
    AmisDevice* dev = amis_device(device);
    std::lock_guard<std::mutex> lock(dev->mu);
    if (!dev->st) {
        CU(cudaStreamCreateWithFlags(&dev->st, cudaStreamNonBlocking));
        cudaMemPool_t pool;                                   // keep freed blocks in the pool instead of returning them to the OS
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    bildk_amis* h = new bildk_amis();
    struct Guard { bildk_amis* h; ~Guard() { if (h) bildk_amis_destroy(h); } } guard{h};
    h->device = device; h->K1 = K1; h->S = S; h->dev = dev;
    h->transitions.assign(transitions, transitions + static_cast<size_t>(S) * S);
    CU(cudaMallocAsync(&h->head, (4 + 2 * static_cast<size_t>(K1) + static_cast<size_t>(S) * K1) * sizeof(double), dev->st));
    guard.h = nullptr;
    *out = h;
    return BILDK_OK;
}

extern "C" int bildk_amis_size(bildk_amis_t h, int* n_samples, int* n_proposals) {
    if (!h) return fail(BILDK_EINVAL, "NULL handle");
    if (n_samples) *n_samples = h->n;
    if (n_proposals) *n_proposals = h->n_par;
    return BILDK_OK;
}

// grow the samples / proposals block (stream ordered: the copies queue behind earlier steps, the old block is freed after them)
static int amis_grow(bildk_amis* h, size_t need_samples, size_t need_par, cudaStream_t st) {
    const int K1 = h->K1, S = h->S;
    if (need_samples > h->cap) {
        bildk_amis old = *h;
        const size_t cap = std::max<size_t>(std::max<size_t>(need_samples, 2 * h->cap), 1024);
        char* blk = nullptr;
        cudaError_t e = cudaMallocAsync(&blk, bildk_amis::sbytes(cap, K1), st);
        if (e != cudaSuccess) return fail(BILDK_ENOMEM, "cudaMallocAsync(%zu bytes): %s", bildk_amis::sbytes(cap, K1), cudaGetErrorString(e));
        h->sblock = blk; h->cap = cap;
        if (old.sblock && old.n) {
            const size_t n = old.n;
            CU(cudaMemcpyAsync(h->ss(), old.ss(), n * K1 * 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(h->logs(), old.logs(), n * K1 * 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(h->logL(), old.logL(), n * 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(h->per(), old.per(), 3 * n * 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(h->thetas(), old.thetas(), n * K1, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(h->flags(), old.flags(), n, cudaMemcpyDeviceToDevice, st));
        }
        if (old.sblock) CU(cudaFreeAsync(old.sblock, st));
    }
    if (need_par > h->capp) {
        bildk_amis old = *h;
        const size_t capp = std::max<size_t>(std::max<size_t>(need_par, 2 * h->capp), 32);
        char* blk = nullptr;
        cudaError_t e = cudaMallocAsync(&blk, bildk_amis::pbytes(capp, K1, S), st);
        if (e != cudaSuccess) return fail(BILDK_ENOMEM, "cudaMallocAsync(%zu bytes): %s", bildk_amis::pbytes(capp, K1, S), cudaGetErrorString(e));
        h->pblock = blk; h->capp = capp;
        if (old.pblock && old.n_par) {
            const size_t n = old.n_par, sk = static_cast<size_t>(S) * K1;
            CU(cudaMemcpyAsync(h->A(), old.A(), n * K1 * 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(h->lognorm(), old.lognorm(), n * 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(h->norm0(), old.norm0(), n * 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(h->logp(), old.logp(), n * sk * 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(h->reach(), old.reach(), n * sk * 8, cudaMemcpyDeviceToDevice, st));
        }
        if (old.pblock) CU(cudaFreeAsync(old.pblock, st));
    }
    return BILDK_OK;
}

// Staging of one AMIS step (doubles): ss (n K1) | prop: A (K1) lognorm norm0 logp (S K1) reach (S K1) | thetas bytes (n K1)
static size_t amis_stage_doubles(int n_new, int K1, int S) {
    const size_t nk = static_cast<size_t>(n_new) * K1, sk = static_cast<size_t>(S) * K1;
    return nk + K1 + 2 + 2 * sk + (nk + 7) / 8;
}
static size_t amis_head_doubles(int K1, int S) { return 4 + 2 * static_cast<size_t>(K1) + static_cast<size_t>(S) * K1; }

// Fill the pinned staging of one step (see amis_stage_doubles); the normalisers of the joining proposal
// (amis.py:83-108, 258-281) are computed here, once per proposal.
static void amis_fill_stage(const bildk_amis* h, int n_new, const double* ss, const int64_t* thetas, const double* A_cur,
                            const double* logp_cur, double* pin) {
    const int K1 = h->K1, S = h->S;
    const size_t nk = static_cast<size_t>(n_new) * K1, sk = static_cast<size_t>(S) * K1;
    const size_t o_A = nk, o_ln = o_A + K1, o_n0 = o_ln + 1, o_lp = o_n0 + 1, o_rc = o_lp + sk, o_th = o_rc + sk;
    std::memcpy(pin, ss, nk * sizeof(double));
    std::memcpy(pin + o_A, A_cur, K1 * sizeof(double));
    const double inf = std::numeric_limits<double>::infinity();
    auto lse = [&](const double* v, int stride, const uint8_t* allowed) {
        double top = -inf;
        for (int m = 0; m < S; ++m)
            if (!allowed || allowed[m]) top = std::max(top, v[m * stride]);
        if (!std::isfinite(top)) top = 0.0;
        double acc = 0.0;
        for (int m = 0; m < S; ++m)
            if (!allowed || allowed[m]) acc += std::exp(v[m * stride] - top);
        return std::log(acc) + top;
    };
    double asum = 0.0, lg = 0.0;
    for (int c = 0; c < K1; ++c) { asum += A_cur[c]; lg += std::lgamma(A_cur[c]); }
    pin[o_ln] = std::lgamma(asum) - lg;
    pin[o_n0] = lse(logp_cur, K1, nullptr);
    std::memcpy(pin + o_lp, logp_cur, sk * sizeof(double));
    for (int m = 0; m < S; ++m) {
        pin[o_rc + static_cast<size_t>(m) * K1] = 0.0;
        for (int c = 1; c < K1; ++c) pin[o_rc + static_cast<size_t>(m) * K1 + c] = lse(logp_cur + c, K1, h->transitions.data() + static_cast<size_t>(m) * S);
    }
    uint8_t* thb = reinterpret_cast<uint8_t*>(pin + o_th);
    for (size_t i = 0; i < nk; ++i) thb[i] = static_cast<uint8_t>(thetas[i]);
}

// Describe one AMIS step as a device job: the step's inputs will be on the device at `d_stage` (layout of
// amis_stage_doubles), the new likelihoods at `d_logL` (device, valid in stream order - e.g. the output of the filter
// kernel enqueued just before); the statistics go to `d_head` (device).  Grows the ensemble's blocks (stream ordered on
// `st`) and updates its sizes.  Nothing here waits for the GPU.
static int amis_make_job(bildk_amis* h, int n_new, const double* d_stage, const double* d_logL, double* d_head, cudaStream_t st, AmisJob* job) {
    const size_t n_old = h->n, n_tot = n_old + n_new, n_par = h->n_par + 1;
    int rc = amis_grow(h, n_tot, n_par, st);
    if (rc) return rc;
    AmisParams& ap = job->p;
    ap = AmisParams{};
    ap.n_old = static_cast<int>(n_old); ap.n_new = n_new; ap.K1 = h->K1; ap.S = h->S; ap.n_par = static_cast<int>(n_par);
    ap.logs = h->logs(); ap.thetas = h->thetas(); ap.flags = h->flags(); ap.ss = h->ss(); ap.logL = h->logL(); ap.per = h->per();
    ap.A = h->A(); ap.lognorm = h->lognorm(); ap.logp = h->logp(); ap.reach = h->reach(); ap.norm0 = h->norm0();
    ap.log_nsteps = std::log(static_cast<double>(n_par));
    ap.out = d_head;
    job->stage = d_stage;
    job->logL_new = d_logL;
    h->n = static_cast<int>(n_tot);
    h->n_par = static_cast<int>(n_par);
    return BILDK_OK;
}

// Launch the jobs at `d_jobs` (device copy; the first n16 have K1 <= 16, the next n32 more slots - the column width of
// the reduction is a function of the ensemble alone, so a result never depends on its companions): one append launch and
// one cluster launch per width class, for all ensembles together.
static int amis_launch(const AmisJob* d_jobs, int n16, int n32, int max_new, cudaStream_t st) {
    const int n_jobs = n16 + n32;
    if (n_jobs == 0) return BILDK_OK;
    k_amis_append<<<dim3((max_new + 127) / 128, n_jobs), 128, 0, st>>>(d_jobs);
    CU(cudaGetLastError());
    g_launches++;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(AMIS_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = AMIS_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (n16) {
        cfg.gridDim = dim3(AMIS_CLUSTER, n16);
        CU(cudaLaunchKernelEx(&cfg, k_amis_step<16>, d_jobs));
        g_launches++;
    }
    if (n32) {
        cfg.gridDim = dim3(AMIS_CLUSTER, n32);
        CU(cudaLaunchKernelEx(&cfg, k_amis_step<32>, d_jobs + n16));
        g_launches++;
    }
    return BILDK_OK;
}

extern "C" int bildk_amis_step(bildk_amis_t h, int n_new, const double* ss, const int64_t* thetas, const double* logL,
                               const double* A_cur, const double* logp_cur, double* head, double* per_sample) {
    if (!h) return fail(BILDK_EINVAL, "NULL handle");
    if (n_new < 1 || !ss || !thetas || !logL || !A_cur || !logp_cur || !head) return fail(BILDK_EINVAL, "bad argument");
    const int K1 = h->K1, S = h->S;
    const size_t nk = static_cast<size_t>(n_new) * K1;
    for (size_t i = 0; i < nk; ++i)
        if (thetas[i] < 0 || thetas[i] >= S) return fail(BILDK_EINVAL, "state %lld out of range [0,%d)", static_cast<long long>(thetas[i]), S);
    NvtxRange nvtx("bildk_amis_step");
    AmisDevice* dev = h->dev;
    std::lock_guard<std::mutex> lock(dev->mu);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = dev->st;
    const size_t n_tot = static_cast<size_t>(h->n) + n_new;
    // ---- one packed upload: [job | step staging | logL]
    constexpr size_t n_job = sizeof(AmisJob) / sizeof(double);
    static_assert(sizeof(AmisJob) % sizeof(double) == 0, "the job sits in front of the staged doubles");
    const size_t n_stage = amis_stage_doubles(n_new, K1, S), in_doubles = n_job + n_stage + n_new;
    const size_t n_head = amis_head_doubles(K1, S);
    const size_t out_doubles = n_head + (per_sample ? 3 * n_tot : 0);
    const size_t need_pinned = std::max(in_doubles, out_doubles);
    if (need_pinned > dev->cap_pinned) {
        if (dev->pinned) cudaFreeHost(dev->pinned);
        dev->pinned = nullptr; dev->cap_pinned = 0;
        const size_t want = std::max<size_t>(need_pinned * 2, 1 << 16);
        cudaError_t e = cudaMallocHost(&dev->pinned, want * sizeof(double));
        if (e != cudaSuccess) return fail(BILDK_ENOMEM, "cudaMallocHost(%zu bytes): %s", want * sizeof(double), cudaGetErrorString(e));
        dev->cap_pinned = want;
    }
    if (in_doubles > dev->cap_stage) {
        if (dev->stage) cudaFree(dev->stage);
        dev->stage = nullptr; dev->cap_stage = 0;
        const size_t want = std::max<size_t>(in_doubles * 2, 1 << 14);
        cudaError_t e = cudaMalloc(&dev->stage, want * sizeof(double));
        if (e != cudaSuccess) return fail(BILDK_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
        dev->cap_stage = want;
    }
    double* pin = dev->pinned;
    int rc = amis_make_job(h, n_new, dev->stage + n_job, dev->stage + n_job + n_stage, h->head, st, reinterpret_cast<AmisJob*>(pin));
    if (rc) return rc;
    amis_fill_stage(h, n_new, ss, thetas, A_cur, logp_cur, pin + n_job);
    std::memcpy(pin + n_job + n_stage, logL, n_new * sizeof(double));
    CU(cudaMemcpyAsync(dev->stage, pin, in_doubles * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = amis_launch(reinterpret_cast<const AmisJob*>(dev->stage), K1 <= 16 ? 1 : 0, K1 <= 16 ? 0 : 1, n_new, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(pin, h->head, n_head * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (per_sample) CU(cudaMemcpyAsync(pin + n_head, h->per(), 3 * n_tot * sizeof(double), cudaMemcpyDeviceToHost, st));
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(BILDK_ECUDA, "AMIS step failed: %s", cudaGetErrorString(e));
    std::memcpy(head, pin, n_head * sizeof(double));
    if (per_sample) std::memcpy(per_sample, pin + n_head, 3 * n_tot * sizeof(double));
    return BILDK_OK;
}

// ------------------------------------------------------------------------------------------------
static size_t align16(size_t x) { return (x + 15) / 16 * 16; }

extern "C" int bildk_logl_wait(void* ticket) {
    if (!ticket) return BILDK_OK;                      // an empty batch has no ticket
    bildk_model::Slot* sl = static_cast<bildk_model::Slot*>(ticket);
    if (!sl->busy) return fail(BILDK_EINVAL, "ticket is not in flight");
    NvtxRange nvtx("bildk_logl_wait");
    cudaError_t e = cudaEventSynchronize(sl->done);
    int rc = BILDK_OK;
    if (e != cudaSuccess) rc = fail(BILDK_ECUDA, "kernel or copy-back failed: %s", cudaGetErrorString(e));
    else {
        std::memcpy(sl->user_out, sl->pin + sl->off_out, static_cast<size_t>(sl->P) * sizeof(double));
        for (const auto& pa : sl->amis) {
            std::memcpy(pa.head, sl->pin + pa.off_head, pa.n_head * sizeof(double));
            if (pa.per) std::memcpy(pa.per, sl->pin + pa.off_per, pa.n_per * sizeof(double));
        }
    }
    std::lock_guard<std::mutex> lock(sl->owner->mu);
    sl->amis.clear();
    sl->busy = false;
    return rc;
}

extern "C" int bildk_logl_ready(void* ticket) {
    if (!ticket) return 1;                             // an empty batch is always done
    bildk_model::Slot* sl = static_cast<bildk_model::Slot*>(ticket);
    if (!sl->busy) return fail(BILDK_EINVAL, "ticket is not in flight");
    cudaError_t e = cudaEventQuery(sl->done);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) return 0;
    return fail(BILDK_ECUDA, "kernel or copy-back failed: %s", cudaGetErrorString(e));
}

extern "C" int bildk_logl_runs_multi_submit(int n_traj, const bildk_traj_t* trajs, const int32_t* offsets, int K1,
                                            const int32_t* starts, const uint8_t* states, double* out,
                                            const bildk_amis_req* amis, void** ticket) {
    if (!ticket) return fail(BILDK_EINVAL, "ticket is NULL");
    *ticket = nullptr;
    if (n_traj < 1 || !trajs || !offsets) return fail(BILDK_EINVAL, "need at least one trajectory");
    if (K1 < 1) return fail(BILDK_EINVAL, "K1 must be >= 1");
    bildk_model* m = trajs[0] ? trajs[0]->m : nullptr;
    if (!m) return fail(BILDK_EINVAL, "NULL trajectory");
    const int P = offsets[n_traj];
    if (offsets[0] != 0 || P < 0) return fail(BILDK_EINVAL, "offsets must start at 0");
    if (P == 0) return BILDK_OK;
    if (!starts || !states || !out) return fail(BILDK_EINVAL, "NULL array argument");
    std::vector<int> hf(offsets, offsets + n_traj + 1);
    for (int i = 0; i < n_traj; ++i) {
        bildk_traj* t = trajs[i];
        if (!t || t->m != m) return fail(BILDK_EINVAL, "all trajectories must belong to one model");
        if (hf[i + 1] < hf[i]) return fail(BILDK_EINVAL, "offsets must be non-decreasing");
        if (t->dstar != trajs[0]->dstar) return fail(BILDK_EINVAL, "trajectories in one batch must share the localisation-error structure");
        for (int e = 0; e < t->dstar; ++e) {
            if (t->s2[e] != trajs[0]->s2[e] || t->ncols[e] != trajs[0]->ncols[e]) return fail(BILDK_EINVAL, "trajectories in one batch must share the localisation error");
            for (int c = 0; c < t->ncols[e]; ++c) if (t->cols[e][c] != trajs[0]->cols[e][c]) return fail(BILDK_EINVAL, "trajectories in one batch must share the localisation error");
        }
        int rc = validate_runs(m, t->T, hf[i + 1] - hf[i], K1, starts + static_cast<size_t>(hf[i]) * K1, states + static_cast<size_t>(hf[i]) * K1, hf[i]);
        if (rc) return rc;
    }
    // fused AMIS bookkeeping (optional): validate, and size the staging / result areas
    size_t amis_in = 0, amis_heads = 0, amis_per = 0;
    std::vector<int> amis_order;                  // trajectories with a fused step: K1 <= 16 first, then the wider ones
    int amis_n16 = 0;
    if (amis) {
        for (int i = 0; i < n_traj; ++i) {
            const bildk_amis_req& rq = amis[i];
            const int n_i = hf[i + 1] - hf[i];
            if (!rq.ens || n_i == 0) continue;
            if (!rq.ss || !rq.thetas || !rq.A_cur || !rq.logp_cur || !rq.head) return fail(BILDK_EINVAL, "AMIS request %d: NULL array", i);
            if (rq.ens->device != m->device) return fail(BILDK_EINVAL, "AMIS request %d: ensemble lives on another device", i);
            const size_t nk = static_cast<size_t>(n_i) * rq.ens->K1;
            for (size_t e = 0; e < nk; ++e)
                if (rq.thetas[e] < 0 || rq.thetas[e] >= rq.ens->S) return fail(BILDK_EINVAL, "AMIS request %d: state out of range", i);
            amis_in += amis_stage_doubles(n_i, rq.ens->K1, rq.ens->S);
            amis_heads += amis_head_doubles(rq.ens->K1, rq.ens->S);
            if (rq.per_sample) amis_per += 3 * (static_cast<size_t>(rq.ens->n) + n_i);
            for (int j : amis_order)
                if (amis[j].ens == rq.ens) return fail(BILDK_EINVAL, "AMIS request %d: the same ensemble twice in one batch", i);
            amis_order.push_back(i);
        }
        std::stable_partition(amis_order.begin(), amis_order.end(), [&](int i) { return amis[i].ens->K1 <= 16; });
        for (int i : amis_order) amis_n16 += amis[i].ens->K1 <= 16;
    }
    const size_t amis_jobs = amis_order.size();
    NvtxRange nvtx("bildk_logl_runs_multi_submit");
    std::lock_guard<std::mutex> lock(m->mu);
    CU(cudaSetDevice(m->device));
    bildk_model::Slot* sl = nullptr;
    for (auto& cand : m->slots)
        if (!cand.busy) { sl = &cand; break; }
    if (!sl) return fail(BILDK_EINVAL, "both batch slots of this model are in flight: call bildk_logl_wait first");
    // ---- slot layout (bytes): run starts | run states | T per trajectory | first filter per trajectory | x pointers |
    //      valid-mask pointers | jobs and staging of the fused AMIS steps | CTA maps | filter -> trajectory | out | AMIS statistics
    //      (device mirror up to here) | per-sample AMIS arrays (pinned only: copied straight from the ensembles)
    const size_t nrun = static_cast<size_t>(P) * K1;
    const size_t o_st = 0, o_rs = align16(o_st + nrun * sizeof(int32_t)), o_T = align16(o_rs + nrun), o_first = align16(o_T + n_traj * sizeof(int)),
                 o_x = align16(o_first + (n_traj + 1) * sizeof(int)), o_v = align16(o_x + n_traj * sizeof(void*)),
                 o_jobs = align16(o_v + n_traj * sizeof(void*)), o_ain = align16(o_jobs + amis_jobs * sizeof(AmisJob)),
                 o_cta = align16(o_ain + amis_in * sizeof(double)),
                 o_pt = align16(o_cta + 2 * (static_cast<size_t>(P) + n_traj) * sizeof(int)),
                 o_out = align16(o_pt + static_cast<size_t>(P) * sizeof(int)), o_heads = o_out + static_cast<size_t>(P) * sizeof(double),
                 total_dev = align16(o_heads + amis_heads * sizeof(double)), total = align16(total_dev + amis_per * sizeof(double));
    if (total > sl->cap) {
        if (sl->pin) cudaFreeHost(sl->pin);
        sl->pin = nullptr; sl->cap = 0;
        const size_t want = total * 2;
        cudaError_t e = cudaMallocHost(&sl->pin, want);
        if (e != cudaSuccess) return fail(BILDK_ENOMEM, "pinned slot allocation (%zu bytes): %s", want, cudaGetErrorString(e));
        sl->cap = want;
    }
    if (total_dev > sl->cap_dev) {
        if (sl->dev) cudaFree(sl->dev);
        sl->dev = nullptr; sl->cap_dev = 0;
        const size_t want = total_dev * 2;
        cudaError_t e = cudaMalloc(&sl->dev, want);
        if (e != cudaSuccess) return fail(BILDK_ENOMEM, "device slot allocation (%zu bytes): %s", want, cudaGetErrorString(e));
        sl->cap_dev = want;
    }
    std::memcpy(sl->pin + o_st, starts, nrun * sizeof(int32_t));
    std::memcpy(sl->pin + o_rs, states, nrun);
    int* hT = reinterpret_cast<int*>(sl->pin + o_T);
    int* hfirst = reinterpret_cast<int*>(sl->pin + o_first);
    const double** hx = reinterpret_cast<const double**>(sl->pin + o_x);
    const uint8_t** hv = reinterpret_cast<const uint8_t**>(sl->pin + o_v);
    for (int i = 0; i < n_traj; ++i) { hT[i] = trajs[i]->T; hx[i] = trajs[i]->dx; hv[i] = trajs[i]->dvalid; }
    for (int i = 0; i <= n_traj; ++i) hfirst[i] = hf[i];
    // Two batches in flight run CONCURRENTLY on the device (one stream per slot) when they share no per-model scratch:
    // d* = 1 (no partial-logL buffer) and a kernel that keeps its filter state on chip (no L2 workspace).  Otherwise
    // both slots use the model's stream and the second batch queues behind the first.
    bool own_stream = trajs[0]->dstar == 1 && sl == &m->slots[1];
    if (own_stream) {
        const Plan pl = make_plan(m, P / n_traj);
        own_stream = pl.mma || pl.mmar || pl.mmar2 || pl.mma2 || pl.mmac || pl.tile;
    }
    cudaStream_t st = own_stream ? m->st2 : m->st;
    sl->st = st;
    // ---- fused AMIS bookkeeping: every trajectory's step consumes its likelihoods straight from the kernel's output,
    //      in stream order behind the filter kernel - no host round trip, no extra synchronisation
    sl->amis.clear();
    int amis_max_new = 0;
    std::unique_lock<std::mutex> amis_lock;       // the ensembles change size here: keep synchronous bildk_amis_step calls out
    if (amis_jobs) {
        amis_lock = std::unique_lock<std::mutex>(amis[amis_order[0]].ens->dev->mu);
        double* ain = reinterpret_cast<double*>(sl->pin + o_ain);
        AmisJob* jobs = reinterpret_cast<AmisJob*>(sl->pin + o_jobs);
        size_t off_ain = o_ain, off_head = o_heads;
        for (size_t j = 0; j < amis_jobs; ++j) {
            const int i = amis_order[j];
            const bildk_amis_req& rq = amis[i];
            const int n_i = hf[i + 1] - hf[i];
            const size_t n_head = amis_head_doubles(rq.ens->K1, rq.ens->S);
            amis_fill_stage(rq.ens, n_i, rq.ss, rq.thetas, rq.A_cur, rq.logp_cur, ain);
            int rc = amis_make_job(rq.ens, n_i, reinterpret_cast<const double*>(sl->dev + off_ain),
                                   reinterpret_cast<const double*>(sl->dev + o_out) + hf[i], reinterpret_cast<double*>(sl->dev + off_head), st, &jobs[j]);
            if (rc) return rc;
            sl->amis.push_back(bildk_model::Slot::PendingAmis{rq.head, off_head, n_head, rq.per_sample, 0, 0, rq.ens});
            ain += amis_stage_doubles(n_i, rq.ens->K1, rq.ens->S);
            off_ain += amis_stage_doubles(n_i, rq.ens->K1, rq.ens->S) * sizeof(double);
            off_head += n_head * sizeof(double);
            amis_max_new = std::max(amis_max_new, n_i);
        }
    }
    CU(cudaMemcpyAsync(sl->dev, sl->pin, o_cta, cudaMemcpyHostToDevice, st));     // everything up to the CTA maps in one copy
    MapArena arena;
    arena.h_cta = reinterpret_cast<int*>(sl->pin + o_cta); arena.d_cta = reinterpret_cast<int*>(sl->dev + o_cta);
    arena.h_pt = reinterpret_cast<int*>(sl->pin + o_pt); arena.d_pt = reinterpret_cast<int*>(sl->dev + o_pt);
    int rc = launch_device(m, trajs[0], n_traj, reinterpret_cast<const double* const*>(sl->dev + o_x),
                           reinterpret_cast<const uint8_t* const*>(sl->dev + o_v), reinterpret_cast<const int*>(sl->dev + o_T),
                           reinterpret_cast<const int*>(sl->dev + o_first), hf, P, K1, reinterpret_cast<const int32_t*>(sl->dev + o_st),
                           reinterpret_cast<const uint8_t*>(sl->dev + o_rs), reinterpret_cast<double*>(sl->dev + o_out), st, &arena);
    if (rc) return rc;
    if (amis_jobs) {
        rc = amis_launch(reinterpret_cast<const AmisJob*>(sl->dev + o_jobs), amis_n16, static_cast<int>(amis_jobs) - amis_n16, amis_max_new, st);
        if (rc) return rc;
        size_t off_per = total_dev;
        for (auto& pa : sl->amis) {
            if (!pa.per) continue;
            pa.off_per = off_per;
            pa.n_per = 3 * static_cast<size_t>(pa.ens->n);
            CU(cudaMemcpyAsync(sl->pin + off_per, pa.ens->per(), pa.n_per * sizeof(double), cudaMemcpyDeviceToHost, st));
            off_per += pa.n_per * sizeof(double);
        }
    }
    CU(cudaMemcpyAsync(sl->pin + o_out, sl->dev + o_out, static_cast<size_t>(P) * sizeof(double) + amis_heads * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(sl->done, st));
    sl->user_out = out; sl->off_out = o_out; sl->P = P; sl->busy = true;
    *ticket = sl;
    return BILDK_OK;
}

extern "C" int bildk_logl_runs_multi(int n_traj, const bildk_traj_t* trajs, const int32_t* offsets, int K1,
                                     const int32_t* starts, const uint8_t* states, double* out) {
    void* ticket = nullptr;
    int rc = bildk_logl_runs_multi_submit(n_traj, trajs, offsets, K1, starts, states, out, nullptr, &ticket);
    if (rc) return rc;
    return bildk_logl_wait(ticket);
}

extern "C" int bildk_logl_runs(bildk_traj_t t, int P, int K1, const int32_t* starts, const uint8_t* states, double* out) {
    if (!t) return fail(BILDK_EINVAL, "NULL trajectory");
    if (P < 0) return fail(BILDK_EINVAL, "P must be >= 0");
    int32_t off[2] = {0, P};
    bildk_traj_t arr[1] = {t};
    return bildk_logl_runs_multi(1, arr, off, K1, starts, states, out);
}

extern "C" int bildk_logl_st(bildk_traj_t t, int P, int K1, const double* ss, const int64_t* thetas, double* out) {
    if (!t) return fail(BILDK_EINVAL, "NULL trajectory");
    if (P < 0 || K1 < 1) return fail(BILDK_EINVAL, "need P >= 0 and K1 >= 1");
    if (P == 0) return BILDK_OK;
    if (!ss || !thetas || !out) return fail(BILDK_EINVAL, "NULL array argument");
    const int T = t->T, S = t->m->S;
    thread_local std::vector<int32_t> rs;
    thread_local std::vector<uint8_t> rt;
    rs.resize(static_cast<size_t>(P) * K1);
    rt.resize(static_cast<size_t>(P) * K1);
    const double Tm1 = static_cast<double>(T - 1);
    for (int p = 0; p < P; ++p) {
        const double* s = ss + static_cast<size_t>(p) * K1;
        const int64_t* th = thetas + static_cast<size_t>(p) * K1;
        int32_t* r = rs.data() + static_cast<size_t>(p) * K1;
        uint8_t* q = rt.data() + static_cast<size_t>(p) * K1;
        double cs = 0.0;   // np.cumsum: sequential adds in index order
        r[0] = 0;
        for (int i = 0; i < K1; ++i) {
            if (th[i] < 0 || th[i] >= S) return fail(BILDK_EINVAL, "profile %d: state %lld out of range [0,%d)", p, static_cast<long long>(th[i]), S);
            q[i] = static_cast<uint8_t>(th[i]);
            if (i + 1 < K1) {
                cs += s[i];
                const double f = std::floor(cs * Tm1);
                if (!(f >= -1.0 && f <= 2147483000.0)) return fail(BILDK_EINVAL, "profile %d: interval lengths are not finite", p);
                r[i + 1] = static_cast<int32_t>(f) + 1;
            }
        }
    }
    return bildk_logl_runs(t, P, K1, rs.data(), rt.data(), out);
}

extern "C" int bildk_logl_states(bildk_traj_t t, int P, const int32_t* states, double* out) {
    if (!t) return fail(BILDK_EINVAL, "NULL trajectory");
    if (P < 0) return fail(BILDK_EINVAL, "P must be >= 0");
    if (P == 0) return BILDK_OK;
    if (!states || !out) return fail(BILDK_EINVAL, "NULL array argument");
    const int T = t->T, S = t->m->S;
    // run-length code every profile; K1 = longest
    std::vector<int> nruns(P);
    int K1 = 1;
    for (int p = 0; p < P; ++p) {
        const int32_t* s = states + static_cast<size_t>(p) * T;
        int n = 1;
        for (int i = 0; i < T; ++i) {
            if (s[i] < 0 || s[i] >= S) return fail(BILDK_EINVAL, "profile %d frame %d: state %d out of range [0,%d)", p, i, s[i], S);
            if (i && s[i] != s[i - 1]) ++n;
        }
        nruns[p] = n;
        K1 = std::max(K1, n);
    }
    std::vector<int32_t> rs(static_cast<size_t>(P) * K1, T);
    std::vector<uint8_t> rt(static_cast<size_t>(P) * K1, 0);
    for (int p = 0; p < P; ++p) {
        const int32_t* s = states + static_cast<size_t>(p) * T;
        int r = 0;
        rs[static_cast<size_t>(p) * K1] = 0;
        rt[static_cast<size_t>(p) * K1] = static_cast<uint8_t>(s[0]);
        for (int i = 1; i < T; ++i)
            if (s[i] != s[i - 1]) {
                ++r;
                rs[static_cast<size_t>(p) * K1 + r] = i;
                rt[static_cast<size_t>(p) * K1 + r] = static_cast<uint8_t>(s[i]);
            }
        for (++r; r < K1; ++r) rt[static_cast<size_t>(p) * K1 + r] = rt[static_cast<size_t>(p) * K1 + r - 1];
    }
    return bildk_logl_runs(t, P, K1, rs.data(), rt.data(), out);
}

// ------------------------------------------------------------------------------------------------
// one cluster of AW_CLUSTER CTAs (bildk_kernels.cuh)
static cudaError_t launch_amis_weights(int n, const double* logL, const double* logdelta, const double* curlp, double log_nsteps,
                                       double* log_w, double* stats, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(AW_CLUSTER);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = AW_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_amis_weights, n, logL, logdelta, curlp, log_nsteps, log_w, stats);
    if (e == cudaSuccess) g_launches++;
    return e;
}

extern "C" int bildk_amis_weights(int n, const double* logL, const double* logdelta, const double* curlp,
                                  double log_nsteps, double* log_w, double stats[4], int device) {
    if (n < 1 || !logL || !logdelta || !curlp || !stats) return fail(BILDK_EINVAL, "bad argument");
    int ndev = bildk_device_count();
    if (ndev == 0) return fail(BILDK_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(BILDK_EINVAL, "device %d out of range", device);
    CU(cudaSetDevice(device));
    // one grow-only scratch buffer per device and thread, one packed copy in, one out (a cudaMalloc / cudaFree pair
    // and five small copies per call used to cost more than the reduction itself: the AMIS loop calls this every step)
    // ... on a private non-blocking stream: a call from one host thread never queues behind a filter launch that
    // another thread has in flight on the default stream
    struct Scratch { DevBuf<double> dev; std::vector<double> host; cudaStream_t st = nullptr; };
    static thread_local std::vector<Scratch> scratch;
    if (scratch.size() < static_cast<size_t>(ndev)) scratch.resize(ndev);
    Scratch& sc = scratch[device];
    const size_t nn = static_cast<size_t>(n);
    int rc = sc.dev.reserve(4 * nn + 4);
    if (rc) return rc;
    if (sc.host.size() < 3 * nn + 4) sc.host.resize(2 * (3 * nn + 4));
    if (!sc.st) CU(cudaStreamCreateWithFlags(&sc.st, cudaStreamNonBlocking));
    double* d = sc.dev.p;
    std::memcpy(sc.host.data(), logL, nn * 8);
    std::memcpy(sc.host.data() + nn, logdelta, nn * 8);
    std::memcpy(sc.host.data() + 2 * nn, curlp, nn * 8);
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d, sc.host.data(), 3 * nn * 8, cudaMemcpyHostToDevice, sc.st)) != cudaSuccess)
        return fail(BILDK_ECUDA, "copy failed: %s", cudaGetErrorString(e));
    NvtxRange nvtx("bildk_amis_weights");
    e = launch_amis_weights(n, d, d + nn, d + 2 * nn, log_nsteps, log_w ? d + 3 * nn : nullptr, d + 4 * nn, sc.st);
    // log_w (n) and the four statistics are contiguous on the device: [3n, 4n + 4)
    const size_t off = log_w ? 3 * nn : 4 * nn, cnt = log_w ? nn + 4 : 4;
    if (e != cudaSuccess ||
        (e = cudaMemcpyAsync(sc.host.data(), d + off, cnt * 8, cudaMemcpyDeviceToHost, sc.st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(sc.st)) != cudaSuccess)
        return fail(BILDK_ECUDA, "weights kernel failed: %s", cudaGetErrorString(e));
    if (log_w) std::memcpy(log_w, sc.host.data(), nn * 8);
    std::memcpy(stats, sc.host.data() + (log_w ? nn : 0), 4 * 8);
    return BILDK_OK;
}

extern "C" int bildk_marginal_posterior(int n, int K1, int T, int S, const int32_t* run_starts, const uint8_t* run_states,
                                        const double* log_w, double* out, int device) {
    if (n < 1 || K1 < 1 || T < 1 || S < 1 || S > 255 || !run_starts || !run_states || !log_w || !out) return fail(BILDK_EINVAL, "bad argument");
    int ndev = bildk_device_count();
    if (ndev == 0) return fail(BILDK_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(BILDK_EINVAL, "device %d out of range", device);
    const size_t nn = static_cast<size_t>(n), nr = nn * K1;
    for (size_t i = 0; i < nr; ++i) if (run_states[i] >= S) return fail(BILDK_EINVAL, "state %d out of range [0,%d)", run_states[i], S);
    CU(cudaSetDevice(device));
    NvtxRange nvtx("bildk_marginal_posterior");
    // grow-only scratch per device and thread, one packed copy in, one out, private stream (as bildk_amis_weights)
    struct Scratch { DevBuf<double> dev; std::vector<double> host; cudaStream_t st = nullptr; };
    static thread_local std::vector<Scratch> scratch;
    if (scratch.size() < static_cast<size_t>(ndev)) scratch.resize(ndev);
    Scratch& sc = scratch[device];
    // packed layout in units of 8 bytes: log_w (n) | out (S*T) | run_starts (ceil(nr/2)) | run_states (ceil(nr/8))
    const size_t n_out = static_cast<size_t>(S) * T, o_out = nn, o_starts = o_out + n_out, o_states = o_starts + (nr + 1) / 2,
                 total = o_states + (nr + 7) / 8;
    int rc = sc.dev.reserve(total);
    if (rc) return rc;
    if (sc.host.size() < total) sc.host.resize(2 * total);
    if (!sc.st) CU(cudaStreamCreateWithFlags(&sc.st, cudaStreamNonBlocking));
    std::memcpy(sc.host.data(), log_w, nn * 8);
    std::memcpy(sc.host.data() + o_starts, run_starts, nr * sizeof(int32_t));
    std::memcpy(sc.host.data() + o_states, run_states, nr);
    double* d = sc.dev.p;
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d, sc.host.data(), nn * 8, cudaMemcpyHostToDevice, sc.st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d + o_starts, sc.host.data() + o_starts, (total - o_starts) * 8, cudaMemcpyHostToDevice, sc.st)) != cudaSuccess)
        return fail(BILDK_ECUDA, "copy failed: %s", cudaGetErrorString(e));
    const int cache_n = nn <= 40 * 1024 ? n : 0;   // one byte of shared memory per sample (48 KB need no opt-in)
    k_marginal_posterior<<<T, 256, static_cast<size_t>(cache_n), sc.st>>>(n, K1, T, S, reinterpret_cast<const int32_t*>(d + o_starts),
                                                                        reinterpret_cast<const uint8_t*>(d + o_states), d, d + o_out, cache_n);
    g_launches++;
    if ((e = cudaGetLastError()) != cudaSuccess ||
        (e = cudaMemcpyAsync(sc.host.data() + o_out, d + o_out, n_out * 8, cudaMemcpyDeviceToHost, sc.st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(sc.st)) != cudaSuccess)
        return fail(BILDK_ECUDA, "marginal posterior kernel failed: %s", cudaGetErrorString(e));
    std::memcpy(out, sc.host.data() + o_out, n_out * 8);
    return BILDK_OK;
}

// Host-side AMIS bookkeeping: proposal densities of n samples under n_par proposals (see include/bild_b200.h).
extern "C" int bildk_amis_log_proposal(int n_par, int n, int K1, int S, const double* A, const double* logp,
                                       const uint8_t* transitions, const double* ss, const int64_t* thetas, double* out) {
    if (n_par < 0 || n < 0 || K1 < 1 || S < 1 || S > 255) return fail(BILDK_EINVAL, "bad sizes (n_par=%d n=%d K1=%d S=%d)", n_par, n, K1, S);
    if (n_par == 0 || n == 0) return BILDK_OK;
    if (!A || !logp || !transitions || !ss || !thetas || !out) return fail(BILDK_EINVAL, "NULL array argument");
    for (size_t i = 0; i < static_cast<size_t>(n) * K1; ++i)
        if (thetas[i] < 0 || thetas[i] >= S) return fail(BILDK_EINVAL, "state %lld out of range [0,%d)", static_cast<long long>(thetas[i]), S);
    const double inf = std::numeric_limits<double>::infinity();
    // log-sum-exp over the entries selected by `allowed`; -inf for an empty selection (amis.py uses scipy's logsumexp)
    auto lse = [&](const double* v, int stride, const uint8_t* allowed) {
        double top = -inf;
        for (int m = 0; m < S; ++m)
            if (!allowed || allowed[m]) top = std::max(top, v[m * stride]);
        if (!std::isfinite(top)) top = 0.0;
        double acc = 0.0;
        for (int m = 0; m < S; ++m)
            if (!allowed || allowed[m]) acc += std::exp(v[m * stride] - top);
        return std::log(acc) + top;
    };
    std::vector<double> logs(static_cast<size_t>(n) * K1);   // log(s), shared by all proposals
    std::vector<uint8_t> bad0(n), haszero(n);
    for (int i = 0; i < n; ++i) {
        double sum = 0.0;
        bool bad = false, z = false;
        for (int c = 0; c < K1; ++c) {
            const double v = ss[static_cast<size_t>(i) * K1 + c];
            sum += v;
            if (!(v >= 0.0) || v > 1.0) bad = true;
            if (v == 0.0) z = true;
            logs[static_cast<size_t>(i) * K1 + c] = v > 0.0 ? std::log(v) : -inf;
        }
        if (std::fabs(sum - 1.0) > 1e-9 || std::isnan(sum)) bad = true;
        bad0[i] = bad; haszero[i] = z;
    }
    std::vector<double> reach(static_cast<size_t>(S) * K1);   // per proposal: LSE over the states reachable from m at slot c
    for (int j = 0; j < n_par; ++j) {
        const double* a = A + static_cast<size_t>(j) * K1;
        const double* lp = logp + static_cast<size_t>(j) * S * K1;   // [S][K1]
        double asum = 0.0, lg = 0.0;
        for (int c = 0; c < K1; ++c) { asum += a[c]; lg += std::lgamma(a[c]); }
        const double lognorm = std::lgamma(asum) - lg;
        const double norm0 = lse(lp, K1, nullptr);
        for (int m = 0; m < S; ++m)
            for (int c = 1; c < K1; ++c) reach[static_cast<size_t>(m) * K1 + c] = lse(lp + c, K1, transitions + static_cast<size_t>(m) * S);
        double* o = out + static_cast<size_t>(j) * n;
        for (int i = 0; i < n; ++i) {
            const double* ls = logs.data() + static_cast<size_t>(i) * K1;
            const int64_t* th = thetas + static_cast<size_t>(i) * K1;
            double dir = lognorm;
            bool bad = bad0[i];
            for (int c = 0; c < K1; ++c) {
                if (a[c] != 1.0) dir += (a[c] - 1.0) * ls[c];            // xlogy(a - 1, s): zero when a == 1, even at s == 0
                if (haszero[i] && ls[c] == -inf && a[c] < 1.0) bad = true;
            }
            double cat = lp[th[0] * K1] - norm0;
            for (int c = 1; c < K1; ++c) cat += lp[th[c] * K1 + c] - reach[static_cast<size_t>(th[c - 1]) * K1 + c];
            o[i] = (bad ? inf : dir) + cat;
        }
    }
    return BILDK_OK;
}

extern "C" int bildk_amis_weights_device(int n, const double* d_logL, const double* d_logdelta, const double* d_curlp,
                                         double log_nsteps, double* d_log_w, double* d_stats, void* stream) {
    if (n < 1 || !d_logL || !d_logdelta || !d_curlp || !d_stats) return fail(BILDK_EINVAL, "bad argument");
    NvtxRange nvtx("bildk_amis_weights_device");
    CU(launch_amis_weights(n, d_logL, d_logdelta, d_curlp, log_nsteps, d_log_w, d_stats, static_cast<cudaStream_t>(stream)));
    return BILDK_OK;
}

// ------------------------------------------------------------------------------------------------
// ChoiceSampler arithmetic (/root/reference/bild/choicesampler.py:112-175), host side.  Once the likelihood batch is one
// launch, the k-selection heuristic is what a long `bild.sample` run spends its host time on: every AMIS step evaluates the
// selection rule 2 kmax + 2 times on 10 000 Monte-Carlo draws x kmax candidates through numpy temporaries (28 ms per step at
// kmax = 11; one trajectory with 700 steps held an 8-GPU dataset run at 24 s while the other ranks were done after 9 s).
// These helpers do the same arithmetic - the same double additions, the same comparisons, so the same integers - in one pass
// over the draws; the random numbers stay numpy's.

// pick of one draw: first k whose value lies within dE of the row maximum, NaN entries ignored
// (`np.nanargmax(np.nanmax(draws) - dE - draws <= 0)`, choicesampler.py:131-133)
static inline int choice_pick_row(const double* d, int kmax, double dE) {
    double top = -std::numeric_limits<double>::infinity();
    bool any = false;
    for (int j = 0; j < kmax; ++j)
        if (d[j] == d[j]) { top = any ? std::max(top, d[j]) : d[j]; any = true; }
    if (!any) return 0;
    const double thr = top - dE;
    for (int j = 0; j < kmax; ++j)
        if (thr - d[j] <= 0.0) return j;          // false for NaN
    return 0;
}

extern "C" int bildk_choice_pick(int samplesize, int kmax, const double* scaled_rvs, const double* mu, double dE, int64_t* picks) {
    if (samplesize < 0 || kmax < 1 || kmax > 256) return fail(BILDK_EINVAL, "bad sizes (samplesize=%d kmax=%d)", samplesize, kmax);
    if (!scaled_rvs || !mu || !picks) return fail(BILDK_EINVAL, "NULL array argument");
    double d[256];
    for (int i = 0; i < samplesize; ++i) {
        const double* r = scaled_rvs + static_cast<size_t>(i) * kmax;
        for (int j = 0; j < kmax; ++j) d[j] = r[j] + mu[j];
        picks[i] = choice_pick_row(d, kmax, dE);
    }
    return BILDK_OK;
}

extern "C" int bildk_choice_dn(int samplesize, int kmax, const double* scaled_rvs, const double* muhat, const double* Dmu, double dE,
                               int64_t* dn) {
    if (samplesize < 0 || kmax < 1 || kmax > 256) return fail(BILDK_EINVAL, "bad sizes (samplesize=%d kmax=%d)", samplesize, kmax);
    if (!scaled_rvs || !muhat || !Dmu || !dn) return fail(BILDK_EINVAL, "NULL array argument");
    std::fill(dn, dn + static_cast<size_t>(kmax) * kmax, 0);
    double mu_lo[256], mu_hi[256], d[256];
    for (int k = 0; k < kmax; ++k) {       // `mu[k_change] += n_step * self.Dmu[k_change]` (choicesampler.py:126-127)
        mu_lo[k] = muhat[k] + (-0.5) * Dmu[k];
        mu_hi[k] = muhat[k] + 0.5 * Dmu[k];
    }
    for (int i = 0; i < samplesize; ++i) {
        const double* r = scaled_rvs + static_cast<size_t>(i) * kmax;
        // the unshifted draws and their two largest values: the row maximum with column k replaced is max(top without k, new value)
        double top1 = -std::numeric_limits<double>::infinity(), top2 = top1;
        int arg1 = -1;
        for (int j = 0; j < kmax; ++j) {
            d[j] = r[j] + muhat[j];
            if (d[j] > top1) { top2 = top1; top1 = d[j]; arg1 = j; }
            else if (d[j] > top2) top2 = d[j];
        }
        for (int k = 0; k < kmax; ++k) {
            const double rest = (k == arg1) ? top2 : top1;
            const double keep = d[k];
            for (int sgn = 0; sgn < 2; ++sgn) {
                const double v = r[k] + (sgn ? mu_hi[k] : mu_lo[k]);
                const double thr = std::max(rest, v) - dE;
                d[k] = v;
                int pick = 0;
                for (int j = 0; j < kmax; ++j)
                    if (thr - d[j] <= 0.0) { pick = j; break; }
                dn[static_cast<size_t>(k) * kmax + pick] += sgn ? 1 : -1;
            }
            d[k] = keep;
        }
    }
    return BILDK_OK;
}
