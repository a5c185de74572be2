"""
ORACLE - TEST INFRASTRUCTURE ONLY.  Nothing under bild_b200/ may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs use it, as the checker.

CPU restatements of the reference's hot path:

* ``logl_numpy``        numpy restatement of /root/reference/bild/src/MSRouse_logL_py.py:5-121
                        (dense ``B @ C @ B + Sig``; ``/S`` and ``log S`` as in _py.py:50)
* ``logl_c``            ctypes front-end of oracle/kalman_oracle.c (restates MSRouse_logL.pyx:19-256)
* ``st2states``         /root/reference/bild/amis.py:670-695 (``FixedkSampler.st2profile``)
* ``amis_evidence``     /root/reference/bild/amis.py:843-845, 878-900 (weights, evidence, sem, KL)
* ``logl_dense_gaussian`` an INDEPENDENT check that does not use the Kalman recursion at all: joint
                        Gaussian of all valid observations, evaluated with slogdet + solve.
* ``ref_cython``        loader for oracle/_ref/MSRouse_logL*.so - the reference's own .pyx compiled
                        from /root/reference by oracle/Makefile (absent => returns None)

Pinning: tests/test_oracle.py checks logl_c and logl_numpy against tests/golden/*.npz, which
oracle/make_golden.py produced by IMPORTING the unmodified reference from /root/reference
(pure-Python twin + compiled .pyx) - see that script.  The propagators fed to all of them come from
the third-party ``rouse`` package in the reference; that part is restated in oracle/rouse_oracle.py
and is "parity unpinned" (no golden value exists in the reference; see its header).
"""
import ctypes
import glob
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LOG_2PI = np.log(2 * np.pi)


# ------------------------------------------------------------------ input preparation (pyx:143-178)
def noise_to_s2_cind(localization_error):
    """pyx:144-147: unique errors -> s2 (d*,), Cind (d,)"""
    err = np.asarray(localization_error, dtype=float)
    uniq, cind = np.unique(err, return_inverse=True)
    return uniq ** 2, cind.astype(np.uint32)


def model_arrays(models):
    """pyx:152-160: stack per-state dynamics and steady states of rouse.Model-like objects."""
    for m in models:
        m.check_dynamics()
    Bs = np.ascontiguousarray([m._dynamics["B"] for m in models], dtype=float)
    Gs = np.ascontiguousarray([m._dynamics["G"] for m in models], dtype=float)
    Sigs = np.ascontiguousarray([m._dynamics["Sig"] for m in models], dtype=float)
    ss = [m.steady_state() for m in models]
    M0 = np.ascontiguousarray([s[0] for s in ss], dtype=float)
    C0 = np.ascontiguousarray([s[1] for s in ss], dtype=float)
    return Bs, Gs, Sigs, M0, C0


# ------------------------------------------------------------------ numpy restatement (_py.py)
def _kalman_update(w, x, M, C, s2, Cind):
    """_py.py:38-52"""
    xmm = x - w @ M
    Cw = C @ w
    S = Cw @ w + s2
    K = Cw / S[:, None]
    C = C - K[:, :, None] * Cw[:, None, :]
    M = M + K[Cind].T * xmm
    logL = -0.5 * (xmm * xmm / S[Cind] + np.log(S)[Cind] + LOG_2PI)
    return M, C, logL


def logl_numpy(Bs, Gs, Sigs, M0, C0, w, x, s2, Cind, states):
    """_py.py:54-121 with the propagators given as arrays."""
    states = np.asarray(states)
    x = np.asarray(x, dtype=float)
    Cind = np.asarray(Cind, dtype=np.intp)
    M = M0[states[0]].copy()
    C = np.tile(C0[states[0]], (len(s2), 1, 1))
    valid = ~np.any(np.isnan(x), axis=1)
    total = []
    if valid[0]:
        M, C, ll = _kalman_update(w, x[0], M, C, s2, Cind)
        total.append(ll)
    for t in range(1, len(states)):
        B, G, Sig = Bs[states[t]], Gs[states[t]], Sigs[states[t]]
        M = B @ M + G
        C = B @ C @ B + Sig
        if valid[t]:
            M, C, ll = _kalman_update(w, x[t], M, C, s2, Cind)
            total.append(ll)
    return float(np.sum(np.array(total))) if total else 0.0


# ------------------------------------------------------------------ C restatement (pyx)
_LIB = None


def _c_lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "_build", "libbild_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        lib = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.bild_oracle_logl_batch.restype = ctypes.c_int
        lib.bild_oracle_logl_batch.argtypes = [ctypes.c_int] * 6 + [dp] * 8 + [
            ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_int32), dp]
        _LIB = lib
    return _LIB


def logl_c(Bs, Gs, Sigs, M0, C0, w, x, s2, Cind, states):
    """states: (T,) or (P,T) int.  Returns float or (P,) array."""
    lib = _c_lib()
    states = np.ascontiguousarray(states, dtype=np.int32)
    single = states.ndim == 1
    states2 = states[None, :] if single else states
    P, T = states2.shape
    S, N, D = Gs.shape
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (Bs, Gs, Sigs, M0, C0, w, x, s2)]
    assert arrs[6].shape == (T, D)
    Cind = np.ascontiguousarray(Cind, dtype=np.uint32)
    out = np.empty(P, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    rc = lib.bild_oracle_logl_batch(
        N, D, len(arrs[7]), S, T, P, *[a.ctypes.data_as(dp) for a in arrs],
        Cind.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
        np.ascontiguousarray(states2).ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
        out.ctypes.data_as(dp))
    if rc != 0:
        raise ValueError("bild_oracle_logl_batch: bad arguments")
    return float(out[0]) if single else out


# ------------------------------------------------------------------ the reference .pyx, compiled
def ref_cython():
    """Return the reference's own compiled ``MSRouse_logL`` function or None if oracle/_ref is absent."""
    hits = sorted(glob.glob(os.path.join(HERE, "_ref", "MSRouse_logL*.so")))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("MSRouse_logL", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MSRouse_logL


# ------------------------------------------------------------------ profiles (amis.py:670-695)
def st2states(s, theta, T):
    """(s, theta) -> per-frame state array, exactly the numpy expressions of amis.py:685-693."""
    states = theta[0] * np.ones(T)
    if len(s) > 1:
        switchpos = np.cumsum(s)[:-1]
        switches = np.floor(switchpos * (T - 1)).astype(int) + 1
        for i in range(1, len(switches)):
            states[switches[i - 1]:switches[i]] = theta[i]
        states[switches[-1]:] = theta[-1]
    return states.astype(int)


# ------------------------------------------------------------------ weights / evidence (amis.py)
def amis_evidence(logLs, logdeltas, cur_log_proposal, n_steps, logprior):
    """amis.py:843-845 and 878-900 on the concatenated ensemble."""
    from scipy import stats
    log_weights = logLs - logdeltas + np.log(n_steps)
    mx = np.max(log_weights)
    with np.errstate(under="ignore"):
        wo = np.exp(log_weights - mx)
    ev_o = np.mean(wo)
    logev = np.log(ev_o) + mx + logprior
    dlogev = stats.sem(wo) / ev_o
    with np.errstate(under="ignore", invalid="ignore"):
        KL = np.nansum(wo * (logLs - cur_log_proposal)) / len(wo) / ev_o - logev + logprior
    return log_weights, logev, dlogev, KL


# ------------------------------------------------------------------ independent dense check
def logl_dense_gaussian(Bs, Gs, Sigs, M0, C0, w, x, s2, Cind, states):
    """
    Joint Gaussian of all valid observations, no Kalman recursion: for each spatial dimension j,
    y_t = w.x_t + eps,  x_0 ~ N(M0, C0),  x_t = B_t x_{t-1} + G_t + N(0, Sig_t).
    Cov(x_t, x_u) = P_t Phi(u,t)^T for t <= u with Phi(u,t) = B_u ... B_{t+1}.  O(T^2 N^2): small T only.
    """
    states = np.asarray(states)
    x = np.asarray(x, dtype=float)
    T, D = x.shape
    N = len(w)
    s0 = states[0]
    means = [M0[s0].copy()]
    covs = [C0[s0].copy()]
    for t in range(1, T):
        B = Bs[states[t]]
        means.append(B @ means[-1] + Gs[states[t]])
        covs.append(B @ covs[-1] @ B.T + Sigs[states[t]])
    valid = np.nonzero(~np.any(np.isnan(x), axis=1))[0]
    V = len(valid)
    if V == 0:
        return 0.0
    # cross-covariance of the projected process: K[a,b] = w^T Cov(x_ta, x_tb) w
    Kmat = np.empty((V, V))
    for a, ta in enumerate(valid):
        v = covs[ta] @ w            # Cov(x_ta, x_ta) w
        Kmat[a, a] = w @ v
        cur = v                      # Cov(x_u, x_ta) w for u = ta
        u = ta
        for b in range(a + 1, V):
            tb = valid[b]
            while u < tb:
                u += 1
                cur = Bs[states[u]] @ cur
            Kmat[a, b] = Kmat[b, a] = w @ cur
    total = 0.0
    for j in range(D):
        mu = np.array([w @ means[t][:, j] for t in valid])
        Kj = Kmat + s2[Cind[j]] * np.eye(V)
        r = x[valid, j] - mu
        sign, logdet = np.linalg.slogdet(Kj)
        total += -0.5 * (r @ np.linalg.solve(Kj, r) + logdet + V * LOG_2PI)
    return float(total)
