"""Shared helpers for the parity tests (oracle side only; never imported by bild_b200)."""
import glob
import os

import numpy as np

import kalman_oracle as ko
import rouse_oracle as ro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL_KEYS = ("Bs", "Gs", "Sigs", "M0", "C0", "w")


def golden_cases():
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "logl_*.npz")))


def load_golden(path):
    return dict(np.load(path))


def oracle_model(N, d=3, D=1.0, k=5.0, loops=(None, [(0, -1)]), w=None, F=None):
    """Model arrays from the ORACLE's rouse restatement."""
    models = []
    for lp in loops:
        m = ro.Model(N, D, k, d, add_bonds=lp, setup_dynamics=False)
        if F is not None:
            m.F[:] = F
        m.update_dynamics()
        models.append(m)
    Bs, Gs, Sigs, M0, C0 = ko.model_arrays(models)
    if w is None:
        w = np.zeros(N)
        w[0], w[-1] = -1.0, 1.0
    return dict(Bs=Bs, Gs=Gs, Sigs=Sigs, M0=M0, C0=C0, w=np.asarray(w, dtype=float), models=models)


def synth_traj(mod, T, rng, noise, p_nan=0.0, dwell=None):
    """Synthetic trajectory from the generative model (models.py:295-350 restated with a local RNG)."""
    S, N, d = mod["Gs"].shape
    dwell = dwell or max(2, T // 5)
    truth = (np.cumsum(rng.random(T) < 1.0 / dwell) % S).astype(int)
    noise = np.broadcast_to(np.asarray(noise, dtype=float), (d,))

    def sqrt_psd(X):
        lam, V = np.linalg.eigh(X)
        return V * np.sqrt(np.clip(lam, 0, None))

    Ls = [sqrt_psd(mod["Sigs"][s]) for s in range(S)]
    conf = mod["M0"][truth[0]] + sqrt_psd(mod["C0"][truth[0]]) @ rng.normal(size=(N, d))
    x = np.empty((T, d))
    x[0] = mod["w"] @ conf
    for t in range(1, T):
        s = truth[t]
        conf = mod["Bs"][s] @ conf + mod["Gs"][s] + Ls[s] @ rng.normal(size=(N, d))
        x[t] = mod["w"] @ conf
    x += noise * rng.normal(size=x.shape)
    if p_nan > 0:
        x[rng.random(T) < p_nan] = np.nan
    return x, truth


def random_profiles(rng, P, T, S, kmax=10):
    """AMIS-like (ss, thetas) with a common number of runs K1 = kmax+1 (k drawn per profile, padded with empty runs)."""
    K1 = kmax + 1
    ss = np.zeros((P, K1))
    thetas = np.zeros((P, K1), dtype=int)
    for p in range(P):
        k = rng.integers(0, kmax + 1)
        s = rng.dirichlet(np.ones(k + 1))
        ss[p, :k + 1] = s
        th = np.empty(K1, dtype=int)
        th[0] = rng.integers(S)
        for i in range(1, K1):
            th[i] = (th[i - 1] + 1 + rng.integers(max(1, S - 1))) % S
        thetas[p] = th
    return ss, thetas


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def logl_c_parallel(mod, x, s2, Cind, states, threads=None):
    """The C oracle over every host core: ctypes releases the GIL for the duration of the call, so plain threads
    scale (a fork pool is not an option inside a process that has initialised CUDA)."""
    import concurrent.futures as cf
    states = np.asarray(states)
    threads = threads or min(len(states), os.cpu_count() or 1)
    chunks = [c for c in np.array_split(np.arange(len(states)), threads) if len(c)]
    args = [mod[k] for k in MODEL_KEYS]
    with cf.ThreadPoolExecutor(len(chunks)) as ex:
        parts = list(ex.map(lambda c: ko.logl_c(*args, x, s2, Cind, states[c]), chunks))
    return np.concatenate([np.atleast_1d(p) for p in parts])
