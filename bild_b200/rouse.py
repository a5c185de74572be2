"""
Rouse-chain propagators, precomputed once per model (north-star item 1).

The reference takes these from the third-party ``rouse`` package (``rouse.Model(N, D, k, d,
add_bonds=...)``, /root/reference/bild/models.py:246) and re-reads ``_dynamics['B'|'G'|'Sig']`` and
``steady_state()`` on every likelihood call (/root/reference/bild/src/MSRouse_logL.pyx:152-160).
Here one symmetric eigendecomposition of the connectivity matrix yields every quantity in closed
form, and the stacked arrays are uploaded to the GPU once (bild_b200/engine.py).

Model: each spatial dimension is an N-dim Ornstein-Uhlenbeck process

    dx = -k A x dt + F dt + sqrt(2 D) dW ,      A = V diag(lam) V^T  (graph Laplacian + extra bonds)

so over one frame (dt = 1), mode by mode with r = k lam:

    B   = V diag(exp(-r dt))                         V^T
    Sig = V diag(D (1 - exp(-2 r dt)) / r)           V^T      (2 D dt on zero modes)
    G   = V diag((1 - exp(-r dt)) / r)               V^T F    (dt on zero modes)
    steady state:  M = V diag(1 / r) V^T F,  C = V diag(D / r) V^T   with zero modes projected out

The attribute / method names follow what the reference's in-tree code touches, so that a
``Model`` can be handed to reference code (and to the compiled reference .pyx) unchanged.
"""
import numpy as np

__all__ = ["Model", "connectivity"]

_ZERO_MODE_TOL = 1e-10


def connectivity(N, add_bonds=None):
    """Connectivity matrix of a free chain of N beads plus extra bonds ``(i, j[, rel_strength])``."""
    A = np.zeros((N, N))
    idx = np.arange(N - 1)
    A[idx, idx] += 1.0
    A[idx + 1, idx + 1] += 1.0
    A[idx, idx + 1] -= 1.0
    A[idx + 1, idx] -= 1.0
    for bond in (add_bonds or ()):
        i, j = int(bond[0]) % N, int(bond[1]) % N
        rel = float(bond[2]) if len(bond) > 2 else 1.0
        A[i, i] += rel
        A[j, j] += rel
        A[i, j] -= rel
        A[j, i] -= rel
    return A


class Model:
    """
    One Rouse chain (one state of a `MultiStateRouse`).

    Parameters
    ----------
    N : int
        number of monomers
    D, k : float
        monomer diffusivity and backbone spring constant
    d : int
        spatial dimension
    add_bonds : None or list of ``(i, j)`` / ``(i, j, rel_strength)``
        extra bonds; negative indices count from the end (the default loop is ``(0, -1)``)
    """

    def __init__(self, N, D=1.0, k=1.0, d=3, setup_dynamics=True, add_bonds=None):
        self.N, self.D, self.k, self.d = int(N), float(D), float(k), int(d)
        self.F = np.zeros((self.N, self.d))
        self.A = connectivity(self.N, add_bonds)
        self._dynamics = {"needs_updating": True}
        self._modes = None
        if setup_dynamics:
            self.update_dynamics()

    # ------------------------------------------------------------------ precompute
    def update_dynamics(self, dt=1.0):
        lam, V = np.linalg.eigh(self.A)
        zero = np.abs(lam) < _ZERO_MODE_TOL * max(1.0, np.max(np.abs(lam)))
        r = np.where(zero, 1.0, self.k * lam)
        decay = np.where(zero, 1.0, np.exp(-r * dt))
        sig = np.where(zero, 2.0 * self.D * dt, -self.D * np.expm1(-2.0 * r * dt) / r)
        gfac = np.where(zero, dt, -np.expm1(-r * dt) / r)
        rinv = np.where(zero, 0.0, 1.0 / r)

        def spectral(f):
            X = (V * f) @ V.T
            return np.ascontiguousarray(0.5 * (X + X.T))

        self._modes = (lam, V, zero, sig, rinv)
        self._dynamics = {
            "needs_updating": False, "N": self.N, "D": self.D, "k": self.k, "dt": dt,
            "B": spectral(decay),
            "G": np.ascontiguousarray(spectral(gfac) @ self.F),
            "Sig": spectral(sig),
        }
        self._ss = (np.ascontiguousarray(spectral(rinv) @ self.F), spectral(self.D * rinv))

    def check_dynamics(self, dt=1.0, run_if_necessary=True):
        d = self._dynamics
        stale = (d.get("needs_updating", True) or d["N"] != self.N or d["D"] != self.D
                 or d["k"] != self.k or d["dt"] != dt)
        if stale:
            if not run_if_necessary:
                raise RuntimeError("Model changed since last call to update_dynamics()")
            self.update_dynamics(dt)

    # ------------------------------------------------------------------ ensemble moments
    def steady_state(self):
        """``(M (N, d), C (N, N))`` of the stationary ensemble (centre of mass pinned at 0)."""
        self.check_dynamics()
        return self._ss[0].copy(), self._ss[1].copy()

    def propagate_M(self, M, dt=1.0, check_dynamics=True):
        if check_dynamics:
            self.check_dynamics(dt)
        return self._dynamics["B"] @ M + self._dynamics["G"]

    def propagate_C(self, C, dt=1.0, check_dynamics=True):
        if check_dynamics:
            self.check_dynamics(dt)
        B = self._dynamics["B"]
        return B @ C @ B + self._dynamics["Sig"]

    def propagate(self, M, C, dt=1.0, check_dynamics=True):
        return self.propagate_M(M, dt, check_dynamics), self.propagate_C(C, dt, check_dynamics)

    # ------------------------------------------------------------------ generative sampling
    def conf_ss(self):
        """Draw one conformation (N, d) from the steady state (uses the global numpy RNG)."""
        self.check_dynamics()
        lam, V, zero, sig, rinv = self._modes
        z = np.random.normal(size=(self.N, self.d))
        return self._ss[0] + (V * np.sqrt(self.D * rinv)) @ z

    def evolve(self, conf, dt=1.0):
        """Propagate a conformation by one frame, with thermal noise (global numpy RNG)."""
        self.check_dynamics(dt)
        lam, V, zero, sig, rinv = self._modes
        z = np.random.normal(size=conf.shape)
        return self._dynamics["B"] @ conf + self._dynamics["G"] + (V * np.sqrt(sig)) @ z
