"""
ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the Rouse-chain propagators that the reference obtains from the third-party
package ``rouse`` (PyPI / github OpenTrajectoryAnalysis/rouse; the reference pins it as
``rouse >= 0`` in /root/reference/pyproject.toml:22, i.e. UNPINNED, and its sources are not under
/root/reference).  PARITY UNPINNED for this file: there is no golden value in the reference's tests
for B, G, Sig or the steady state, so what is restated here is the published physics that the
reference documents at /root/reference/bild/models.py:163-221 and consumes at

    models.py:246          rouse.Model(N, D, k, d, add_bonds=loop)
    MSRouse_logL.pyx:152-160   m.check_dynamics(); m._dynamics['B'|'G'|'Sig']; m.steady_state()
    MSRouse_logL_py.py:109-110 m.propagate_M(M, check_dynamics=False); m.propagate_C(C, ...)
    models.py:332, 337     m.conf_ss(); m.evolve(conf)
    models.py:366          m.steady_state()

Physics: N beads, each spatial dimension an independent N-dim Ornstein-Uhlenbeck process
    dx = -k A x dt + F dt + sqrt(2 D) dW,          dt = 1 frame,
with A the free-chain graph Laplacian plus rel*(e_i - e_j)(e_i - e_j)^T per extra bond.  Then
    B   = exp(-k A)
    Sig = int_0^1 exp(-k A s) 2D exp(-k A^T s) ds
    G   = int_0^1 exp(-k A s) ds F
    steady state  M = (1/k) A^+ F,   C = (D/k) A^+      (centre-of-mass mode projected out)

To be INDEPENDENT of the product implementation (bild_b200/rouse.py uses one symmetric
eigendecomposition), this file deliberately uses different numerics: scipy.linalg.expm (Pade) for
the exponentials and an SVD pseudo-inverse for the integrals and the steady state.
"""
import numpy as np
from scipy import linalg


def connectivity(N, add_bonds=None):
    """Graph Laplacian of the free chain plus extra bonds [(i, j[, rel_strength])...]."""
    A = 2.0 * np.eye(N) - np.eye(N, k=1) - np.eye(N, k=-1)
    A[0, 0] = A[-1, -1] = 1.0
    if add_bonds is not None:
        for bond in add_bonds:
            i, j = int(bond[0]), int(bond[1])
            rel = float(bond[2]) if len(bond) > 2 else 1.0
            A[i, i] += rel
            A[j, j] += rel
            A[i, j] -= rel
            A[j, i] -= rel
    return A


def dynamics(A, D, k, F, dt=1.0):
    """(B, G, Sig) from Pade matrix exponentials and an SVD pseudo-inverse (A is symmetric, so it
    commutes with its exponential):
        int_0^dt exp(-X s) ds = X^+ (I - exp(-X dt)) + dt P0,   P0 = I - A A^+ (null-space projector).
    (Van Loan's block exponential was tried first and rejected: for k*lam_max*dt ~ 20 the block
    matrix carries exp(+20) terms and loses ~8 digits.)"""
    N = A.shape[0]
    I = np.eye(N)
    Ap = np.linalg.pinv(A, rcond=1e-10)
    P0 = I - A @ Ap
    B = linalg.expm(-k * dt * A)
    B2 = linalg.expm(-2.0 * k * dt * A)
    Sig = (D / k) * Ap @ (I - B2) + 2.0 * D * dt * P0
    G = ((1.0 / k) * Ap @ (I - B) + dt * P0) @ F
    B = 0.5 * (B + B.T)
    Sig = 0.5 * (Sig + Sig.T)
    return np.ascontiguousarray(B), np.ascontiguousarray(G), np.ascontiguousarray(Sig)


def steady_state(A, D, k, F):
    Ap = np.linalg.pinv(A, rcond=1e-10, hermitian=False)
    Ap = 0.5 * (Ap + Ap.T)
    return np.ascontiguousarray(Ap @ F / k), np.ascontiguousarray(D / k * Ap)


class Model:
    """Duck-typed stand-in for ``rouse.Model`` exposing exactly what the reference touches."""

    def __init__(self, N, D=1.0, k=1.0, d=3, setup_dynamics=True, add_bonds=None):
        self.N, self.D, self.k, self.d = N, float(D), float(k), d
        self.F = np.zeros((N, d))
        self.A = connectivity(N, add_bonds)
        self._dynamics = {"needs_updating": True}
        if setup_dynamics:
            self.update_dynamics()

    def update_dynamics(self, dt=1.0):
        B, G, Sig = dynamics(self.A, self.D, self.k, self.F, dt)
        self._dynamics = {"needs_updating": False, "N": self.N, "D": self.D, "k": self.k,
                          "dt": dt, "B": B, "G": G, "Sig": Sig}
        self._ss = steady_state(self.A, self.D, self.k, self.F)

    def check_dynamics(self, dt=1.0, run_if_necessary=True):
        if self._dynamics.get("needs_updating", True):
            if not run_if_necessary:
                raise RuntimeError("dynamics out of date")
            self.update_dynamics(dt)

    def steady_state(self):
        self.check_dynamics()
        M, C = self._ss
        return M.copy(), C.copy()

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

    def _sqrt_psd(self, X):
        lam, V = np.linalg.eigh(X)
        return V * np.sqrt(np.clip(lam, 0.0, None))

    def conf_ss(self):
        M, C = self.steady_state()
        return M + self._sqrt_psd(C) @ np.random.normal(size=(self.N, self.d))

    def evolve(self, conf, dt=1.0):
        self.check_dynamics(dt)
        L = self._sqrt_psd(self._dynamics["Sig"])
        return self._dynamics["B"] @ conf + self._dynamics["G"] + L @ np.random.normal(size=conf.shape)


def twoLocusMSD(dt, Gamma, J):
    """Two-locus Rouse MSD (only referenced by the out-of-scope GenericGaussianModel)."""
    from scipy.special import erfc
    dt = np.asarray(dt, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore", under="ignore"):
        tau = (J / Gamma) ** 2 / np.pi
        out = 2 * Gamma * np.sqrt(dt) * (1 - np.exp(-tau / dt)) + 2 * J * erfc(np.sqrt(tau / dt))
    return np.where(dt == 0, 0.0, out)
