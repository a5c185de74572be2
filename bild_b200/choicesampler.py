"""
Which k to sample next?  Information-based sample selection for the iterative AMIS scheme.

Behavioural mirror of /root/reference/bild/choicesampler.py (`ChoiceSampler`); host-side numpy, tiny
(``samplesize x kmax`` normals), but it draws from the global numpy RNG (choicesampler.py:106), so the
call order is part of the seed-parity contract and is kept.

Provenance note: this class is OUT OF SCOPE of the accelerated path and exists only so that
``import bild_b200 as bild`` is a complete drop-in.  It follows the reference class closely - same attribute and
method names (they are public API: ``n0``, ``samplesize``, ``KLD_moreSamples``, ``KLD_omitK``) and, necessarily, the
same formulas and the same order of random draws - i.e. it is a restatement of choicesampler.py:83-210, not
independent work, and claims no credit.
"""
import numpy as np

__all__ = ["ChoiceSampler"]


def _native():
    """Host-side helpers of libbild_b200.so available?  (Same switch as the AMIS layer: without the built library the
    numpy statements below are used - this is k-selection bookkeeping, not the likelihood.)"""
    from .amis import _native_helpers
    return _native_helpers()


class ChoiceSampler:
    """
    Monte-Carlo "choice distribution" p(k): how often is k the smallest k whose evidence lies within ``dE``
    of the maximum, when the evidence curve is resampled within its error bars?

    Parameters
    ----------
    muhat, shat : (k,) arrays
        evidence estimates and the variances of these estimates
    N : (k,) array (float, ``inf`` allowed)
        number of AMIS steps behind each estimate
    dE : float
        evidence margin
    samplesize : int
        Monte-Carlo sample size
    """

    def __init__(self, muhat, shat, N, dE, samplesize=10000):
        self.dE, self.muhat, self.shat, self.N, self.samplesize = dE, muhat, shat, N, samplesize
        self.kmax = len(muhat)
        self.EDmu2 = self.shat / (self.N + 1)       # expected squared move of the estimate after one more step
        self.Dmu = np.sqrt(self.EDmu2)
        self.init_sample()

    def init_sample(self):
        """(Re)draw the common random numbers all evaluations share, and the point estimate of p(k)."""
        self._scaled_rvs = np.sqrt(self.shat[None, ...]) * np.random.normal(size=(self.samplesize, self.kmax))
        self.bestk = self.evaluate()
        self.n0 = np.bincount(self.bestk, minlength=self.kmax)       # histogram of bestk (= column sums of `best_is_k`)

    @property
    def best_is_k(self):
        """``(samplesize, k)`` truth table of `bestk` (documented attribute of the reference class; built on demand)."""
        return self.bestk[:, None] == np.arange(self.kmax)[None, :]

    def evaluate(self, k_change=None, n_step=0, omit_k=None):
        """Chosen k per Monte-Carlo draw, optionally with ``muhat[k_change]`` shifted by ``n_step * Dmu`` or
        with some k ignored (``omit_k``)."""
        mu = self.muhat.copy()
        if k_change is not None:
            mu[k_change] += n_step * self.Dmu[k_change]
        if omit_k is not None:
            mu[omit_k] = np.nan
        if _native():      # one pass over the draws in libbild_b200.so (bildk_choice_pick): same additions, same comparisons
            from . import _lib
            picks = np.empty(self.samplesize, dtype=np.int64)
            _lib.check(_lib.load().bildk_choice_pick(self.samplesize, self.kmax, _lib.ptr(self._scaled_rvs, _lib.c_double_p),
                                                     _lib.ptr(np.ascontiguousarray(mu, dtype=np.float64), _lib.c_double_p), float(self.dE),
                                                     _lib.ptr(picks, _lib.c_int64_p)))
            return picks
        return self._evaluate_numpy(mu)

    def _evaluate_numpy(self, mu):
        draws = self._scaled_rvs + mu
        top = np.nanmax(draws, axis=1, keepdims=True)
        return np.nanargmax(top - self.dE - draws <= 0, axis=1)      # first k within dE of the maximum

    def Dn(self):
        """``[k1, k2]``: expected change of the count for k2 caused by one more sample at k1."""
        if _native():      # the 2 kmax evaluations below in ONE pass over the draws (bildk_choice_dn): 28 -> 1.5 ms at kmax = 11
            from . import _lib
            dn = np.empty((self.kmax, self.kmax), dtype=np.int64)
            _lib.check(_lib.load().bildk_choice_dn(self.samplesize, self.kmax, _lib.ptr(self._scaled_rvs, _lib.c_double_p),
                                                   _lib.ptr(np.ascontiguousarray(self.muhat, dtype=np.float64), _lib.c_double_p),
                                                   _lib.ptr(np.ascontiguousarray(self.Dmu, dtype=np.float64), _lib.c_double_p),
                                                   float(self.dE), _lib.ptr(dn, _lib.c_int64_p)))
            return dn
        return self._Dn_numpy()

    def _Dn_numpy(self):
        counts = []
        for step in (-0.5, 0.5):
            picks = []
            for k in range(self.kmax):
                mu = self.muhat.copy()
                mu[k] += step * self.Dmu[k]
                picks.append(self._evaluate_numpy(mu))
            picks = np.array(picks)                                                            # (k_change, samp)
            counts.append(np.sum(picks[..., None] == np.arange(self.kmax), axis=-2))          # (k_change, k)
        return counts[1] - counts[0]

    def KLD_moreSamples(self):
        """Expected Kullback-Leibler gain of one more AMIS step at each k."""
        Dn = self.Dn()
        return 0.5 / self.samplesize * np.sum(Dn ** 2 / (self.n0 + 1)[None, :], axis=-1)

    def KLD_omitK(self, omit_k=None):
        """Information contributed by the positions ``omit_k``: KL(full choice distribution || without them)."""
        without = self.evaluate(omit_k=omit_k)
        old_n = np.bincount(without, minlength=self.kmax)
        old_n = old_n / np.sum(old_n) * self.samplesize
        Dn = self.n0 - old_n
        Dn[omit_k] = 0          # would contribute infinite KLD (old_n is 0 there); not of interest
        return 0.5 / self.samplesize * np.sum(Dn ** 2 / (old_n + 1))
