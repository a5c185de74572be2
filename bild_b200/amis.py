"""
AMIS at a fixed number of switches, with the likelihood batch on the GPU.

Behavioural mirror of /root/reference/bild/amis.py (`Dirichlet`, `CFC`, `FixedkSampler`), written so
that a run under a fixed numpy seed consumes the global RNG in exactly the reference's order
(``stats.dirichlet(a).rvs`` amis.py:81 -> ``np.random.choice`` :247 -> ``np.random.rand`` per slot :254)
and therefore proposes the very same profile batches.  What changed:

* `FixedkSampler.logL` (amis.py:717-739) hands the whole batch to ``model.logL_st_batch`` when the model
  has one (the sm_100a engine) instead of looping over profiles in Python;
* the weight normalisation of `FixedkSampler.step` (amis.py:843-845, 878-900) runs as one deterministic
  device reduction when the model offers ``amis_weights``;
* proposal densities are closed-form and vectorised over the whole ensemble (the reference rebuilds
  scipy distribution objects per call, 18 % of ``bild.sample`` at P=100 and quadratic in the number of
  steps - SURVEY.md 8(f) rank 1).
"""
import itertools
import math

import numpy as np
from scipy import stats
from scipy.special import gammaln, xlogy

from .util import Loopingprofile

__all__ = ["Dirichlet", "CFC", "FixedkSampler", "drive"]


def _lse(a, axis=None, mask=None, keepdims=False):
    """log(sum(exp(a))) over ``axis`` restricted to ``mask``; -inf for empty sums; no warnings on underflow."""
    a = np.asarray(a, dtype=float)
    if mask is not None:
        a = np.where(mask, a, -np.inf)
    with np.errstate(under="ignore", invalid="ignore", divide="ignore"):
        amax = np.max(a, axis=axis, keepdims=True)
        amax = np.where(np.isfinite(amax), amax, 0.0)
        out = np.log(np.sum(np.exp(a - amax), axis=axis, keepdims=True)) + amax
    if not keepdims:
        out = np.squeeze(out, axis=axis) if axis is not None else out.reshape(())
    return out


_NATIVE = None


def _native_helpers():
    """True when libbild_b200.so (which also carries the host-side proposal-density helper) can be loaded."""
    global _NATIVE
    if _NATIVE is None:
        from . import _lib
        try:
            _lib.load()
            _NATIVE = True
        except (_lib.BildkError, OSError):
            _NATIVE = False
    return _NATIVE


class LikelihoodRequest(tuple):
    """The ``(ss, thetas)`` batch a sampler yields.  ``amis`` (optional): the sampler's pending device bookkeeping of this
    very batch (`bild_b200.engine.FusedAmisStep`).  A driver that launches asynchronously may let it ride on the
    likelihood launch (`bild_b200.dataset.sample_many`); one that does not simply answers with the likelihoods and the
    sampler runs the step itself."""
    amis = None

    def __new__(cls, ss, thetas, amis=None):
        self = super().__new__(cls, (ss, thetas))
        self.amis = amis
        return self


def drive(gen, evaluate):
    """
    Run a likelihood-requesting generator to completion: every ``(ss, thetas)`` it yields is answered with
    ``evaluate(ss, thetas)``; returns the generator's return value.  The AMIS code is written as generators that
    *yield* their likelihood batches, so that one driver can serve many trajectories from one fused GPU launch
    (`bild_b200.dataset.sample_many`) without threads; the public, synchronous methods drive them with this.
    """
    try:
        request = next(gen)
        while True:
            request = gen.send(evaluate(*request))
    except StopIteration as stop:
        return stop.value


# ------------------------------------------------------------------------------------------------ Dirichlet
class Dirichlet:
    """Dirichlet proposal over the interval lengths ``s`` (sampling, density, weighted method-of-moments fit)."""

    def sample(self, a, N=1):
        # scipy's frozen dirichlet draws with the global RandomState's `dirichlet`; calling it directly
        # consumes the identical stream (amis.py:81) without constructing a distribution object
        return np.random.dirichlet(np.asarray(a, dtype=float), size=N)

    def logpdf(self, a, ss):
        """
        ``(N,)`` log-densities.  Rows that scipy would reject (an ``s_i == 0`` whose ``a_i < 1``, entries
        outside [0, 1], not summing to one) evaluate to ``+inf`` - the reference's convention
        (amis.py:98-108), which zeroes the importance weight of such a sample.
        """
        a = np.asarray(a, dtype=float)
        ss = np.asarray(ss, dtype=float)
        single = ss.ndim == 1
        if single:
            ss = ss[None, :]
        lognorm = gammaln(np.sum(a)) - np.sum(gammaln(a))
        with np.errstate(divide="ignore", invalid="ignore"):
            out = lognorm + np.sum(xlogy(a - 1.0, ss), axis=1)
        bad = (np.any(ss < 0, axis=1) | np.any(ss > 1, axis=1) | np.any((ss == 0) & (a < 1)[None, :], axis=1)
               | (np.abs(np.sum(ss, axis=1) - 1.0) > 1e-9))
        out = np.where(bad, np.inf, out)
        return out[0] if single else out

    def logpdf_multi(self, A, ss):
        """``(n_par, N)`` log-densities of the samples ``ss (N, k+1)`` under every row of ``A (n_par, k+1)``."""
        A = np.asarray(A, dtype=float)
        ss = np.asarray(ss, dtype=float)
        if np.any(ss <= 0) or np.any(ss > 1) or np.any(np.abs(np.sum(ss, axis=1) - 1.0) > 1e-9):
            return np.array([self.logpdf(a, ss) for a in A])        # boundary conventions live in `logpdf`
        lognorm = gammaln(np.sum(A, axis=1)) - np.sum(gammaln(A), axis=1)
        return lognorm[:, None] + (A - 1.0) @ np.log(ss).T

    def estimate(self, ss, log_weights):
        """Weighted method of moments (amis.py:137-151): ``alpha = A m`` with ``A = mean(m (1-m) / v) - 1``."""
        with np.errstate(under="ignore"):
            w = np.exp(log_weights - np.max(log_weights))
            w /= np.sum(w)
            m = w @ ss
            v = w @ (ss - m[None, :]) ** 2
        return self.estimate_from_moments(m, v)

    @staticmethod
    def estimate_from_moments(m, v):
        """``alpha`` from the weighted mean ``m`` and variance ``v`` of the interval lengths (amis.py:145-151)."""
        if np.any(v == 0):
            A = 1e10   # degenerate ensemble: very concentrated but finite; the concentration brake takes over
        else:
            A = np.mean(m * (1 - m) / v) - 1
        return A * m


# ------------------------------------------------------------------------------------------------ CFC
class CFC:
    """
    Conflict-free categorical: proposal over state traces ``theta`` in which consecutive entries must be
    an allowed transition.  Parametrised by log-weights ``logp`` of shape ``(n, k+1)``; sampling is causal
    (slot i is drawn from column i restricted to the states reachable from theta[i-1]).
    """

    def __init__(self, transitions):
        self.transitions = np.array(transitions, dtype=bool, copy=True)
        self.MOM_maxiter = 1000
        self.MOM_precision = 1e-2

    @property
    def n(self):
        return self.transitions.shape[0]

    def sample(self, logp, N=1):
        k = logp.shape[1] - 1
        assert k >= 0
        with np.errstate(under="ignore"):
            p = np.exp(logp - _lse(logp, axis=0, keepdims=True))
        thetas = np.empty((N, k + 1), dtype=int)
        thetas[:, 0] = np.random.choice(self.n, size=N, p=p[:, 0])
        for i in range(1, k + 1):
            cdf = np.cumsum(p[None, :, i] * self.transitions[thetas[:, i - 1]], axis=1)
            cdf /= cdf[:, [-1]]
            thetas[:, i] = np.argmax(cdf > np.random.rand(N, 1), axis=1)   # first state whose cdf exceeds u
        return thetas

    def logpmf(self, logp, thetas):
        thetas = np.asarray(thetas)
        slots = np.arange(logp.shape[1])
        picked = logp[thetas, slots[None, :]]                               # (N, k+1)
        # normaliser of slot i given the previous state: LSE over the states reachable from it
        reach = _lse(logp.T[None, 1:, :], axis=-1, mask=self.transitions[thetas[:, :-1]])   # (N, k)
        return np.sum(picked, axis=1) - np.sum(reach, axis=1) - _lse(logp[:, 0])

    def logpmf_multi(self, logps, thetas):
        """``(n_par, N)`` log-probabilities of the traces ``thetas (N, k+1)`` under every ``logps[j] (n, k+1)``."""
        L = np.asarray(logps, dtype=float)                                   # (n_par, n, k+1)
        thetas = np.asarray(thetas)
        slots = np.arange(L.shape[2])
        picked = L[:, thetas, slots[None, :]]                                # (n_par, N, k+1)
        out = np.sum(picked, axis=2) - _lse(L[:, :, 0], axis=1)[:, None]
        if L.shape[2] > 1:
            mask = self.transitions[thetas[:, :-1]]                          # (N, k, n) states reachable from the previous one
            reach = _lse(L.transpose(0, 2, 1)[:, None, 1:, :], axis=-1, mask=mask[None])   # (n_par, N, k)
            out = out - np.sum(reach, axis=2)
        return out

    def estimate(self, thetas, log_weights):
        """Method of marginals (amis.py:283-305): weighted state marginals per slot -> weight parameters."""
        onehot = thetas[None, :, :] == np.arange(self.n)[:, None, None]     # (n, N, k+1)
        log_marginals = _lse(np.broadcast_to(log_weights[None, :, None], onehot.shape), axis=1, mask=onehot)
        return self.estimate_from_log_marginals(log_marginals)

    def estimate_from_log_marginals(self, log_marginals):
        """Weight parameters from the (unnormalised) per-slot log marginals ``(n, k+1)`` (amis.py:302-305)."""
        log_marginals = log_marginals - _lse(log_marginals, axis=0, keepdims=True)
        return self.logp_from_marginals(log_marginals)

    def logp_from_marginals(self, log_marginals):
        k = log_marginals.shape[1] - 1
        assert k >= 0
        logp = np.empty(log_marginals.shape, dtype=float)
        logp[:, 0] = log_marginals[:, 0]
        for i in range(1, k + 1):
            logp[:, i] = self.solve_marginals_single(log_marginals[:, i], log_marginals[:, i - 1])
        return logp

    def solve_marginals_single(self, logf, logg):
        """
        Fixed point of ``p_n = f_n / sum_{m -> n} g_m / sum_{m -> j} p_j`` (amis.py:336-392), iterated from
        ``p = f`` until successive iterates differ by less than ``MOM_precision``.
        """
        if np.any(logf == 0):                      # Kronecker-delta marginal
            return logf.copy()
        if np.any(logg == 0):
            assert np.all(logf[logg == 0] == -np.inf)
            return logf.copy()
        f_zero = logf == -np.inf
        g_zero = logg == -np.inf
        cur = logf
        for _ in range(self.MOM_maxiter):
            norm_from = _lse(cur[None, :], axis=1, mask=self.transitions)          # from state m: LSE over targets
            norm_from = np.where(g_zero, 0.0, norm_from)
            inflow = _lse((logg - norm_from)[:, None], axis=0, mask=self.transitions)   # into state i
            inflow = np.where(f_zero, 0.0, inflow)
            new = logf - inflow
            new = new - _lse(new)
            if np.max(np.abs(new[~f_zero] - cur[~f_zero])) < self.MOM_precision:
                return new
            cur = new
        raise RuntimeError("Iteration did not converge")

    def _int_transitions(self):
        return self.transitions.astype(int).astype(object)   # python ints: no overflow for long traces

    def uniform_marginals(self, k):
        """Slot marginals of the uniform distribution over allowed traces, by path counting in big integers."""
        Tm = self._int_transitions()
        powers = [np.linalg.matrix_power(Tm, i) for i in range(k + 1)]
        counts = np.empty((self.n, k + 1), dtype=object)
        for i in range(k + 1):
            counts[:, i] = powers[i].sum(axis=0) * powers[k - i].sum(axis=1)

        def biglog(x):
            return math.log(x) if x > 0 else -np.inf

        total = counts.sum(axis=0)
        out = np.empty((self.n, k + 1), dtype=float)
        for s in range(self.n):
            for i in range(k + 1):
                out[s, i] = biglog(counts[s, i]) - biglog(total[i])
        return out

    def logp_uniform(self, k):
        return self.logp_from_marginals(self.uniform_marginals(k))

    def N_total(self, k, log=False):
        total = np.sum(np.linalg.matrix_power(self._int_transitions(), k))
        return math.log(total) if log else total

    def full_sample(self, k, Nmax=1000):
        """All allowed traces with ``k`` switches, in the reference's enumeration order (first slot slowest)."""
        total = self.N_total(k)
        if total > Nmax:
            raise ValueError(f"Full sample would be {total} > Nmax = {Nmax} traces")
        successors = [np.nonzero(row)[0].tolist() for row in self.transitions]
        traces = [[s] for s in range(self.n)]
        for _ in range(k):
            traces = [tr + [nxt] for tr in traces for nxt in successors[tr[-1]]]
        return np.array(traces, dtype=int).reshape(len(traces), k + 1)


# ------------------------------------------------------------------------------------------------ sampler
class FixedkSampler:
    """
    AMIS (Cornuet et al. 2012) over profiles with exactly ``k`` switches; same constructor, attributes
    and method names as the reference (amis.py:540-972).  ``step()`` draws ``N`` profiles from the current
    Dirichlet x CFC proposal, evaluates their likelihood in one batch, recomputes the deterministic-mixture
    weights of the whole ensemble, refits the proposal (with the concentration / polarisation brakes) and
    appends ``(log evidence, its standard error, KL)`` to ``evidences``.
    """

    class ExhaustionImpractical(ValueError):
        pass

    def __init__(self, traj, model, k, N=100, concentration_brake=1e-2, polarization_brake=1e-3,
                 max_fev=20000, max_fcomplete=1000, _defer=False):
        # _defer (internal): do not evaluate the exhaustive sample here; the caller runs `start_gen()` itself
        self.k, self.N = k, N
        self.brakes = (concentration_brake, polarization_brake)
        self.max_fev, self.max_fcomplete = max_fev, max_fcomplete
        self.exhausted = False
        self.traj, self.model = traj, model

        self._started = False
        if self.k >= len(self.traj):      # more switches than frames: unidentifiable by construction
            self.evidences = [(-np.inf, 1e-10, np.inf)]
            self.exhausted = True
            self._started = True
            return

        self.dirichlet = Dirichlet()
        self.cfc = CFC(model.transitions)
        self.parameters = [(np.ones(self.k + 1), self.cfc.logp_uniform(self.k))]
        # uniform prior over profiles: k! / N_total(k)  (simplex volume 1/k! x number of traces)
        self.logprior = np.sum(np.log(np.arange(self.k) + 1)) - self.cfc.N_total(self.k, log=True)
        self.samples = []       # dicts with 'ss', 'thetas', 'logLs' [, 'logδs', 'log_weights', 'cur_log_proposal']
        self.evidences = []     # (logev, dlogev, KL) per step
        if not _defer:
            drive(self.start_gen(), self.logL)

    def start_gen(self):
        """Generator form of the construction-time work: the exhaustive evaluation if the space is small enough."""
        if self._started:
            return
        self._started = True
        try:
            yield from self.fix_exhaustive_gen()
        except FixedkSampler.ExhaustionImpractical:
            pass

    # ------------------------------------------------------------------ profiles
    def st2profile(self, s, theta):
        """(s, theta) -> `Loopingprofile`; frame f gets the state of the interval containing it, switch i at
        ``floor(cumsum(s)[i] * (T-1)) + 1`` (amis.py:685-693)."""
        T = len(self.traj)
        states = theta[0] * np.ones(T)
        if len(s) > 1:
            switches = np.floor(np.cumsum(s)[:-1] * (T - 1)).astype(int) + 1
            for i in range(1, len(switches)):
                states[switches[i - 1]:switches[i]] = theta[i]
            states[switches[-1]:] = theta[-1]
        return Loopingprofile(states)

    def log_proposal(self, parameters, ss, thetas):
        return self.log_proposal_multi([parameters], ss, thetas)[0]

    def log_proposal_multi(self, parameters, ss, thetas):
        """
        ``(len(parameters), N)``: the samples under every proposal of the list in one pass of the native helper
        ``bildk_amis_log_proposal`` (the reference evaluates proposals one by one through scipy distribution
        objects, amis.py:697-715, 836-839 - O(steps) constructions per step).  `Dirichlet.logpdf_multi` +
        `CFC.logpmf_multi` are the numpy statement of the same arithmetic (tests compare the two).
        """
        from . import _lib
        n_par, n = len(parameters), len(ss)
        out = np.empty((n_par, n))
        if n_par == 0 or n == 0:
            return out
        A = np.ascontiguousarray([par[0] for par in parameters], dtype=np.float64)
        L = np.ascontiguousarray([par[1] for par in parameters], dtype=np.float64)
        if not _native_helpers():
            # host-side bookkeeping only (no likelihood is evaluated here): without the built library - pure-CPU models
            # such as `FactorizedModel` on a machine without nvcc - use the numpy statement of the same arithmetic,
            # as the reference's AMIS layer (pure numpy/scipy) does.  The LIKELIHOOD has no such fallback.
            return self.dirichlet.logpdf_multi(A, ss) + self.cfc.logpmf_multi(L, thetas)
        ss = np.ascontiguousarray(ss, dtype=np.float64)
        thetas = np.ascontiguousarray(thetas, dtype=np.int64)
        trans = np.ascontiguousarray(self.cfc.transitions, dtype=np.uint8)
        _lib.check(_lib.load().bildk_amis_log_proposal(
            n_par, n, ss.shape[1], self.cfc.n, _lib.ptr(A, _lib.c_double_p), _lib.ptr(L, _lib.c_double_p),
            _lib.ptr(trans, _lib.c_uint8_p), _lib.ptr(ss, _lib.c_double_p),
            thetas.ctypes.data_as(_lib.ctypes.POINTER(_lib.ctypes.c_int64)), _lib.ptr(out, _lib.c_double_p)))
        return out

    def logL(self, ss, thetas):
        """Model likelihood of a batch of samples -> (N,) float64.  One GPU launch when the model can."""
        if hasattr(self.model, "logL_st_batch"):
            return np.asarray(self.model.logL_st_batch(ss, thetas, self.traj), dtype=float)
        if hasattr(self.model, "logL_st"):
            return np.array([self.model.logL_st(s, theta, self.traj) for s, theta in zip(ss, thetas)])
        return np.array([self.model.logL(self.st2profile(s, theta), self.traj) for s, theta in zip(ss, thetas)])

    # ------------------------------------------------------------------ exhaustive evaluation for tiny spaces
    def fix_exhaustive(self):
        return drive(self.fix_exhaustive_gen(), self.logL)

    def fix_exhaustive_gen(self):
        """
        If there are at most ``min(max_fcomplete, max_fev)`` profiles with ``k`` switches, evaluate all of
        them (one batch) and compute the evidence exactly as the mean likelihood under the uniform prior;
        otherwise raise `ExhaustionImpractical` (amis.py:741-803).
        """
        Nmax = min(self.max_fcomplete, self.max_fev)
        T = len(self.traj)
        count = self.cfc.N_total(self.k)
        for i in range(self.k):
            count *= T - i - 1
            if count > Nmax:
                raise self.ExhaustionImpractical(
                    f"Parameter space too large for exhaustive sampling (number of profiles = {count} > Nmax = {Nmax})")

        grid = np.array(list(itertools.combinations(np.arange(T - 1) + 0.5, self.k))) / (T - 1)   # switch positions
        grid = np.append(np.insert(grid, 0, 0, axis=1), np.ones((len(grid), 1)), axis=1)
        ss = np.diff(grid, axis=1)
        thetas = self.cfc.full_sample(self.k, Nmax=Nmax)
        n_ss = len(ss)
        ss = np.tile(ss, (len(thetas), 1))
        thetas = np.repeat(thetas, n_ss, axis=0)

        sample = {"ss": ss, "thetas": thetas}
        sample["logLs"] = yield (ss, thetas)              # the one likelihood batch of this sampler
        self.samples.append(sample)

        top = np.max(sample["logLs"])
        with np.errstate(under="ignore"):
            w = np.exp(sample["logLs"] - top)
            ev = np.mean(w)
            logev = np.log(ev) + top
            KL = np.mean(sample["logLs"] * w) / ev - logev
        self.evidences.append((logev, 1e-10, KL))    # exact evidence: standard error "zero"
        self.exhausted = True

    # ------------------------------------------------------------------ one AMIS iteration
    def _device_ensemble(self):
        """The device-resident ensemble of this sampler (created on first use), or None: model without a GPU engine, or a
        shape beyond the device kernel (more than 32 slots / 4 states)."""
        if not hasattr(self, "_ens"):
            make = getattr(self.model, "amis_ensemble", None)
            self._ens = make(self.k + 1, self.cfc.transitions) if make is not None and not self.samples else None
        return self._ens

    def _step_host(self, cur):
        """Host (numpy) bookkeeping of one AMIS iteration - models without a GPU engine.  Returns
        ``(summary or None, log_w, new_a, new_logp)``."""
        # the current proposal joins the mixture: update the denominators of all previous samples
        if self.samples:
            sizes = [len(smp["logLs"]) for smp in self.samples]
            old_lp = self.log_proposal(cur, np.concatenate([smp["ss"] for smp in self.samples]),
                                       np.concatenate([smp["thetas"] for smp in self.samples]))
            for smp, lp in zip(self.samples, np.split(old_lp, np.cumsum(sizes)[:-1])):
                smp["cur_log_proposal"] = lp
                with np.errstate(under="ignore"):
                    smp["logδs"] = np.logaddexp(smp["logδs"], lp)

        # new sample: RNG order Dirichlet -> CFC, as the reference
        new = {"ss": self.dirichlet.sample(cur[0], self.N), "thetas": self.cfc.sample(cur[1], self.N)}
        new["logLs"] = yield (new["ss"], new["thetas"])
        all_lp = self.log_proposal_multi(self.parameters, new["ss"], new["thetas"])   # past proposals and the current one
        new["cur_log_proposal"] = all_lp[-1]
        new["logδs"] = _lse(all_lp, axis=0)
        self.samples.append(new)

        # deterministic-mixture weights of the full ensemble
        n_steps = len(self.parameters)
        ens = {key: np.concatenate([smp[key] for smp in self.samples], axis=0) for key in self.samples[-1]}
        summary = None
        if hasattr(self.model, "amis_weights"):
            log_w, summary = self.model.amis_weights(ens["logLs"], ens["logδs"], ens["cur_log_proposal"], np.log(n_steps))
        else:
            log_w = ens["logLs"] - ens["logδs"] + np.log(n_steps)
        ens["log_weights"] = log_w
        for smp, lw in zip(self.samples, np.split(log_w, np.cumsum([len(smp["logLs"]) for smp in self.samples])[:-1])):
            smp["log_weights"] = lw

        # refit the proposal
        new_a = self.dirichlet.estimate(ens["ss"], log_w)
        new_logp = self.cfc.estimate(ens["thetas"], log_w)
        self._host_ens = {"logLs": ens["logLs"], "cur_log_proposal": ens["cur_log_proposal"]}
        return summary, log_w, new_a, new_logp

    def step(self):
        """Returns False (and does nothing) if the sampler is exhausted, True otherwise."""
        return drive(self.step_gen(), self.logL)

    def step_gen(self):
        """Generator form of `step`: yields the ``(ss, thetas)`` batch whose likelihoods it needs."""
        if self.exhausted:
            return False
        cur = self.parameters[-1]

        device = self._device_ensemble()
        if device is not None:
            # ---- bookkeeping on the device (bildk_amis_step): the ensemble and all proposals live in HBM; only the new
            #      batch goes up, only the statistics of the refit (and the per-sample log weights) come back
            new = {"ss": self.dirichlet.sample(cur[0], self.N), "thetas": self.cfc.sample(cur[1], self.N)}   # RNG order: Dirichlet -> CFC
            fused = device.fused_step(new["ss"], new["thetas"], cur[0], cur[1])
            new["logLs"] = yield LikelihoodRequest(new["ss"], new["thetas"], fused)
            new["logLs"] = np.asarray(new["logLs"], dtype=float)
            if fused.submitted:                      # the driver let the step ride on the likelihood launch
                summary, mom_m, mom_v, log_marginals, per = fused.result()
            else:
                summary, mom_m, mom_v, log_marginals, per = device.step(new["ss"], new["thetas"], new["logLs"], cur[0], cur[1])
            self.samples.append(new)
            log_w = per[:, 0]
            lo = 0
            for smp in self.samples:                 # views into the ensemble-level buffer the device call refreshed
                hi = lo + len(smp["logLs"])
                smp["log_weights"], smp["logδs"], smp["cur_log_proposal"] = per[lo:hi, 0], per[lo:hi, 1], per[lo:hi, 2]
                lo = hi
            old_a, old_logp = cur
            new_a = self.dirichlet.estimate_from_moments(mom_m.copy(), mom_v)
            new_logp = self.cfc.estimate_from_log_marginals(log_marginals)
        else:
            summary, log_w, new_a, new_logp = yield from self._step_host(cur)
            ens = self._host_ens
            old_a, old_logp = cur

        ratio = np.log(np.sum(new_a) / np.sum(old_a))           # concentration brake
        cap = self.N * self.brakes[0]
        if np.abs(ratio) > cap:
            new_a *= np.exp(np.sign(ratio) * cap - ratio)

        with np.errstate(under="ignore"):                        # polarisation brake, slot by slot, in linear space
            old_p, new_p = np.exp(old_logp), np.exp(new_logp)
        cap = self.N * self.brakes[1]
        for i in range(new_p.shape[1]):
            delta = new_p[:, i] - old_p[:, i]
            biggest = np.max(np.abs(delta))
            if biggest > cap:
                new_logp[:, i] = np.log(old_p[:, i] + cap * delta / biggest)
        self.parameters.append((new_a, new_logp))

        # evidence, its standard error, and KL(posterior || proposal)
        n = len(log_w)
        if summary is not None:
            top, s1, ssd, s3 = summary            # max, sum w, sum (w - mean)^2, nansum w (logL - log q)
            ev = s1 / n
            sem = math.sqrt(ssd / (n - 1)) / math.sqrt(n) if n > 1 else np.nan
        else:
            top = np.max(log_w)
            with np.errstate(under="ignore"):
                w = np.exp(log_w - top)
            ev = np.mean(w)
            sem = np.std(w, ddof=1) / np.sqrt(n) if n > 1 else np.nan
            with np.errstate(under="ignore", invalid="ignore"):
                s3 = np.nansum(w * (ens["logLs"] - ens["cur_log_proposal"]))
        logev = np.log(ev) + top + self.logprior
        dlogev = sem / ev
        KL = s3 / n / ev - logev + self.logprior
        self.evidences.append((logev, dlogev, KL))

        if (len(self.samples) + 1) * self.N >= self.max_fev:
            self.exhausted = True
        return True

    # ------------------------------------------------------------------ results
    def tstat(self, other):
        """Evidence separation from another sampler in units of the combined standard error."""
        a, da = self.evidences[-1][:2]
        b, db = other.evidences[-1][:2]
        return (a - b) / np.sqrt(da ** 2 + db ** 2)

    def MAP_profile(self):
        """The sampled profile with the highest likelihood so far."""
        best_in = np.array([np.argmax(smp["logLs"]) for smp in self.samples])
        best_val = np.array([smp["logLs"][i] for smp, i in zip(self.samples, best_in)])
        j = np.argmax(best_val)
        return self.st2profile(self.samples[j]["ss"][best_in[j]], self.samples[j]["thetas"][best_in[j]])

    def _ensemble_states(self):
        """(n_samples, T) per-frame states of the whole ensemble, without a Python loop over samples."""
        ss = np.concatenate([smp["ss"] for smp in self.samples])
        thetas = np.concatenate([smp["thetas"] for smp in self.samples])
        T = len(self.traj)
        n, K1 = ss.shape
        if K1 == 1:
            return np.broadcast_to(thetas[:, :1], (n, T)).copy()
        switches = np.floor(np.cumsum(ss, axis=1)[:, :-1] * (T - 1)).astype(int) + 1      # (n, k)
        run = np.sum(np.arange(T)[None, :, None] >= switches[:, None, :], axis=2)         # index of the run of each frame
        return np.take_along_axis(thetas, run, axis=1)

    def log_marginal_posterior(self):
        """``(n_states, T)`` normalised log posterior probability of each state at each frame."""
        try:
            log_w = np.concatenate([smp["log_weights"] for smp in self.samples])
        except KeyError:                              # exhaustive sampling: weights are the likelihoods
            log_w = np.concatenate([smp["logLs"] for smp in self.samples])
        if hasattr(self.model, "marginal_posterior"):      # weighted per-frame state histogram on the device
            ss = np.concatenate([smp["ss"] for smp in self.samples])
            thetas = np.concatenate([smp["thetas"] for smp in self.samples])
            return self.model.marginal_posterior(ss, thetas, len(self.traj), log_w)
        states = self._ensemble_states()
        n = self.model.nStates
        onehot = states[:, None, :] == np.arange(n)[None, :, None]
        logpost = _lse(np.broadcast_to(log_w[:, None, None], onehot.shape), axis=0, mask=onehot)
        return logpost - _lse(logpost, axis=0, keepdims=True)
