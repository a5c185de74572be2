"""
`sample` - the entry point of BILD - and its result container.

Behavioural mirror of /root/reference/bild/core.py: the same state machine over the number of switches
k (core.py:202-229), the same RNG call order, the same ``log`` contents.  The likelihood batches inside
each `FixedkSampler` go to the GPU when the model is a `bild_b200.models.MultiStateRouse`.
"""
import numpy as np
from tqdm.auto import tqdm

from .amis import FixedkSampler, _lse, drive
from .choicesampler import ChoiceSampler
from .trajectory import make_Trajectory

__all__ = ["sample", "sample_gen", "SamplingResults"]


def sample(traj, model, dE=0, init_runs=20, certainty_in_k=0.99, k_lookahead=2, k_max=20,
           sampler_kw={}, choice_kw={}, show_progress=False):
    """
    Infer the looping profile of a trajectory: run AMIS samplers for k = 0, 1, ... switches, always
    spending the next AMIS step where it is expected to be most informative about the best k (evidence
    within ``dE`` of the maximum, smallest such k), until that choice is ``certainty_in_k`` certain.

    Parameters
    ----------
    traj : Trajectory or array
    model : MultiStateModel
    dE : float
        evidence margin
    init_runs : int
        AMIS steps a new sampler runs before it competes
    certainty_in_k : float in (0, 1)
        stop when the choice distribution puts this much mass on one k
    k_lookahead : int
        how many ks beyond the current best must be explored
    k_max : int
        largest k for which a sampler may be created
    sampler_kw, choice_kw : dict
        forwarded to `FixedkSampler` / `ChoiceSampler`
    show_progress : bool

    Returns
    -------
    SamplingResults
    """
    traj = make_Trajectory(traj)
    # every likelihood batch of the run goes through FixedkSampler.logL of a throw-away sampler bound to (traj, model)
    probe = FixedkSampler.__new__(FixedkSampler)
    probe.traj, probe.model = traj, model
    return drive(sample_gen(traj, model, dE=dE, init_runs=init_runs, certainty_in_k=certainty_in_k, k_lookahead=k_lookahead,
                            k_max=k_max, sampler_kw=sampler_kw, choice_kw=choice_kw, show_progress=show_progress), probe.logL)


def sample_gen(traj, model, dE=0, init_runs=20, certainty_in_k=0.99, k_lookahead=2, k_max=20,
               sampler_kw={}, choice_kw={}, show_progress=False):
    """
    Generator form of `sample` (same arguments): yields every ``(ss, thetas)`` batch whose likelihoods the run needs
    and expects them to be sent back; returns the `SamplingResults`.  `sample` drives it with the model's own batched
    likelihood; `bild_b200.dataset.sample_many` drives many of them and fuses their batches into one launch.
    """
    bar = tqdm(disable=not show_progress)
    traj = make_Trajectory(traj)
    samplers = []
    log = {"k": [], "pk": [], "KLD": [], "I_la": []}
    state = {"fresh": False}

    def take_step(k):
        if (yield from samplers[k].step_gen()):            # no-op for exhausted samplers
            bar.update()
            for key in log:
                log[key].append(None)
            log["k"][-1] = k
            state["fresh"] = True

    def new_sampler(k):
        assert k == len(samplers)
        fresh = FixedkSampler(traj, model, k=k, _defer=True, **sampler_kw)
        yield from fresh.start_gen()
        samplers.append(fresh)
        for _ in range(init_runs):
            yield from take_step(k)

    def next_k():
        k_new = len(samplers)
        if not state["fresh"]:
            return k_new if len(log["k"]) == 0 else log["k"][-1]

        logE = np.array([s.evidences[-1][0] for s in samplers])
        dlogE = np.array([s.evidences[-1][1] for s in samplers])
        steps = np.array([np.inf if s.exhausted else len(s.samples) for s in samplers])
        cs = ChoiceSampler(logE, dlogE ** 2, steps, dE, **choice_kw)
        pk = cs.n0 / cs.samplesize

        if k_new < k_lookahead + 1 and k_new <= k_max:
            # every sampler so far lies inside the lookahead region: go on to the next k unconditionally
            choice, KLD, I_la = k_new, None, np.inf
        else:
            KLD = cs.KLD_moreSamples()
            choice = np.argmax(KLD)
            I_la = cs.KLD_omitK(np.arange(k_new - k_lookahead, k_new)) if k_new >= k_lookahead + 1 else np.inf
            if I_la > KLD[choice] and k_new <= k_max:
                choice = k_new

        log["pk"][-1], log["KLD"][-1], log["I_la"][-1] = pk, KLD, I_la
        state["fresh"] = False
        return choice

    k_next = 0
    running = True
    failure = None
    try:
        while running:
            if k_next < len(samplers):
                yield from take_step(k_next)
            elif k_next == len(samplers):
                yield from new_sampler(k_next)
            else:  # pragma: no cover
                raise RuntimeError("Trying to sample outside of existing range; this is a bug")
            k_next = next_k()
            if k_next == len(samplers):
                running = True                     # a higher k is needed: takes precedence over certainty
            else:
                running = np.max(log["pk"][-1]) < certainty_in_k
                if log["KLD"][-1] is not None:     # ... and the proposed step must still carry information
                    running &= log["KLD"][-1][k_next] > 0
        bar.close()
    except KeyboardInterrupt:  # pragma: no cover
        pass                                       # hand back what we have
    except Exception as err:  # noqa: BLE001
        # The reference returns from a ``finally`` block (core.py:231-236), i.e. it hands back the partial
        # results on ANY exception (e.g. "Iteration did not converge" from the CFC fit, amis.py:392).  Same here,
        # but the exception is kept on the results instead of being dropped silently.
        failure = err
    out = SamplingResults(traj, model, dE, samplers, log)
    out.error = failure
    return out


class SamplingResults:
    """
    Output of `sample`: the samplers, the evidence curve and convenience accessors.

    Attributes
    ----------
    traj, model, dE, samplers
    log : dict of arrays - per AMIS step: ``k`` sampled, choice distribution ``pk``, expected gains ``KLD``,
        lookahead importance ``I_la`` (ragged entries are NaN padded)
    k, evidence, evidence_se : arrays over the samplers
    error : None, or the exception that ended the run early (the results are then partial)
    """

    error = None

    def __init__(self, traj, model, dE, samplers, log=None):
        self.traj, self.model, self.dE, self.samplers = traj, model, dE, samplers
        self.log = {}
        if log is not None:
            for key, rows in log.items():
                if key in ("k", "I_la"):
                    self.log[key] = np.array(rows)
                else:
                    width = max([1 if r is None else len(r) for r in rows], default=0)
                    arr = np.full((len(rows), width), np.nan)
                    for i, r in enumerate(rows):
                        if r is not None:
                            arr[i, :len(r)] = r
                    self.log[key] = arr

    @property
    def k(self):
        return np.array([s.k for s in self.samplers])

    @property
    def evidence(self):
        return np.array([s.evidences[-1][0] for s in self.samplers])

    @property
    def evidence_se(self):
        return np.array([s.evidences[-1][1] for s in self.samplers])

    def best_k(self, dE=None):
        """Smallest k whose evidence is within ``dE`` (default: the run's) of the maximum."""
        if dE is None:
            dE = self.dE
        return np.min(self.k[self.evidence >= np.max(self.evidence) - dE])

    def best_profile(self, dE=None):
        return self.samplers[self.best_k(dE)].MAP_profile()

    def log_marginal_posterior(self, dE=None):
        """``(n_states, T)`` log posterior state probabilities, for the best k or (``dE='average'``)
        evidence-averaged over k."""
        if isinstance(dE, str) and dE == "average":
            terms = [s.log_marginal_posterior() + ev for s, ev in zip(self.samplers, self.evidence) if ev > -np.inf]
            logpost = _lse(np.array(terms), axis=0)
            return logpost - _lse(logpost, axis=0, keepdims=True)
        return self.samplers[self.best_k(dE)].log_marginal_posterior()
