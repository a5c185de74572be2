"""
Inference models behind the reference's model interface.

Mirrors /root/reference/bild/models.py for the hot path:

* `MultiStateModel`  - the interface the samplers program against (models.py:24-160)
* `MultiStateRouse`  - the multi-state Rouse model (models.py:163-370).  ``logL(profile, traj)`` keeps
  the reference signature; the work is done by the sm_100a engine (bild_b200/engine.py) instead of
  the Cython module the reference plugs in at cython_imports.py:3-7.  New: ``logL_batch`` /
  ``logL_st_batch`` evaluate a whole AMIS batch in one launch - `FixedkSampler.logL` uses them.
* `FactorizedModel`  - the table-lookup model (models.py:372-534); host-only, kept because
  ``MultiStateRouse.initial_loopingprofile`` and the samplers' own tests need it.

`GenericGaussianModel` (models.py:536-728) depends on the third-party ``bayesmsd`` and is a different
likelihood family; it is out of scope.
"""
import abc
from collections import OrderedDict

import numpy as np
import scipy.stats

from . import rouse
from .trajectory import Trajectory
from .util import Loopingprofile

__all__ = ["MultiStateModel", "MultiStateRouse", "FactorizedModel"]


class MultiStateModel(metaclass=abc.ABCMeta):
    """
    Interface of an inference model: a likelihood ``logL(profile, traj)``, the number of states, the
    spatial dimension, and which state transitions are allowed (``transitions[i, j]``: i -> j).
    Implementations call ``init_transitions(n)`` at the end of their constructor.
    """

    def init_transitions(self, n):
        self.transitions = ~np.eye(n, dtype=bool)

    @property
    def nStates(self):
        return self.transitions.shape[0]

    @property
    def d(self):
        raise NotImplementedError  # pragma: no cover

    def initial_loopingprofile(self, traj):
        """Default guess: a uniformly random profile (global numpy RNG, as the reference does)."""
        return Loopingprofile(np.random.choice(self.nStates, size=len(traj)))

    @abc.abstractmethod
    def logL(self, loopingprofile, traj):
        """log-likelihood of ``traj`` given ``loopingprofile``"""
        raise NotImplementedError  # pragma: no cover

    # -- shared argument handling for the generative models (models.py:139-157) -------------------
    def _resolve_localization_error(self, localization_error):
        if np.isscalar(localization_error):
            localization_error = self.d * [localization_error]
        localization_error = np.asarray(localization_error)
        if localization_error.shape != (self.d,):
            raise ValueError("Did not understand localization_error")
        return localization_error

    @staticmethod
    def _resolve_missing_frames(missing_frames, T):
        """None/0: none; float in (0,1): each frame independently; int: that many; array: those indices."""
        if missing_frames is None or (np.isscalar(missing_frames) and missing_frames == 0):
            return np.array([], dtype=int)
        if np.isscalar(missing_frames):
            if 0 < missing_frames < 1:
                return np.nonzero(np.random.rand(T) < missing_frames)[0]
            return np.random.choice(T, size=missing_frames, replace=False).astype(int)
        return np.asarray(missing_frames, dtype=int)

    def trajectory_from_loopingprofile(self, profile, localization_error=None, missing_frames=None, preproc=None):
        if preproc == "localization_error":
            return self._resolve_localization_error(localization_error)
        if preproc == "missing_frames":
            return self._resolve_missing_frames(missing_frames, len(profile))
        raise NotImplementedError  # pragma: no cover


class MultiStateRouse(MultiStateModel):
    """
    Multi-state Rouse model; same constructor as the reference (models.py:222-249).

    Parameters
    ----------
    N : int
        number of monomers
    D, k : float
        1d diffusion constant of a free monomer; backbone spring constant
    d : int
        spatial dimension
    looppositions : sequence
        one entry per state: ``None`` (free chain), a bond ``(left, right[, rel_strength])`` or a list of bonds
    measurement : "end2end" or (N,) array
        measured distance vector; "end2end" is ``e_{N-1} - e_0``
    localization_error : None, float or (d,) array
        overrides ``traj.localization_error`` when given
    device : int, optional
        CUDA device the likelihood engine lives on (default: ``BILD_B200_DEVICE`` or ``LOCAL_RANK`` or 0)
    """

    def __init__(self, N, D, k, d=3, looppositions=(None, (0, -1)), measurement="end2end",
                 localization_error=None, device=None):
        self._d = d
        if isinstance(measurement, str):
            if measurement != "end2end":
                raise ValueError(f"unknown measurement {measurement!r}")
            measurement = np.zeros(N)
            measurement[0], measurement[-1] = -1.0, 1.0
        measurement = np.asarray(measurement, dtype=float)
        assert len(measurement) == N
        self.measurement = measurement

        if localization_error is not None and np.isscalar(localization_error):
            localization_error = localization_error * np.ones(d)
        self.localization_error = localization_error

        self.models = []
        for loop in looppositions:
            if loop is not None and np.isscalar(loop[0]):
                loop = [loop]
            self.models.append(rouse.Model(N, D, k, d, add_bonds=loop))
        self.init_transitions(len(self.models))

        self.device = device
        self._sharder = None
        self._engine = None
        self._engine_built_for = None
        self._handles = OrderedDict()   # id(traj) -> (fingerprint, TrajectoryHandle)
        self._max_handles = 8192

    @property
    def d(self):
        return self._d

    def _get_noise(self, traj):
        """Localisation error that applies to ``traj``: the model's, else the trajectory's (models.py:255-263)."""
        if self.localization_error is not None:
            return np.asarray(self.localization_error)
        if getattr(traj, "localization_error", None) is not None:
            return np.asarray(traj.localization_error)
        raise ValueError("No localization error specified (use MultiStateModel.localization_error or Trajectory.localization_error)")

    # ------------------------------------------------------------------ GPU plumbing
    def _engine_key(self):
        """What the GPU copy depends on: the per-state ``_dynamics`` (by identity - `update_dynamics` installs a new
        dict) and the measurement vector.  The reference re-reads these on every call (pyx:150-160), so an edit of
        ``models[i]`` (D, k, `update_dynamics`) or of ``measurement`` between two likelihood calls must take effect."""
        for m in self.models:
            m.check_dynamics()             # as pyx:152-153: refreshes dynamics whose N / D / k / dt changed
        return [m._dynamics for m in self.models], np.array(self.measurement, dtype=float)

    @property
    def engine(self):
        """The GPU-resident model: built on first use, rebuilt when the dynamics or the measurement vector changed."""
        dyn, w = self._engine_key()
        if self._engine is not None:
            old_dyn, old_w = self._engine_built_for
            if len(old_dyn) == len(dyn) and all(a is b for a, b in zip(old_dyn, dyn)) and np.array_equal(old_w, w):
                return self._engine
            self.reset_engine()
        import os
        from .engine import RouseEngine
        dev = self.device
        if dev is None:
            dev = int(os.environ.get("BILD_B200_DEVICE", os.environ.get("LOCAL_RANK", 0)))
        self._engine = RouseEngine.from_models(self.models, self.measurement, device=dev)
        self._engine_built_for = (dyn, w)
        return self._engine

    def reset_engine(self):
        self._engine = None
        self._engine_built_for = None
        self._handles.clear()

    def _handle(self, traj):
        noise = np.asarray(self._get_noise(traj), dtype=float)
        if noise.ndim == 0:
            noise = noise * np.ones(self.d)
        data = np.ascontiguousarray(traj[:], dtype=float)
        fp = (hash(data.tobytes()), hash(noise.tobytes()), data.shape)
        key = id(traj)
        hit = self._handles.get(key)
        if hit is not None and hit[0] == fp:
            self._handles.move_to_end(key)
            return hit[1]
        h = self.engine.trajectory(data, noise)
        self._handles[key] = (fp, h)
        while len(self._handles) > self._max_handles:
            self._handles.popitem(last=False)
        return h

    # ------------------------------------------------------------------ likelihood
    def logL(self, profile, traj):
        """Rouse likelihood by Kalman filter, one profile (signature of models.py:265-278)."""
        return float(self.engine.logl_states(self._handle(traj), np.asarray(profile[:], dtype=np.int32))[0])

    def logL_batch(self, profiles, traj):
        """Many profiles (Loopingprofiles or a (P, T) int array), one launch -> (P,) float64."""
        if isinstance(profiles, np.ndarray):
            states = profiles
        else:
            states = np.array([p[:] for p in profiles], dtype=np.int32).reshape(len(profiles), len(traj))
        return self.engine.logl_states(self._handle(traj), states)

    def logL_st_batch(self, ss, thetas, traj):
        """Batched ``logL(st2profile(s, theta), traj)`` for AMIS samples (replaces the loop at amis.py:735-739).
        With `shard_over` set, the batch is split across the ranks of a process group (bild_b200/dist.py)."""
        if self._sharder is not None:
            return self._sharder(lambda a, b: self._logL_st_local(a, b, traj), np.asarray(ss), np.asarray(thetas))
        return self._logL_st_local(ss, thetas, traj)

    def logL_runs_multi(self, trajs, offsets, starts, run_states):
        """Run-length profiles of MANY trajectories in one launch: profiles [offsets[i], offsets[i+1]) belong
        to ``trajs[i]`` (used by `bild_b200.dataset.sample_many`)."""
        return self.engine.logl_runs_multi([self._handle(t) for t in trajs], offsets, starts, run_states)

    def logL_runs_multi_submit(self, trajs, offsets, starts, run_states, amis=None):
        """Asynchronous `logL_runs_multi`: returns an object whose ``wait()`` gives the log-likelihoods; the launch runs
        while the caller does host work (at most two batches in flight).  ``amis``: per trajectory the fused AMIS step
        of its batch (`bild_b200.engine.FusedAmisStep`) or None."""
        return self.engine.logl_runs_multi_submit([self._handle(t) for t in trajs], offsets, starts, run_states, amis=amis)

    def _logL_st_local(self, ss, thetas, traj):
        return self.engine.logl_st(self._handle(traj), ss, thetas)

    def shard_over(self, group=None, device=None):
        """Evaluate every AMIS batch sharded over the ranks of ``group`` (one all-gather of logL per batch)."""
        from .dist import ShardedEvaluator
        if device is None:
            device = f"cuda:{self.engine.device}"
        self._sharder = ShardedEvaluator(group, device)
        return self

    def amis_weights(self, logLs, logdeltas, cur_log_proposal, log_nsteps):
        """
        AMIS weight normalisation on the device (amis.py:843-845, 878-900): returns ``log_weights`` and
        ``(max, sum w, sum (w - mean)^2, nansum w (logL - log q))`` with ``w = exp(log_weights - max)``,
        reduced in a fixed order (identical bits on every rank).
        """
        import ctypes
        from . import _lib
        a, b, c = (_lib.as_f64(v) for v in (logLs, logdeltas, cur_log_proposal))
        n = len(a)
        log_w = np.empty(n)
        st = np.empty(4)
        dp = _lib.c_double_p
        _lib.check(_lib.load().bildk_amis_weights(n, _lib.ptr(a, dp), _lib.ptr(b, dp), _lib.ptr(c, dp), float(log_nsteps),
                                                  _lib.ptr(log_w, dp), _lib.ptr(st, dp), self.engine.device))
        return log_w, tuple(st)

    def amis_ensemble(self, K1, transitions):
        """Device-resident ensemble for one `FixedkSampler` (bild_b200.engine.AmisEnsemble), or None when the shape is
        outside what the device kernel handles (the sampler then keeps its bookkeeping on the host)."""
        from .engine import AmisEnsemble
        if not AmisEnsemble.supports(K1, len(transitions)):
            return None
        return AmisEnsemble(K1, transitions, device=self.engine.device)

    def marginal_posterior(self, ss, thetas, T, log_weights):
        """
        ``(n_states, T)`` normalised log posterior probability of each state at each frame from the weighted
        ensemble ``(ss, thetas, log_weights)`` (amis.py:942-972), reduced on the device from the run-length profiles
        (the reference builds an ``(n, S, T)`` boolean tensor on the host).
        """
        from . import _lib
        from .engine import st_to_runs
        starts, states = st_to_runs(ss, thetas, T)
        lw = _lib.as_f64(log_weights)
        out = np.empty((self.nStates, T))
        _lib.check(_lib.load().bildk_marginal_posterior(len(lw), starts.shape[1], T, self.nStates,
                                                        _lib.ptr(starts, _lib.c_int32_p), _lib.ptr(states, _lib.c_uint8_p),
                                                        _lib.ptr(lw, _lib.c_double_p), _lib.ptr(out, _lib.c_double_p),
                                                        self.engine.device))
        return out

    # ------------------------------------------------------------------ helpers shared with the reference API
    def initial_loopingprofile(self, traj):
        return self.toFactorized().initial_loopingprofile(traj)

    def trajectory_from_loopingprofile(self, profile, localization_error=None, missing_frames=None):
        """Generative model (models.py:295-350): steady-state start, one `evolve` per frame, NaN rows, noise."""
        if localization_error is None:
            if self.localization_error is None:
                raise ValueError("Need to specify either localization_error or model.localization_error")
            localization_error = self.localization_error
        localization_error = self._resolve_localization_error(localization_error)
        missing_frames = self._resolve_missing_frames(missing_frames, len(profile))

        data = np.full((len(profile), self.d), np.nan)
        conf = self.models[profile[0]].conf_ss()
        data[0] = self.measurement @ conf
        for i in range(1, len(profile)):
            conf = self.models[profile[i]].evolve(conf)
            data[i] = self.measurement @ conf
        data[missing_frames, :] = np.nan
        data += localization_error[None, :] * np.random.normal(size=data.shape)
        return Trajectory(data, localization_error=localization_error, loopingprofile=profile)

    def toFactorized(self):
        """`FactorizedModel` built from the exact steady-state distance distributions (models.py:352-370)."""
        noise2 = 0 if self.localization_error is None else np.sum(np.asarray(self.localization_error) ** 2) / self.d
        dists = []
        for mod in self.models:
            _, C = mod.steady_state()
            var = self.measurement @ C @ self.measurement + noise2
            dists.append(scipy.stats.maxwell(scale=np.sqrt(var)))
        return FactorizedModel(dists, d=self.d)


class FactorizedModel(MultiStateModel):
    """
    Time-scale-separated model: every frame's distance is drawn independently from the distribution of
    the frame's state (anything with ``logpdf`` and, for sampling, ``rvs``).  Ignores the trajectory's
    localisation error (it is assumed to be part of the distributions).  Per-trajectory log-pdf tables
    are memoised; `clear_memo` drops them.
    """

    def __init__(self, distributions, d=3):
        self.distributions = distributions
        self._d = d
        self._known_trajs = dict()
        self.init_transitions(len(self.distributions))

    @property
    def d(self):
        return self._d

    def clear_memo(self):
        self._known_trajs = dict()

    def _table(self, traj):
        entry = self._known_trajs.get(traj)
        if entry is None:
            dist = traj.abs()[:][:, 0]
            with np.errstate(divide="ignore"):   # NaN frames
                entry = {"logL_table": np.array([dd.logpdf(dist) for dd in self.distributions])}
            self._known_trajs[traj] = entry
        return entry["logL_table"]

    def initial_loopingprofile(self, traj):
        """Frame-wise maximum-likelihood state; missing frames take the state of the next valid frame."""
        table = self._table(traj)
        T = len(traj)
        valid = np.nonzero(~np.any(np.isnan(traj[:]), axis=1))[0]
        best = np.argmax(table[:, valid], axis=0)
        # frame t is assigned the state of the first valid frame >= t (the last valid one beyond the end)
        nxt = np.minimum(np.searchsorted(valid, np.arange(T), side="left"), len(valid) - 1)
        return Loopingprofile(best[nxt])

    def logL(self, profile, traj):
        table = self._table(traj)
        return np.nansum(table[np.asarray(profile[:]), np.arange(len(profile))])

    def trajectory_from_loopingprofile(self, profile, localization_error=0.0, missing_frames=None):
        localization_error = self._resolve_localization_error(localization_error)
        missing_frames = self._resolve_missing_frames(missing_frames, len(profile))
        magnitudes = np.array([self.distributions[state].rvs() for state in profile[:]])
        data = np.random.normal(size=(len(magnitudes), self.d))
        data *= (magnitudes / np.linalg.norm(data, axis=1))[:, None]
        data[missing_frames, :] = np.nan
        return Trajectory(data, localization_error=localization_error, loopingprofile=profile)
