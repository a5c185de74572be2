"""
Host-side engine: owns the GPU-resident model and trajectory handles and turns profile batches into
one kernel launch through the C ABI (include/bild_b200.h).

What the reference does per likelihood call (/root/reference/bild/src/MSRouse_logL.pyx:143-199:
``np.unique`` of the noise, restacking B/G/Sig of every state, ``steady_state()``, NaN mask) happens
here ONCE per model / trajectory; what it does per profile in a Python loop
(/root/reference/bild/amis.py:735-739) happens here once per batch.
"""
import ctypes
import weakref

import numpy as np

from . import _lib
from ._lib import c_double_p, c_int32_p, c_uint8_p, c_uint32_p, as_f64, ptr

__all__ = ["RouseEngine", "TrajectoryHandle", "AmisEnsemble", "FusedAmisStep", "st_to_runs", "states_to_runs"]


def st_to_runs(ss, thetas, T):
    """
    Run-length form of ``FixedkSampler.st2profile`` for a whole batch.

    Uses the identical numpy expression as /root/reference/bild/amis.py:687-688
    (``floor(cumsum(s)[:-1] * (T-1)).astype(int) + 1``) so that the discrete profiles are the very
    same ones; run ``r`` of profile ``p`` covers frames ``[starts[p, r], starts[p, r+1])`` with state
    ``thetas[p, r]`` (last run to ``T``); empty runs vanish like the empty numpy slices do.
    """
    ss = np.asarray(ss, dtype=float)
    thetas = np.asarray(thetas)
    if ss.ndim == 1:
        ss, thetas = ss[None, :], thetas[None, :]
    P, K1 = ss.shape
    starts = np.zeros((P, K1), dtype=np.int32)
    if K1 > 1:
        switchpos = np.cumsum(ss, axis=1)[:, :-1]
        starts[:, 1:] = np.floor(switchpos * (T - 1)).astype(int) + 1
    return starts, _as_state_bytes(thetas)


def _as_state_bytes(states):
    """uint8 copy of a state array; refuses values that would wrap (the C ABI checks < S, a silent wrap would pass it)."""
    states = np.asarray(states)
    if states.size and (states.min() < 0 or states.max() > 255):
        raise ValueError("state indices must lie in [0, 255]")
    return np.ascontiguousarray(states, dtype=np.uint8)


def states_to_runs(states):
    """Run-length code per-frame state arrays (P, T) -> (starts (P, K1) int32, run_states (P, K1) uint8)."""
    states = np.asarray(states)
    if states.size and (states.min() < 0 or states.max() > 255):
        raise ValueError("state indices must lie in [0, 255]")
    if states.ndim == 1:
        states = states[None, :]
    P, T = states.shape
    change = np.ones((P, T), dtype=bool)
    change[:, 1:] = states[:, 1:] != states[:, :-1]
    nruns = change.sum(axis=1)
    K1 = int(nruns.max())
    starts = np.full((P, K1), T, dtype=np.int32)
    rstates = np.zeros((P, K1), dtype=np.uint8)
    pi, ti = np.nonzero(change)
    ri = (np.cumsum(change, axis=1) - 1)[pi, ti]
    starts[pi, ri] = ti
    rstates[pi, ri] = states[pi, ti]
    return starts, rstates


class TrajectoryHandle:
    """A trajectory resident on the GPU: data, missing-frame mask and localisation-error structure."""

    def __init__(self, engine, x, localization_error):
        x = as_f64(x)
        if x.ndim == 1:
            x = x[:, None]
        if x.ndim != 2 or x.shape[1] != engine.d:
            raise ValueError(f"trajectory must have shape (T, {engine.d}), got {x.shape}")
        err = np.asarray(localization_error, dtype=float)
        if err.shape != (engine.d,):
            raise ValueError(f"localization_error must have shape ({engine.d},), got {err.shape}")
        # MSRouse_logL.pyx:145-147
        uniq, cind = np.unique(err, return_inverse=True)
        self.s2 = as_f64(uniq * uniq)
        self.Cind = np.ascontiguousarray(cind, dtype=np.uint32)
        self.T = x.shape[0]
        self.engine = engine
        lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(lib.bildk_traj_create(engine._h, self.T, ptr(x, c_double_p), len(self.s2), ptr(self.s2, c_double_p),
                                         ptr(self.Cind, c_uint32_p), ctypes.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, lib.bildk_traj_destroy, h)

    def describe_plan(self, P):
        return _lib.load().bildk_describe_plan(self._h, int(P)).decode()


class AmisEnsemble:
    """
    Device-resident ensemble of one `FixedkSampler` (C ABI ``bildk_amis_*``): samples, likelihoods, mixture
    denominators, weights and all past proposals live in HBM; `step` adds a batch, lets the proposal it was drawn
    from join the mixture, and returns the statistics of amis.py:843-845, 878-900, 137-151, 300-303 in one launch.
    """

    MAX_K1, MAX_STATES = 32, 4

    def __init__(self, K1, transitions, device=0):
        transitions = np.ascontiguousarray(transitions, dtype=np.uint8)
        self.K1, self.S = int(K1), transitions.shape[0]
        lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(lib.bildk_amis_create(self.K1, self.S, ptr(transitions, c_uint8_p), int(device), ctypes.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, lib.bildk_amis_destroy, h)
        self.n = 0
        self._per = np.empty((1024, 3))

    @classmethod
    def supports(cls, K1, n_states):
        return K1 <= cls.MAX_K1 and n_states <= cls.MAX_STATES

    def _check_batch(self, ss, thetas, a, logp):
        ss = np.ascontiguousarray(ss, dtype=np.float64)
        thetas = np.ascontiguousarray(thetas, dtype=np.int64)
        a = np.ascontiguousarray(a, dtype=np.float64)
        logp = np.ascontiguousarray(logp, dtype=np.float64)
        if ss.ndim != 2 or ss.shape[1] != self.K1 or thetas.shape != ss.shape or a.shape != (self.K1,) or logp.shape != (self.S, self.K1):
            raise ValueError("inconsistent AMIS batch")
        n_tot = self.n + len(ss)
        if n_tot > len(self._per):
            self._per = np.empty((max(n_tot, 2 * len(self._per)), 3))   # contents are rewritten in full by every step
        return ss, thetas, a, logp, np.empty(4 + 2 * self.K1 + self.S * self.K1)

    def _unpack(self, head, n_tot):
        K1 = self.K1
        return tuple(head[:4]), head[4:4 + K1], head[4 + K1:4 + 2 * K1], head[4 + 2 * K1:].reshape(self.S, K1), self._per[:n_tot]

    def step(self, ss, thetas, logLs, a, logp):
        """-> ((max, sum w, sum (w - mean)^2, nansum w (logL - log q)), mean (K1,), var (K1,), log marginals (S, K1),
        per-sample array (n, 3): log_w | logdelta | log q_cur of the whole ensemble - a view that later steps overwrite)."""
        ss, thetas, a, logp, head = self._check_batch(ss, thetas, a, logp)
        logLs = np.ascontiguousarray(logLs, dtype=np.float64)
        n_new = len(logLs)
        if len(ss) != n_new:
            raise ValueError("inconsistent AMIS batch")
        _lib.check(_lib.load().bildk_amis_step(self._h, n_new, ptr(ss, c_double_p), ptr(thetas, _lib.c_int64_p),
                                               ptr(logLs, c_double_p), ptr(a, c_double_p), ptr(logp, c_double_p), ptr(head, c_double_p),
                                               ptr(self._per, c_double_p)))
        self.n += n_new
        return self._unpack(head, self.n)

    def fused_step(self, ss, thetas, a, logp):
        """The same step riding on a likelihood launch (``bildk_logl_runs_multi_submit`` with `amis`): the likelihoods go
        from the filter kernel's output straight into the ensemble on the device.  Returns a `FusedAmisStep`; hand it to
        `RouseEngine.logl_runs_multi_submit` and read ``.result()`` after the batch's ``wait()``."""
        return FusedAmisStep(self, *self._check_batch(ss, thetas, a, logp))


class FusedAmisStep:
    """One AMIS step in flight with a likelihood batch; owns the arrays the C side reads and writes until `wait`."""

    def __init__(self, ens, ss, thetas, a, logp, head):
        self.ens, self.ss, self.thetas, self.a, self.logp, self.head = ens, ss, thetas, a, logp, head
        self.per = ens._per                    # pinned by this reference: a later growth must not free it under the copy
        self.submitted = False

    def fill(self, req):
        """Write the C struct `bildk_amis_req` for this step."""
        req.ens = self.ens._h
        req.ss, req.thetas = ptr(self.ss, c_double_p), ptr(self.thetas, _lib.c_int64_p)
        req.A_cur, req.logp_cur = ptr(self.a, c_double_p), ptr(self.logp, c_double_p)
        req.head, req.per_sample = ptr(self.head, c_double_p), ptr(self.per, c_double_p)

    def mark_submitted(self):
        self.submitted = True
        self.ens.n += len(self.ss)             # the library's ensemble has grown (in stream order)
        self.n_tot = self.ens.n

    def result(self):
        """As `AmisEnsemble.step` (valid after the carrying batch's ``wait()``)."""
        if not self.submitted:
            raise RuntimeError("this AMIS step was not carried by a launch")
        return self.ens._unpack(self.head, self.n_tot)


class PendingBatch:
    """A batch in flight (C ABI ``bildk_logl_runs_multi_submit`` / ``bildk_logl_wait``)."""

    def __init__(self, ticket, out, keep):
        self._ticket, self._out, self._keep = ticket, out, keep      # `keep`: the handles must outlive the launch

    def ready(self):
        """True once `wait` will not block."""
        if self._ticket is None or not self._ticket.value:
            return True
        rc = _lib.load().bildk_logl_ready(self._ticket)
        if rc < 0:
            _lib.check(rc)
        return bool(rc)

    def wait(self):
        if self._ticket is not None:
            ticket, self._ticket = self._ticket, None
            if ticket.value:
                _lib.check(_lib.load().bildk_logl_wait(ticket))
        return self._out


class RouseEngine:
    """
    GPU-resident multi-state Rouse model.

    Parameters
    ----------
    Bs, Gs, Sigs : (S, N, N), (S, N, d), (S, N, N)
        per-state propagators (``m._dynamics['B'|'G'|'Sig']``)
    M0, C0 : (S, N, d), (S, N, N)
        per-state steady states
    w : (N,)
        measurement vector
    device : int
        CUDA device ordinal
    """

    def __init__(self, Bs, Gs, Sigs, M0, C0, w, device=0):
        Bs, Gs, Sigs, M0, C0, w = (as_f64(a) for a in (Bs, Gs, Sigs, M0, C0, w))
        S, N, d = Gs.shape
        if Bs.shape != (S, N, N) or Sigs.shape != (S, N, N) or C0.shape != (S, N, N) or M0.shape != (S, N, d) or w.shape != (N,):
            raise ValueError("inconsistent model array shapes")
        self.S, self.N, self.d, self.device = S, N, d, int(device)
        lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(lib.bildk_model_create(N, d, S, *(ptr(a, c_double_p) for a in (Bs, Gs, Sigs, M0, C0, w)),
                                          self.device, ctypes.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, lib.bildk_model_destroy, h)

    @classmethod
    def from_models(cls, models, measurement, device=0):
        """Build from ``rouse.Model``-like objects (what MSRouse_logL.pyx:152-160 reads on every call)."""
        for m in models:
            m.check_dynamics()
        ss = [m.steady_state() for m in models]
        return cls([m._dynamics["B"] for m in models], [m._dynamics["G"] for m in models],
                   [m._dynamics["Sig"] for m in models], [s[0] for s in ss], [s[1] for s in ss],
                   measurement, device=device)

    def trajectory(self, x, localization_error):
        return TrajectoryHandle(self, x, localization_error)

    # ------------------------------------------------------------------ batched likelihood
    def logl_runs(self, traj, starts, run_states):
        starts = np.ascontiguousarray(starts, dtype=np.int32)
        run_states = np.ascontiguousarray(run_states, dtype=np.uint8)
        if starts.ndim != 2 or starts.shape != run_states.shape:
            raise ValueError("starts and run_states must both have shape (P, K1)")
        P, K1 = starts.shape
        out = np.empty(P, dtype=np.float64)
        if P:
            _lib.check(_lib.load().bildk_logl_runs(traj._h, P, K1, ptr(starts, c_int32_p), ptr(run_states, c_uint8_p),
                                                   ptr(out, c_double_p)))
        return out

    def logl_st(self, traj, ss, thetas):
        """Batched ``logL(st2profile(s, theta), traj)`` (amis.py:717-739); the (s, theta) -> run-length conversion
        happens inside the library with numpy's exact arithmetic (`st_to_runs` is the Python statement of it)."""
        ss = np.ascontiguousarray(ss, dtype=np.float64)
        thetas = np.ascontiguousarray(thetas, dtype=np.int64)
        if ss.ndim == 1:
            ss, thetas = ss[None, :], thetas[None, :]
        if ss.ndim != 2 or ss.shape != thetas.shape:
            raise ValueError("ss and thetas must both have shape (P, K1)")
        P, K1 = ss.shape
        out = np.empty(P, dtype=np.float64)
        if P:
            _lib.check(_lib.load().bildk_logl_st(traj._h, P, K1, ptr(ss, c_double_p), ptr(thetas, _lib.c_int64_p),
                                                 ptr(out, c_double_p)))
        return out

    def logl_states(self, traj, states):
        """Per-frame state arrays (P, T) or (T,) -> (P,) log-likelihoods."""
        states = np.ascontiguousarray(states, dtype=np.int32)
        if states.ndim == 1:
            states = states[None, :]
        if states.shape[1] != traj.T:
            raise ValueError(f"profile length {states.shape[1]} does not match trajectory length {traj.T}")
        out = np.empty(states.shape[0], dtype=np.float64)
        if len(out):
            _lib.check(_lib.load().bildk_logl_states(traj._h, states.shape[0], ptr(states, c_int32_p), ptr(out, c_double_p)))
        return out

    def logl_runs_multi(self, trajs, offsets, starts, run_states):
        """Profiles [offsets[i], offsets[i+1]) belong to trajs[i]; one launch for the whole dataset."""
        starts = np.ascontiguousarray(starts, dtype=np.int32)
        run_states = np.ascontiguousarray(run_states, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        if len(offsets) != len(trajs) + 1 or starts.shape != run_states.shape or starts.shape[0] != offsets[-1]:
            raise ValueError("inconsistent multi-trajectory batch")
        out = np.empty(starts.shape[0], dtype=np.float64)
        if len(out):
            arr = (ctypes.c_void_p * len(trajs))(*[t._h for t in trajs])
            _lib.check(_lib.load().bildk_logl_runs_multi(len(trajs), arr, ptr(offsets, c_int32_p), starts.shape[1],
                                                         ptr(starts, c_int32_p), ptr(run_states, c_uint8_p), ptr(out, c_double_p)))
        return out

    def logl_runs_multi_submit(self, trajs, offsets, starts, run_states, amis=None):
        """Asynchronous `logl_runs_multi`: stages and enqueues the batch (nothing waits for the GPU) and returns a
        `PendingBatch`; ``.wait()`` returns the (P,) log-likelihoods.  At most two batches in flight per model.
        ``amis``: per trajectory a `FusedAmisStep` (`AmisEnsemble.fused_step`) or None - the AMIS bookkeeping of that
        trajectory's batch is enqueued behind the filter kernel in the same stream."""
        starts = np.ascontiguousarray(starts, dtype=np.int32)
        run_states = np.ascontiguousarray(run_states, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        if len(offsets) != len(trajs) + 1 or starts.shape != run_states.shape or starts.shape[0] != offsets[-1]:
            raise ValueError("inconsistent multi-trajectory batch")
        out = np.empty(starts.shape[0], dtype=np.float64)
        ticket = ctypes.c_void_p()
        reqs, steps = None, []
        if amis is not None and any(a is not None for a in amis):
            if len(amis) != len(trajs):
                raise ValueError("one AMIS entry (or None) per trajectory")
            reqs = (_lib.AmisReq * len(trajs))()
            for i, a in enumerate(amis):
                if a is None:
                    continue
                if len(a.ss) != offsets[i + 1] - offsets[i]:
                    raise ValueError("AMIS step and likelihood batch differ in size")
                a.fill(reqs[i])
                steps.append(a)
        if len(out):
            arr = (ctypes.c_void_p * len(trajs))(*[t._h for t in trajs])
            _lib.check(_lib.load().bildk_logl_runs_multi_submit(len(trajs), arr, ptr(offsets, c_int32_p), starts.shape[1],
                                                                ptr(starts, c_int32_p), ptr(run_states, c_uint8_p), ptr(out, c_double_p),
                                                                reqs, ctypes.byref(ticket)))
            for a in steps:
                a.mark_submitted()
        return PendingBatch(ticket, out, (trajs, steps))

    def logl_runs_device(self, traj, P, K1, d_starts, d_states, d_out, stream=0):
        """Device pointers (ints); asynchronous on ``stream``."""
        _lib.check(_lib.load().bildk_logl_runs_device(traj._h, int(P), int(K1), ctypes.c_void_p(d_starts), ctypes.c_void_p(d_states),
                                                      ctypes.c_void_p(d_out), ctypes.c_void_p(stream)))
