"""
GPU tests of the drop-in API: ``bild_b200.sample`` with the real engine against the run recorded from the
unmodified reference (BASELINE.json configs[0]), the device AMIS-weight reduction against the oracle's
restatement of amis.py:843-900, and the batched postproc caller.
"""
import os

import numpy as np
import pytest

import kalman_oracle as ko
import bild_b200 as bild
from test_amis_host import check_c1_evidence, replay_reference_batches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def runs(golden_dir):
    return np.load(os.path.join(golden_dir, "sample_runs.npz"))


def test_models_logL_reference_fixture():
    """tests/test_bild.py:135-148 through the GPU: range, error-source equivalence, ValueError."""
    traj = bild.Trajectory(np.array([1, 2, np.nan, 4]), localization_error=[0.5])
    profile = bild.Loopingprofile([1, 1, 0, 0])
    model = bild.models.MultiStateRouse(20, 1, 5, d=1)
    v = model.logL(profile, traj)
    assert -100 < v < 0
    assert abs(v - (-10.2226225359098)) < 1e-9 * 10.3
    model2 = bild.models.MultiStateRouse(20, 1, 5, d=1, localization_error=0.5)
    assert model2.logL(profile, traj) == v
    traj.localization_error = None
    with pytest.raises(ValueError):
        model.logL(profile, traj)
    assert np.array_equal(model2.initial_loopingprofile(traj).state, [1, 0, 0, 0])


@pytest.fixture(scope="module")
def replay(golden_dir):
    return np.load(os.path.join(golden_dir, "sample_c1_replay.npz"))


def test_replay_reference_batches(replay):
    """The reference's own 142 profile batches through the GPU engine and the device weight reduction: likelihoods, weights,
    evidence, standard error, KL and refitted proposals within 1e-9 of the reference after every AMIS step."""
    model = bild.models.MultiStateRouse(20, 1, 5, d=3, localization_error=0.3)
    worst = replay_reference_batches(model, replay, tol=1e-9)
    assert max(w["logL"] for w in worst.values()) < 1e-12
    from bild_b200 import _lib
    assert _lib.load().bildk_launch_count() >= 142


def test_sample_config1_matches_reference_run(runs, replay):
    model = bild.models.MultiStateRouse(20, 1, 5, d=3, localization_error=0.3)
    traj = bild.Trajectory(runs["c1_x"], localization_error=[0.3] * 3)
    np.random.seed(1234)
    res = bild.sample(traj, model)
    assert np.array_equal(res.k, runs["c1_k"])
    assert np.array_equal(res.log["k"], runs["c1_logk"])
    check_c1_evidence(res, runs, replay)
    # every likelihood batch went through the CUDA library
    from bild_b200 import _lib
    assert _lib.load().bildk_launch_count() >= int(runs["c1_n_logl_batches"])


def test_marginal_posterior_kernel(runs):
    """Device marginal posterior (amis.py:942-972 replacement) vs the host tensor formulation, and vs the reference run."""
    model = bild.models.MultiStateRouse(20, 1, 5, d=3, localization_error=0.3)
    traj = bild.Trajectory(runs["c1_x"], localization_error=[0.3] * 3)
    np.random.seed(1234)
    res = bild.sample(traj, model)
    np.testing.assert_allclose(np.exp(res.log_marginal_posterior()), np.exp(runs["c1_post"]), atol=1e-4)   # as in test_amis_host
    for smp in res.samplers:
        if not smp.samples:
            continue
        dev = smp.log_marginal_posterior()
        # host formulation: hide the device method from the sampler
        class HostOnly:
            nStates = model.nStates
        keep, smp.model = smp.model, HostOnly()
        try:
            host = smp.log_marginal_posterior()
        finally:
            smp.model = keep
        assert dev.shape == host.shape == (2, len(traj))
        both = np.isfinite(host)
        assert np.array_equal(both, np.isfinite(dev))
        assert np.max(np.abs(dev[both] - host[both])) < 1e-10
    # evidence-averaged posterior goes through the same kernel
    avg = res.log_marginal_posterior(dE="average")
    np.testing.assert_allclose(np.exp(avg).sum(axis=0), 1.0, atol=1e-12)
    # synthetic ensemble with empty runs, padding runs and -inf weights, three states
    rng = np.random.default_rng(5)
    n, K1, T = 777, 6, 41
    ss = rng.dirichlet(np.ones(K1), n)
    ss[::7, 2] = 0
    ss /= ss.sum(axis=1, keepdims=True)
    thetas = rng.integers(0, 2, size=(n, K1))
    lw = rng.normal(-50, 8, n)
    lw[::11] = -np.inf
    got = model.marginal_posterior(ss, thetas, T, lw)
    switches = np.floor(np.cumsum(ss, axis=1)[:, :-1] * (T - 1)).astype(int) + 1
    run = np.sum(np.arange(T)[None, :, None] >= switches[:, None, :], axis=2)
    st = np.take_along_axis(thetas, run, axis=1)
    w = np.exp(lw - lw[np.isfinite(lw)].max())
    want = np.array([[w[st[:, t] == s].sum() for t in range(T)] for s in range(2)])
    want = np.log(want / want.sum(axis=0, keepdims=True))
    assert np.max(np.abs(got - want)) < 1e-10


def test_amis_weights_kernel():
    rng = np.random.default_rng(0)
    model = bild.models.MultiStateRouse(8, 1, 5, d=1, localization_error=0.5)
    for n in (1, 2, 31, 1000, 20000):
        logL = rng.normal(-400, 30, size=n)
        logd = rng.normal(-30, 5, size=n)
        cur = rng.normal(-30, 5, size=n)
        if n > 10:
            cur[3] = -np.inf            # weight-zero sample, impossible under the new proposal: 0 * inf = NaN, skipped
            logd[3] = np.inf
            logd[5] = np.inf
        want_lw, logev, dlogev, KL = ko.amis_evidence(logL, logd, cur, 7, -3.3) if n > 1 else (logL - logd + np.log(7), None, None, None)
        lw, (mx, s1, ssd, s3) = model.amis_weights(logL, logd, cur, np.log(7))
        assert np.array_equal(lw, want_lw)
        if n > 1:
            ev = s1 / n
            assert abs(np.log(ev) + mx - 3.3 - logev) < 1e-12 * abs(logev)
            sem = np.sqrt(ssd / (n - 1)) / np.sqrt(n)
            assert abs(sem / ev - dlogev) < 1e-9 * dlogev
            assert abs(s3 / n / ev - (np.log(ev) + mx - 3.3) - 3.3 - KL) < 1e-9 * max(1, abs(KL))
    a = model.amis_weights(logL, logd, cur, np.log(7))
    b = model.amis_weights(logL, logd, cur, np.log(7))
    assert a[1] == b[1]                 # fixed-order reduction: bitwise reproducible


def test_postproc_batched_on_gpu():
    from bild_b200 import postproc
    model = bild.models.MultiStateRouse(10, 1, 5, d=2, localization_error=0.2)
    np.random.seed(4)
    truth = bild.Loopingprofile([0] * 12 + [1] * 14 + [0] * 10)
    traj = model.trajectory_from_loopingprofile(truth)
    start = bild.Loopingprofile([0] * 10 + [1] * 18 + [0] * 8)
    lr = postproc.logLR_boundaries(start, traj, model)
    base = model.logL(start, traj)
    moved = start.copy()
    moved[28] = 1
    assert abs(lr[1, 1] - (model.logL(moved, traj) - base)) < 1e-9
    opt = postproc.optimize_boundary(start, traj, model)
    assert model.logL(opt, traj) >= base and opt.count_switches() == 2


def test_postproc_on_gpu_against_oracle():
    """postproc.py:13-117 with the engine vs the SAME host code with the C oracle as likelihood: every log-likelihood ratio
    of every pass within 1e-9 of the oracle's, identical sequence of boundary moves, identical optimum - on a profile
    with several boundaries, missing frames and N = 20 (k_mmar) / N = 50 (k_mma2)."""
    from bild_b200 import postproc
    from test_amis_host import OracleBackedRouse
    for N, T, seed in ((20, 120, 6), (50, 90, 7)):
        gpu = bild.models.MultiStateRouse(N, 1, 5, d=3, localization_error=0.3)
        cpu = OracleBackedRouse(N, 1, 5, d=3, localization_error=0.3)
        np.random.seed(seed)
        edges = np.sort(np.random.choice(np.arange(8, T - 8), 5, replace=False))
        truth = np.zeros(T, dtype=int)
        for i, e in enumerate(edges):
            truth[e:] = (i + 1) % 2
        traj = gpu.trajectory_from_loopingprofile(bild.Loopingprofile(truth), missing_frames=0.1)
        start = truth.copy()                                  # move every boundary by a frame or two
        for e, shift in zip(edges, (2, -1, 1, -2, 1)):
            if shift > 0:
                start[e:e + shift] = truth[e - 1]            # later
            else:
                start[e + shift:e] = truth[e]                # earlier
        start = bild.Loopingprofile(start)
        a, b = postproc.logLR_boundaries(start, traj, gpu), postproc.logLR_boundaries(start, traj, cpu)
        assert a.shape == b.shape and a.shape[1] == 2 and len(a) >= 3
        assert np.max(np.abs(a - b)) < 1e-9 * max(1.0, abs(cpu.logL(start, traj)))
        opt_gpu, opt_cpu = postproc.optimize_boundary(start, traj, gpu), postproc.optimize_boundary(start, traj, cpu)
        assert opt_gpu == opt_cpu
        assert abs(gpu.logL(opt_gpu, traj) - cpu.logL(opt_cpu, traj)) < 1e-9 * abs(cpu.logL(opt_cpu, traj))
        assert cpu.logL(opt_cpu, traj) >= cpu.logL(start, traj)


def test_sample_many_on_gpu_equals_one_by_one():
    """Dataset driver with the real engine: fused multi-trajectory launches == sequential bild.sample runs."""
    from bild_b200.dataset import sample_many
    model = bild.models.MultiStateRouse(12, 1, 5, d=3, localization_error=0.3)
    np.random.seed(3)
    trajs = [model.trajectory_from_loopingprofile(bild.Loopingprofile([0] * a + [1] * b + [0] * c))
             for a, b, c in [(10, 12, 8), (15, 15, 0), (7, 6, 12)]]
    kw = dict(init_runs=5, sampler_kw={"N": 40}, k_max=5)
    seeds = [11, 12, 13]
    res, stats = sample_many(trajs, model, seeds=seeds, **kw)
    assert stats["launches"] == stats["rounds"] > 0
    for i, tr in enumerate(trajs):
        np.random.seed(seeds[i])
        solo = bild.sample(tr, model, **kw)
        assert np.array_equal(solo.log["k"], res[i].log["k"])
        assert np.array_equal(solo.evidence, res[i].evidence)       # bitwise: same kernels, same random numbers


def test_device_ensemble_matches_host_bookkeeping():
    """`bildk_amis_step` (device-resident ensemble: mixture denominators, weights, evidence sums, Dirichlet moments, CFC
    marginals in one launch) against the host numpy bookkeeping of the same sampler, step by step, incl. the +inf
    conventions for rejected samples (amis.py:98-108) and three states."""
    from bild_b200 import amis
    model = bild.models.MultiStateRouse(12, 1, 5, d=2, looppositions=(None, (0, -1), (2, 8)), localization_error=0.3)
    np.random.seed(8)
    truth = bild.Loopingprofile([0] * 15 + [1] * 15 + [2] * 15 + [0] * 15)
    traj = model.trajectory_from_loopingprofile(truth, missing_frames=0.1)
    for k in (2, 5, 17):                                      # 17: K1 = 18 > 16 -> the 32-column kernel variant
        dev = amis.FixedkSampler(traj, model, k=k, N=64)
        host = amis.FixedkSampler(traj, model, k=k, N=64)
        host._ens = None                                      # force the numpy bookkeeping
        assert dev._device_ensemble() is not None
        for step in range(6):
            np.random.seed(100 * k + step)
            cur = dev.parameters[-1]                          # ONE batch, drawn from the device sampler's proposal, fed to both
            ss, th = dev.dirichlet.sample(cur[0], dev.N), dev.cfc.sample(cur[1], dev.N)
            if step == 2:                                     # a rejected sample: zero-length interval, and one off the simplex
                ss[0, 0] += ss[0, 1]; ss[0, 1] = 0.0
                ss[1] *= 1.001
            for smp in (dev, host):
                smp.dirichlet.sample = lambda a, N, ss=ss: ss.copy()
                smp.cfc.sample = lambda logp, N, th=th: th.copy()
                assert smp.step() is True
                del smp.dirichlet.sample, smp.cfc.sample
            np.testing.assert_allclose(dev.evidences[-1], host.evidences[-1], rtol=1e-10, atol=1e-10)
            np.testing.assert_allclose(dev.parameters[-1][0], host.parameters[-1][0], rtol=1e-9)
            np.testing.assert_allclose(np.exp(dev.parameters[-1][1]), np.exp(host.parameters[-1][1]), atol=1e-12)
        for key in ("log_weights", "logδs", "cur_log_proposal"):
            a = np.concatenate([b[key] for b in dev.samples])
            b = np.concatenate([b[key] for b in host.samples])
            assert np.array_equal(np.isfinite(a), np.isfinite(b)) and np.array_equal(a[~np.isfinite(a)], b[~np.isfinite(b)])
            ok = np.isfinite(a)
            np.testing.assert_allclose(a[ok], b[ok], rtol=1e-11, atol=1e-11)
        assert np.isposinf(np.concatenate([b["logδs"] for b in dev.samples])).sum() >= 1
        np.testing.assert_allclose(np.exp(dev.log_marginal_posterior()), np.exp(host.log_marginal_posterior()), atol=1e-10)


def test_fused_amis_steps_match_synchronous_steps_bitwise():
    """`bildk_logl_runs_multi_submit` with `amis` requests (the bookkeeping of several samplers riding on ONE likelihood
    launch, two AMIS launches for all of them) against `bildk_amis_step` called sampler by sampler: identical bits in
    the statistics and in every per-sample array, for both column-width classes (K1 <= 16 and K1 = 18) in one batch,
    three states, and a rejected sample."""
    from bild_b200 import amis
    model = bild.models.MultiStateRouse(12, 1, 5, d=2, looppositions=(None, (0, -1), (2, 8)), localization_error=0.3)
    np.random.seed(9)
    truth = bild.Loopingprofile([0] * 15 + [1] * 15 + [2] * 15 + [0] * 15)
    trajs = [model.trajectory_from_loopingprofile(truth, missing_frames=0.1) for _ in range(3)]
    ks = (17, 2, 5)                                           # mixed order: the library sorts the width classes itself
    fused = [amis.FixedkSampler(t, model, k=k, N=48 + 8 * i) for i, (t, k) in enumerate(zip(trajs, ks))]
    solo = [amis.FixedkSampler(t, model, k=k, N=48 + 8 * i) for i, (t, k) in enumerate(zip(trajs, ks))]
    for step in range(5):
        gens, reqs = [], []
        for i, smp in enumerate(fused):
            np.random.seed(1000 * step + i)
            g = smp.step_gen()
            req = next(g)
            if step == 2 and i == 1:
                req[0][0, 0] += req[0][0, 1]; req[0][0, 1] = 0.0      # a rejected sample (zero-length interval)
            gens.append(g)
            reqs.append(req)
        runs = [bild.engine.st_to_runs(r[0], r[1], len(t)) for r, t in zip(reqs, trajs)]
        K1 = max(a.shape[1] for a, _ in runs)
        starts = np.concatenate([np.concatenate([a, np.full((len(a), K1 - a.shape[1]), len(t), dtype=a.dtype)], axis=1)
                                 for (a, _), t in zip(runs, trajs)])
        states = np.concatenate([np.concatenate([b, np.repeat(b[:, -1:], K1 - b.shape[1], axis=1)], axis=1) for _, b in runs])
        offsets = np.concatenate([[0], np.cumsum([len(a) for a, _ in runs])])
        batch = model.logL_runs_multi_submit(trajs, offsets, starts, states, amis=[r.amis for r in reqs])
        out = batch.wait()
        for i, (g, smp) in enumerate(zip(gens, fused)):
            assert reqs[i].amis.submitted
            try:
                g.send(out[offsets[i]:offsets[i + 1]])
            except StopIteration as stop:
                assert stop.value is True
            # the same batch through the synchronous call
            other = solo[i]
            other.dirichlet.sample = lambda a, N, ss=reqs[i][0]: ss.copy()
            other.cfc.sample = lambda logp, N, th=reqs[i][1]: th.copy()
            assert other.step() is True
            del other.dirichlet.sample, other.cfc.sample
            assert np.array_equal(smp.evidences[-1], other.evidences[-1])
            assert np.array_equal(smp.parameters[-1][0], other.parameters[-1][0])
            assert np.array_equal(smp.parameters[-1][1], other.parameters[-1][1])
            for key in ("logLs", "log_weights", "logδs", "cur_log_proposal"):
                a = np.concatenate([b[key] for b in smp.samples])
                b = np.concatenate([b[key] for b in other.samples])
                assert np.array_equal(a, b, equal_nan=True), (step, i, key)
