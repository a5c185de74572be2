"""
ORACLE - TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by IMPORTING THE UNMODIFIED REFERENCE
from /root/reference (it cannot travel to the GPU box, the vectors can).

Run here (CPU container):   python oracle/make_golden.py

What is executed from the reference:
  * bild.models.MultiStateRouse / bild.Loopingprofile            (/root/reference/bild/models.py, util.py)
  * bild.src.MSRouse_logL_py.MSRouse_logL  (pure-Python twin)    -> ``logL_py``
  * oracle/_ref/MSRouse_logL*.so = the reference's .pyx compiled by oracle/Makefile -> ``logL_cy``
  * bild.amis.FixedkSampler.st2profile, Dirichlet.sample, CFC.sample  (profile batches)
The three third-party packages the reference imports but does not vendor (rouse, noctiluca,
bayesmsd) are provided by oracle/shims/ - so the PROPAGATORS in these vectors follow
oracle/rouse_oracle.py ("parity unpinned" for that part, see its header); the filter arithmetic
on top of them is the reference's own code.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    import bild                                        # the reference package, unmodified
    from bild.src.MSRouse_logL_py import MSRouse_logL as ref_logL_py
import noctiluca as nl                                  # shim
import kalman_oracle as ko

ref_logL_cy = ko.ref_cython()
assert ref_logL_cy is not None, "run `make -C oracle ref` first"
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def pack(model, traj, profiles):
    """Evaluate the reference on every profile and collect raw arrays for the parity tests."""
    Bs, Gs, Sigs, M0, C0 = ko.model_arrays(model.models)
    s2, Cind = ko.noise_to_s2_cind(model._get_noise(traj))
    states = np.array([p[:] for p in profiles], dtype=np.int32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lpy = np.array([ref_logL_py(model, p, traj) for p in profiles])
        lcy = np.array([ref_logL_cy(model, p, traj) for p in profiles])
    return dict(Bs=Bs, Gs=Gs, Sigs=Sigs, M0=M0, C0=C0, w=np.asarray(model.measurement, dtype=float),
                x=np.asarray(traj[:], dtype=float), s2=s2, Cind=Cind, states=states,
                logL_py=lpy, logL_cy=lcy)


def amis_profiles(model, traj, rng_seed, ks, per_k):
    """Profile batches the way FixedkSampler.step draws them (amis.py:830-833) for several k."""
    np.random.seed(rng_seed)
    out, sts = [], []
    for k in ks:
        smp = bild.amis.FixedkSampler.__new__(bild.amis.FixedkSampler)
        smp.traj, smp.model, smp.k = traj, model, k
        cfc = bild.amis.CFC(model.transitions)
        ss = bild.amis.Dirichlet().sample(np.ones(k + 1), per_k)
        th = cfc.sample(cfc.logp_uniform(k), per_k)
        for s, t in zip(ss, th):
            out.append(smp.st2profile(s, t))
            sts.append((s, t))
    return out, sts


def main():
    cases = {}

    # (1) the reference's own fixture: tests/test_bild.py:125-136
    traj = nl.Trajectory(np.array([1, 2, np.nan, 4]), localization_error=[0.5])
    model = bild.models.MultiStateRouse(20, 1, 5, d=1)
    cases["fixture_test_bild"] = pack(model, traj, [bild.Loopingprofile([1, 1, 0, 0]),
                                                     bild.Loopingprofile([0, 0, 0, 0]),
                                                     bild.Loopingprofile([0, 1, 0, 1])])

    # (2) config-1-like: N=20, d=3, T=100, 10% missing, AMIS-style profile batch
    np.random.seed(685441950)                           # tests/test_bild.py:9
    model = bild.models.MultiStateRouse(20, 1, 5, d=3, localization_error=0.3)
    truth = bild.Loopingprofile((np.arange(100) // 20) % 2)
    traj = model.trajectory_from_loopingprofile(truth, missing_frames=0.1)
    profs, _ = amis_profiles(model, traj, 1, ks=[0, 1, 2, 3, 5, 10], per_k=4)
    cases["n20_d3_t100_nan"] = pack(model, traj, profs)

    # (3) anisotropic localisation error -> d* = 2 (pyx:145)
    model = bild.models.MultiStateRouse(10, 1, 5, d=3, localization_error=[0.1, 0.1, 0.3])
    truth = bild.Loopingprofile((np.arange(50) // 10) % 2)
    traj = model.trajectory_from_loopingprofile(truth, missing_frames=0.2)
    profs, _ = amis_profiles(model, traj, 2, ks=[0, 2, 4], per_k=4)
    cases["n10_aniso_dstar2"] = pack(model, traj, profs)

    # (4) larger polymer N=50, short T; error taken from the trajectory, not the model
    model = bild.models.MultiStateRouse(50, 1, 5, d=3)
    gen = bild.models.MultiStateRouse(50, 1, 5, d=3, localization_error=0.2)
    truth = bild.Loopingprofile((np.arange(60) // 15) % 2)
    traj = gen.trajectory_from_loopingprofile(truth, missing_frames=5)
    profs, _ = amis_profiles(model, traj, 3, ks=[1, 3], per_k=3)
    cases["n50_t60_trajerr"] = pack(model, traj, profs)

    # (5) first frame missing / long gap / only one valid frame
    model = bild.models.MultiStateRouse(12, 1, 5, d=2, localization_error=0.25)
    truth = bild.Loopingprofile([0] * 10 + [1] * 10)
    traj = model.trajectory_from_loopingprofile(truth)
    traj.data[0, :] = np.nan
    traj.data[5:14, 0] = np.nan                          # one NaN component invalidates the frame
    profs, _ = amis_profiles(model, traj, 4, ks=[0, 1, 3], per_k=3)
    cases["n12_first_missing_gap"] = pack(model, traj, profs)
    traj1 = nl.Trajectory(traj.data.copy(), localization_error=[0.25, 0.25])
    traj1.data[:, :] = np.nan
    traj1.data[7, :] = [0.3, -0.2]
    cases["n12_single_valid"] = pack(model, traj1, profs[:4])

    # (6) three states, odd N, dense (non end-to-end) measurement vector, unequal loop strength
    w = np.linspace(-1, 1, 17)
    w -= w.mean()
    model = bild.models.MultiStateRouse(17, 0.7, 2.5, d=3, looppositions=(None, (0, -1), [(3, 12, 2.0), (1, 5)]),
                                        measurement=w, localization_error=0.15)
    truth = bild.Loopingprofile(np.repeat([0, 2, 1, 0], 10))
    traj = model.trajectory_from_loopingprofile(truth, missing_frames=0.15)
    profs, _ = amis_profiles(model, traj, 5, ks=[0, 2, 5], per_k=4)
    cases["n17_three_states_dense_w"] = pack(model, traj, profs)

    for name, c in cases.items():
        d = np.max(np.abs(c["logL_py"] - c["logL_cy"]) / np.maximum(1.0, np.abs(c["logL_cy"])))
        print(f"{name:28s} P={len(c['states']):3d}  logL[0]={c['logL_cy'][0]: .15g}  max rel |py-cy|={d:.2e}")
        np.savez_compressed(os.path.join(OUT, f"logl_{name}.npz"), **c)

    # (7) st2profile golden (amis.py:670-695), incl. the probe cases of SURVEY.md appendix A
    smp = bild.amis.FixedkSampler.__new__(bild.amis.FixedkSampler)
    rows = []
    hand = [([.25, .5, .25], [0, 1, 0], 6), ([0, .5, .5], [0, 1, 0], 6), ([.5, 0, .5], [0, 1, 0], 6),
            ([.5, .5, 0], [0, 1, 0], 6), ([.999, .0005, .0005], [0, 1, 0], 6), ([1.0], [1], 5)]
    np.random.seed(7)
    for T in (2, 3, 7, 100, 501):
        for k in (0, 1, 2, 5, 9):
            for _ in range(6):
                s = np.random.dirichlet(np.ones(k + 1) * np.random.choice([0.05, 1.0, 20.0]))
                th = np.zeros(k + 1, dtype=int)
                th[0] = np.random.randint(3)
                for i in range(1, k + 1):
                    th[i] = (th[i - 1] + 1 + np.random.randint(2)) % 3
                hand.append((s, th, T))
    KMAX = 10
    S_arr = np.zeros((len(hand), KMAX)); TH = np.zeros((len(hand), KMAX), dtype=np.int32)
    K1 = np.zeros(len(hand), dtype=np.int32); TT = np.zeros(len(hand), dtype=np.int32)
    for i, (s, th, T) in enumerate(hand):
        smp.traj = np.zeros((T, 1))
        st = smp.st2profile(np.asarray(s, dtype=float), np.asarray(th))[:]
        K1[i], TT[i] = len(s), T
        S_arr[i, :len(s)] = s; TH[i, :len(s)] = th
        rows.append(np.asarray(st, dtype=np.int32))
    flat = np.concatenate(rows)
    np.savez_compressed(os.path.join(OUT, "st2profile.npz"), ss=S_arr, thetas=TH, k1=K1, T=TT,
                        states_flat=flat, offsets=np.cumsum([0] + [len(r) for r in rows]))
    print("st2profile cases:", len(hand))


def sample_goldens():
    """Full ``bild.sample`` runs of the unmodified reference under fixed seeds (host AMIS parity)."""
    import scipy.stats
    out = {}
    # (a) FactorizedModel - the model the reference's own TestCore uses (tests/test_bild.py:224-283)
    for seed in (1, 2, 3):
        np.random.seed(seed)
        model = bild.models.FactorizedModel([scipy.stats.maxwell(scale=1), scipy.stats.maxwell(scale=4)], d=1)
        traj = nl.Trajectory(np.array([0.5, 0.7, 3.0, 5.0, 4.0, 0.8, 0.9, 6.0, 5.5, 0.3, 0.2, 4.0]))
        res = bild.sample(traj, model, sampler_kw={"N": 50}, dE=1.0)
        out[f"fact{seed}_k"] = res.k
        out[f"fact{seed}_evidence"] = res.evidence
        out[f"fact{seed}_evidence_se"] = res.evidence_se
        out[f"fact{seed}_logk"] = res.log["k"]
        out[f"fact{seed}_best"] = res.best_profile()[:]
        out[f"fact{seed}_post"] = res.log_marginal_posterior()
        out[f"fact{seed}_post_avg"] = res.log_marginal_posterior(dE="average")
    out["fact_data"] = np.array([0.5, 0.7, 3.0, 5.0, 4.0, 0.8, 0.9, 6.0, 5.5, 0.3, 0.2, 4.0])

    # (b) BASELINE.json configs[0]: MultiStateRouse N=20, d=3, T=100, localisation error, defaults
    import bild.models as bm
    bm.MSRouse_logL = ref_logL_cy                       # the reference's own compiled .pyx (its plugin slot)
    np.random.seed(685441950)
    model = bild.models.MultiStateRouse(20, 1, 5, d=3, localization_error=0.3)
    truth = (np.arange(100) // 20) % 2
    traj = model.trajectory_from_loopingprofile(bild.Loopingprofile(truth))
    np.random.seed(1234)
    res = bild.sample(traj, model)
    replay_golden(res, traj, truth)
    out["c1_x"] = traj[:]
    out["c1_truth"] = truth
    out["c1_k"] = res.k
    out["c1_evidence"] = res.evidence
    out["c1_evidence_se"] = res.evidence_se
    out["c1_logk"] = res.log["k"]
    out["c1_best"] = res.best_profile()[:]
    out["c1_post"] = res.log_marginal_posterior()
    out["c1_n_logl_batches"] = np.array(sum(len(smp.samples) for smp in res.samplers))
    print("C1: k", res.k, "best k", res.best_k(), "steps", len(res.log["k"]),
          "profile recovered:", bool(np.all(res.best_profile()[:] == truth)))
    np.savez_compressed(os.path.join(OUT, "sample_runs.npz"), **out)


def replay_golden(res, traj, truth):
    """
    Everything needed to replay the config-1 reference run step by step on IDENTICAL profile batches
    (tests/test_gpu_sample.py::test_replay_reference_batches): per sampler the concatenated samples in draw order
    (``ss``, ``thetas``, ``logLs``, and - for the AMIS samplers - ``logδs``, ``cur_log_proposal``, ``log_weights`` as
    they stand at the END of the run), the batch sizes, the evidence triple after every step (amis.py:878-900) and
    the proposal parameters after every step (amis.py:847-876).
    """
    rp = {"x": traj[:], "truth": truth, "n_samplers": np.array(len(res.samplers))}
    for smp in res.samplers:
        k = smp.k
        rp[f"k{k}_sizes"] = np.array([len(b["logLs"]) for b in smp.samples])
        rp[f"k{k}_ss"] = np.concatenate([b["ss"] for b in smp.samples])
        rp[f"k{k}_thetas"] = np.concatenate([b["thetas"] for b in smp.samples]).astype(np.int8)
        rp[f"k{k}_logLs"] = np.concatenate([b["logLs"] for b in smp.samples])
        rp[f"k{k}_evidences"] = np.array(smp.evidences, dtype=float)
        rp[f"k{k}_exhausted"] = np.array(bool(smp.exhausted))
        if "log_weights" in smp.samples[-1]:
            for key, name in (("logδs", "logdeltas"), ("cur_log_proposal", "cur_log_proposal"), ("log_weights", "log_weights")):
                rp[f"k{k}_{name}"] = np.concatenate([b[key] for b in smp.samples])
            rp[f"k{k}_par_a"] = np.array([p[0] for p in smp.parameters])
            rp[f"k{k}_par_logp"] = np.array([p[1] for p in smp.parameters])
    rp["post"] = res.log_marginal_posterior()
    rp["logk"] = res.log["k"]
    np.savez_compressed(os.path.join(OUT, "sample_c1_replay.npz"), **rp)
    print("replay golden:", sum(int(np.sum(rp[f"k{s.k}_sizes"])) for s in res.samplers), "profiles in",
          sum(len(rp[f"k{s.k}_sizes"]) for s in res.samplers), "batches")


if __name__ == "__main__":
    if "--sample-only" not in sys.argv:
        main()
    sample_goldens()
