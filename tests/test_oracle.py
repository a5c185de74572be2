"""
CPU tests of the ORACLE itself: the C and numpy restatements against the golden vectors that
oracle/make_golden.py produced by running the unmodified reference (pure-Python twin and compiled
.pyx), and against an independent dense-Gaussian evaluation.
"""
import os

import numpy as np
import pytest

import kalman_oracle as ko
import rouse_oracle as ro
from helpers import MODEL_KEYS, golden_cases, load_golden, oracle_model, rel_err, synth_traj

CPU_TOL = 1e-12   # CPU-vs-CPU: observed <= 3e-15


def args_of(g):
    return [g[k] for k in MODEL_KEYS] + [g["x"], g["s2"], g["Cind"]]


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[5:-4])
def test_c_oracle_matches_reference_golden(path):
    g = load_golden(path)
    got = ko.logl_c(*args_of(g), g["states"])
    assert rel_err(got, g["logL_cy"]) < CPU_TOL     # reference .pyx, compiled
    assert rel_err(got, g["logL_py"]) < CPU_TOL     # reference pure-Python twin


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[5:-4])
def test_numpy_oracle_matches_reference_golden(path):
    g = load_golden(path)
    got = np.array([ko.logl_numpy(*args_of(g), s) for s in g["states"]])
    assert rel_err(got, g["logL_py"]) < CPU_TOL


def test_reference_fixture_value():
    """tests/test_bild.py:125-138: range check of the reference, plus the restatement-convention value."""
    g = load_golden([p for p in golden_cases() if "fixture_test_bild" in p][0])
    v = g["logL_cy"][0]
    assert -100 < v < 0
    assert abs(v - (-10.2226225359098)) < 1e-10


@pytest.mark.parametrize("path", golden_cases()[:4], ids=lambda p: os.path.basename(p)[5:-4])
def test_dense_gaussian_independent_check(path):
    g = load_golden(path)
    for i in range(min(3, len(g["states"]))):
        dense = ko.logl_dense_gaussian(*args_of(g), g["states"][i])
        assert abs(dense - g["logL_cy"][i]) < 1e-9 * max(1, abs(dense))


def test_compiled_reference_if_present():
    """oracle/_ref is built from /root/reference here and travels to the GPU box; absent elsewhere."""
    fn = ko.ref_cython()
    if fn is None:
        pytest.skip("oracle/_ref not built")
    from bild_b200.models import MultiStateRouse
    from bild_b200.trajectory import Trajectory
    from bild_b200.util import Loopingprofile
    # the reference's own fixture, driven through OUR model/trajectory/profile objects
    traj = Trajectory(np.array([1, 2, np.nan, 4]), localization_error=[0.5])
    model = MultiStateRouse(20, 1, 5, d=1)
    v = fn(model, Loopingprofile([1, 1, 0, 0]), traj)
    assert abs(v - (-10.2226225359098)) < 1e-9
    model2 = MultiStateRouse(20, 1, 5, d=1, localization_error=0.5)
    assert fn(model2, Loopingprofile([1, 1, 0, 0]), traj) == v      # test_bild.py:145-148
    traj.localization_error = None
    with pytest.raises(ValueError):                                  # test_bild.py:140-143
        fn(model, Loopingprofile([1, 1, 0, 0]), traj)


def test_oracle_properties():
    rng = np.random.default_rng(3)
    mod = oracle_model(14, d=3)
    x, _ = synth_traj(mod, 30, rng, 0.3, p_nan=0.2)
    s2, Cind = ko.noise_to_s2_cind([0.3, 0.3, 0.3])
    states = rng.integers(0, 2, size=(5, 30))
    base = ko.logl_c(*[mod[k] for k in MODEL_KEYS], x, s2, Cind, states)
    # NaN frames are skipped but still propagate: deleting the data of a frame == marking it NaN
    x2 = x.copy()
    x2[7] = np.nan
    x3 = x.copy()
    x3[7, 1] = np.nan                                   # one NaN component invalidates the frame (pyx:178)
    a = ko.logl_c(*[mod[k] for k in MODEL_KEYS], x2, s2, Cind, states)
    b = ko.logl_c(*[mod[k] for k in MODEL_KEYS], x3, s2, Cind, states)
    assert np.array_equal(a, b)
    assert not np.allclose(a, base) or np.isnan(x[7]).any()
    # anisotropic errors that happen to be equal give the isotropic answer (d* sharing)
    s2b, Cindb = ko.noise_to_s2_cind([0.3, 0.3 + 1e-13, 0.3])
    c = ko.logl_c(*[mod[k] for k in MODEL_KEYS], x, s2b, Cindb, states)
    assert len(s2b) == 2 and rel_err(c, base) < 1e-9


def test_rouse_restatements_agree():
    """Product propagators (one eigendecomposition) vs oracle (Pade expm + SVD pinv)."""
    from bild_b200 import rouse as pr
    for N, bonds in [(20, None), (20, [(0, -1)]), (50, [(0, -1)]), (17, [(3, 12, 2.0), (1, 5)]), (10, [(2, 3, -1)])]:
        a, b = pr.Model(N, 1.0, 5.0, 3, add_bonds=bonds), ro.Model(N, 1.0, 5.0, 3, add_bonds=bonds)
        for key in ("B", "G", "Sig"):
            assert np.max(np.abs(a._dynamics[key] - b._dynamics[key])) < 1e-11
        for u, v in zip(a.steady_state(), b.steady_state()):
            assert np.max(np.abs(u - v)) < 1e-10
        B, Sig = a._dynamics["B"], a._dynamics["Sig"]
        assert np.array_equal(B, B.T) and np.array_equal(Sig, Sig.T)
        # stationarity: C = B C B + Sig on the non-centre-of-mass subspace
        if np.count_nonzero(np.abs(np.linalg.eigvalsh(a.A)) < 1e-9) == 1:   # connected chain
            _, C = a.steady_state()
            w = np.zeros(N); w[0], w[-1] = -1, 1
            assert abs(w @ (B @ C @ B + Sig - C) @ w) < 1e-11
