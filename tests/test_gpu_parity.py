"""
GPU parity tests: the CUDA path, called through the C ABI (bild_b200.engine -> libbild_b200.so),
against (i) the golden vectors produced by the unmodified reference and (ii) the C oracle on seeded
synthetic inputs.  Gate: 1e-9 relative on logL (BASELINE.json north_star); observed ~1e-13.
"""
import numpy as np
import pytest

import kalman_oracle as ko
from helpers import MODEL_KEYS, golden_cases, load_golden, logl_c_parallel, oracle_model, random_profiles, rel_err, synth_traj

pytestmark = pytest.mark.gpu

TOL = 1e-9   # relative, FP64 (north_star: "within 1e-9 relative on logL")


def engine_for(mod, device=0):
    from bild_b200.engine import RouseEngine
    return RouseEngine(*(mod[k] for k in MODEL_KEYS), device=device)


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("logl_")[-1][:-4])
def test_golden_vectors(path):
    g = load_golden(path)
    eng = engine_for(g)
    err = np.sqrt(g["s2"])[g["Cind"]]
    traj = eng.trajectory(g["x"], err)
    out = eng.logl_states(traj, g["states"])
    assert rel_err(out, g["logL_cy"]) < TOL
    assert rel_err(out, g["logL_py"]) < TOL


CASES = [
    # N, d, T, P, p_nan, noise, loops, kmax
    (20, 3, 100, 64, 0.1, 0.3, (None, [(0, -1)]), 10),          # config 1/2 shape: warp-scope, 2 filters per warp
    (10, 3, 60, 33, 0.0, 0.2, (None, [(0, -1)]), 4),            # N=10
    (25, 3, 80, 20, 0.2, 0.3, (None, [(0, -1)]), 6),            # G=5: CTA-scope packing
    (50, 3, 120, 19, 0.1, 0.3, (None, [(0, -1)]), 8),           # config 4 shape
    (100, 3, 40, 5, 0.1, 0.3, (None, [(0, -1)]), 3),            # config 3 shape: single resident propagator
    (23, 2, 50, 17, 0.3, [0.1, 0.4], (None, [(0, -1)], [(2, 9), (5, 20, 0.5)]), 5),  # odd N, 3 states, d*=2
    (7, 1, 30, 9, 0.0, 0.5, (None, [(0, -1)]), 2),              # tiny
    (3, 3, 25, 6, 0.0, 0.5, (None, [(0, -1)]), 2),              # N == d
    (2, 3, 25, 6, 0.1, 0.5, (None, [(0, -1, 0.5)]), 2),         # N < d: catch-all kernel
    (60, 3, 30, 4, 0.1, 0.3, (None, [(0, -1)]), 4),             # GT=8: k_mmar2 with four warps per filter
    (64, 3, 20, 3, 0.0, 0.3, (None, [(0, -1)]), 3),             # GT=8, r=8: k_mmar2 (four warps), mean in extra rows
    (68, 3, 24, 4, 0.1, 0.3, (None, [(0, -1)]), 3),             # GT=9: k_mmar2 with five warps per filter
    (70, 3, 24, 4, 0.1, 0.3, (None, [(0, -1)]), 3),             # GT=9, r=6: k_mmar2 (five warps), mean in extra rows
    (108, 3, 12, 2, 0.1, 0.3, (None, [(0, -1)]), 2),            # GT=14 (k_mmact without helpers, 7 slots per warp)
    (96, 2, 16, 3, 0.2, 0.3, (None, [(0, -1)], [(10, 50)]), 3), # GT=12, 3 states: k_mmar8, single resident propagator, TMA swaps
    (100, 2, 14, 3, 0.2, [0.2, 0.4], (None, [(0, -1)], [(10, 50)]), 3),   # GT=13: k_mmar8 (eight warps, TMA propagator swaps), 3 states, d*=2
    (110, 3, 12, 2, 0.0, 0.3, (None, [(0, -1)]), 2),            # GT=14
    (120, 2, 12, 3, 0.1, 0.3, (None, [(0, -1)]), 2),            # GT=15, no padding room: tensor cores with the covariance in L2
    (130, 2, 12, 3, 0.0, 0.3, (None, [(0, -1)]), 2),            # GT=17: beyond the shared-memory limit
    (200, 3, 8, 2, 0.0, 0.3, (None, [(0, -1)], [(20, 150)]), 2),  # BASELINE sweep size N=200, 3 states
]


@pytest.mark.parametrize("N,d,T,P,p_nan,noise,loops,kmax", CASES)
def test_against_c_oracle(N, d, T, P, p_nan, noise, loops, kmax):
    rng = np.random.default_rng(1000 + N * 7 + T)
    mod = oracle_model(N, d=d, loops=loops)
    x, _ = synth_traj(mod, T, rng, noise, p_nan=p_nan)
    ss, thetas = random_profiles(rng, P, T, len(loops), kmax)
    err = np.broadcast_to(np.asarray(noise, dtype=float), (d,))
    s2, Cind = ko.noise_to_s2_cind(err)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)

    eng = engine_for(mod)
    traj = eng.trajectory(x, err)
    got_st = eng.logl_st(traj, ss, thetas)            # run-length path (amis.py:717-739 replacement)
    got_states = eng.logl_states(traj, states)        # per-frame path (models.py:265-278 replacement)
    assert rel_err(got_st, want) < TOL
    assert rel_err(got_states, want) < TOL
    assert np.array_equal(got_st, got_states)          # same filters, same arithmetic


@pytest.mark.parametrize("kernel", ["tile", "mmag", "mmag-one-column", "mmac-column-warps"])
@pytest.mark.parametrize("N", [20, 60, 100])
def test_kernel_variants_agree(kernel, N, monkeypatch):
    """Every kernel family that can run a shape gives the oracle's answer (BILDK_KERNEL forces the family)."""
    if kernel.startswith("mmag") and N < 57:
        pytest.skip("the L2-workspace tensor-core kernel starts at GT = 8")
    if kernel == "mmag-one-column":      # k_mmag (one tile column per warp) instead of k_mmag2 (two)
        monkeypatch.setenv("BILDK_MMAG2", "0")
        kernel = "mmag"
    if kernel == "mmac-column-warps":    # k_mmac with helper warps instead of k_mmact (slots) where GT = 4k + 1
        if N < 57:
            pytest.skip("one CTA per filter starts at GT = 8")
        monkeypatch.setenv("BILDK_MMACT", "0")
        kernel = "mmac"
    rng = np.random.default_rng(N)
    mod = oracle_model(N, d=3)
    T, P = 25, 5
    x, _ = synth_traj(mod, T, rng, 0.3, p_nan=0.1)
    ss, thetas = random_profiles(rng, P, T, 2, 4)
    s2, Cind = ko.noise_to_s2_cind([0.3] * 3)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
    eng = engine_for(mod)
    traj = eng.trajectory(x, [0.3] * 3)
    monkeypatch.setenv("BILDK_KERNEL", kernel)
    assert kernel in traj.describe_plan(P).split()[0]
    got = eng.logl_st(traj, ss, thetas)
    assert rel_err(got, want) < TOL


MMAR_CASES = [
    # N, d, noise, loops                      register-chained kernel: GT <= 4, N mod 8 in 1..4
    (3, 3, 0.5, (None, [(0, -1)])),           # GT=1, r=3
    (4, 4, 0.4, (None, [(0, -1)])),           # GT=1, r=4, d=4: every spare row carries a mean column, no zero row
    (10, 3, 0.2, (None, [(0, -1)])),          # GT=2, r=2 (sweep size N=10)
    (12, 2, [0.1, 0.4], (None, [(0, -1)])),   # GT=2, r=4, d*=2
    (17, 1, 0.3, (None, [(0, -1)])),          # GT=3, r=1, d=1
    (20, 3, 0.3, (None, [(0, -1)], [(2, 9), (5, 15, 0.5)])),   # GT=3, r=4 (configs[1]), 3 states
    (20, 3, [0.2, 0.2, 0.5], (None, [(0, -1)])),               # ... anisotropic error: sub-filters with 2 and 1 columns
    (25, 3, 0.3, (None, [(0, -1)])),          # GT=4, r=1 (sweep size N=25)
    (28, 3, 0.3, (None, [(0, -1)])),          # GT=4, r=4
]


@pytest.mark.parametrize("nb", [0, 1], ids=["default-budget", "smallest-budget"])
@pytest.mark.parametrize("N,d,noise,loops", MMAR_CASES)
def test_register_chained_kernel(N, d, noise, loops, nb, monkeypatch):
    """k_mmar (T chained through registers) in two register budgets vs the C oracle and vs k_mma."""
    rng = np.random.default_rng(77 + N)
    mod = oracle_model(N, d=d, loops=loops)
    T, P = 70, 37
    x, _ = synth_traj(mod, T, rng, noise, p_nan=0.15)
    x[0] = np.nan if N % 2 else x[0]                     # odd N: first frame missing
    ss, thetas = random_profiles(rng, P, T, len(loops), 6)
    err = np.broadcast_to(np.asarray(noise, dtype=float), (d,))
    s2, Cind = ko.noise_to_s2_cind(err)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
    eng = engine_for(mod)
    traj = eng.trajectory(x, err)
    monkeypatch.setenv("BILDK_MMARB", "0")               # N mod 8 in {1, 2}: the all-tensor-core kernel, not k_mmarb
    if nb:   # most resident CTAs per SM = fewest registers (fragments re-read per tile row, some spills)
        monkeypatch.setenv("BILDK_MMAR_NB", "7" if N <= 24 else "5")
    assert traj.describe_plan(P).split()[0] == "mmar" and "border" not in traj.describe_plan(P)
    assert f"CTAs/SM={(7 if N <= 24 else 5) if nb else (4 if N <= 24 else 3)}" in traj.describe_plan(P)
    got = eng.logl_st(traj, ss, thetas)
    assert rel_err(got, want) < TOL
    monkeypatch.setenv("BILDK_KERNEL", "mma1")           # the shared-memory round-trip kernel on the same inputs
    assert traj.describe_plan(P).split()[0] != "mmar"
    assert rel_err(eng.logl_st(traj, ss, thetas), got) < 1e-12


MMAR2_CASES = [
    # N, d, noise, loops                      register-chained kernel with two warps per filter: GT 5..7, N mod 8 in 1..4
    (33, 3, 0.3, (None, [(0, -1)])),          # GT=5, r=1
    (36, 2, [0.1, 0.4], (None, [(0, -1)])),   # GT=5, r=4, d*=2
    (41, 1, 0.3, (None, [(0, -1)])),          # GT=6, r=1, d=1 (rows dealt out 0,3,5 | 1,2,4)
    (44, 4, 0.4, (None, [(0, -1)])),          # GT=6, r=4, d=4: every spare row carries a mean column
    (49, 3, 0.3, (None, [(0, -1)])),          # GT=7, r=1
    (50, 3, 0.3, (None, [(0, -1)], [(5, 30), (12, 44, 0.5)])),   # GT=7, r=2 (north-star N=50), 3 states
    (50, 3, [0.2, 0.2, 0.5], (None, [(0, -1)])),                 # ... anisotropic error: sub-filters with 2 and 1 columns
    (51, 3, 0.3, (None, [(0, -1)])),          # GT=7, r=3
    (52, 3, 0.3, (None, [(0, -1)])),          # GT=7, r=4
]


@pytest.mark.parametrize("fpc", [0, 1, 3], ids=["default", "one-pair-per-cta", "three-pairs"])
@pytest.mark.parametrize("N,d,noise,loops", MMAR2_CASES)
def test_register_chained_two_warp_kernel(N, d, noise, loops, fpc, monkeypatch):
    """k_mmar2 (tile rows of a filter split over a warp pair, T chained through registers) vs the C oracle and vs k_mma2
    (tile columns split, T through shared memory) on the same inputs."""
    rng = np.random.default_rng(177 + N)
    mod = oracle_model(N, d=d, loops=loops)
    T, P = 45, 23
    x, _ = synth_traj(mod, T, rng, noise, p_nan=0.15)
    x[0] = np.nan if N % 2 else x[0]                     # odd N: first frame missing
    ss, thetas = random_profiles(rng, P, T, len(loops), 6)
    err = np.broadcast_to(np.asarray(noise, dtype=float), (d,))
    s2, Cind = ko.noise_to_s2_cind(err)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
    eng = engine_for(mod)
    traj = eng.trajectory(x, err)
    if fpc:
        monkeypatch.setenv("BILDK_FPC2", str(fpc))
    assert traj.describe_plan(P).split()[0] == "mmar2"
    got = eng.logl_st(traj, ss, thetas)
    assert rel_err(got, want) < TOL
    monkeypatch.setenv("BILDK_MMAR2", "0")               # the column-split kernel on the same inputs
    assert traj.describe_plan(P).split()[0] == "mma2"
    assert rel_err(eng.logl_st(traj, ss, thetas), got) < 1e-12


BORDER_CASES = [
    # N, d, noise, loops                      k_mmarb: N mod 8 in {1, 2}, GT 2..4 - border rows / columns of C in DFMAs
    (9, 3, 0.4, (None, [(0, -1)])),           # GT=2, r=1
    (10, 3, 0.2, (None, [(0, -1)])),          # GT=2, r=2 (sweep size N=10)
    (10, 4, 0.3, (None, [(0, -1)])),          # ... d=4: four mean lanes
    (17, 1, 0.3, (None, [(0, -1)])),          # GT=3, r=1, d=1
    (18, 2, [0.1, 0.4], (None, [(0, -1)])),   # GT=3, r=2, d*=2
    (18, 3, 0.3, (None, [(0, -1)], [(2, 9), (5, 15, 0.5)])),   # GT=3, r=2, 3 states
    (25, 3, 0.3, (None, [(0, -1)])),          # GT=4, r=1 (sweep size N=25)
    (25, 3, [0.2, 0.2, 0.5], (None, [(0, -1)])),               # ... anisotropic error: sub-filters with 2 and 1 columns
    (26, 4, 0.3, (None, [(0, -1)])),          # GT=4, r=2, d=4: lanes 26..29 carry the border means
]


@pytest.mark.parametrize("N,d,noise,loops", BORDER_CASES)
def test_border_kernel(N, d, noise, loops, monkeypatch):
    """k_mmarb (tensor cores on the 8 (GT - 1) core rows, the N mod 8 border rows / columns as two DFMA matrix-vector
    products per frame) vs the C oracle and vs k_mmar (everything on padded 8x8 tiles) on the same inputs."""
    rng = np.random.default_rng(377 + N)
    mod = oracle_model(N, d=d, loops=loops)
    T, P = 70, 37
    x, _ = synth_traj(mod, T, rng, noise, p_nan=0.15)
    x[0] = np.nan if N % 2 else x[0]                     # odd N: first frame missing
    ss, thetas = random_profiles(rng, P, T, len(loops), 6)
    err = np.broadcast_to(np.asarray(noise, dtype=float), (d,))
    s2, Cind = ko.noise_to_s2_cind(err)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
    eng = engine_for(mod)
    traj = eng.trajectory(x, err)
    if N not in (17, 25, 26):
        monkeypatch.setenv("BILDK_MMARB", "2")           # compiled, but selected by default for N = 17, 25, 26 only (elsewhere k_mmar is faster)
    assert traj.describe_plan(P).split()[0] == "mmar" and "border-in-DFMA" in traj.describe_plan(P)
    got = eng.logl_st(traj, ss, thetas)
    assert rel_err(got, want) < TOL
    assert np.array_equal(got, eng.logl_states(traj, states))
    monkeypatch.setenv("BILDK_MMARB", "0")
    assert "border" not in traj.describe_plan(P)
    assert rel_err(eng.logl_st(traj, ss, thetas), got) < 1e-12


MX_CASES = [
    # N, d, noise, loops, kernel             register-chained kernels with the mean in an extra row block: N mod 8 in {0, 5, 6, 7}
    (5, 3, 0.5, (None, [(0, -1)]), "mmar"),            # GT=1, r=5
    (8, 4, 0.4, (None, [(0, -1)]), "mmar"),            # GT=1, r=8, d=4: all four mean rows in use
    (14, 2, [0.1, 0.4], (None, [(0, -1)]), "mmar"),    # GT=2, r=6, d*=2
    (16, 3, 0.2, (None, [(0, -1)]), "mmar"),           # GT=2, r=8
    (23, 1, 0.3, (None, [(0, -1)]), "mmar"),           # GT=3, r=7, d=1
    (24, 3, 0.3, (None, [(0, -1)], [(2, 9), (5, 15, 0.5)]), "mmar"),   # GT=3, r=8, 3 states
    (24, 3, [0.2, 0.2, 0.5], (None, [(0, -1)]), "mmar"),               # ... anisotropic error: sub-filters with 2 and 1 columns
    (29, 3, 0.3, (None, [(0, -1)]), "mmar"),           # GT=4, r=5
    (32, 3, 0.3, (None, [(0, -1)]), "mmar"),           # GT=4, r=8
    (37, 3, 0.3, (None, [(0, -1)]), "mmar2"),          # GT=5, r=5
    (40, 2, [0.1, 0.4], (None, [(0, -1)]), "mmar2"),   # GT=5, r=8, d*=2
    (46, 1, 0.3, (None, [(0, -1)]), "mmar2"),          # GT=6, r=6, d=1
    (48, 4, 0.4, (None, [(0, -1)]), "mmar2"),          # GT=6, r=8, d=4
    (55, 3, 0.3, (None, [(0, -1)], [(5, 30), (12, 44, 0.5)]), "mmar2"),   # GT=7, r=7, 3 states
    (56, 3, 0.3, (None, [(0, -1)]), "mmar2"),          # GT=7, r=8
]


@pytest.mark.parametrize("N,d,noise,loops,kernel", MX_CASES)
def test_register_chained_kernels_mean_in_extra_rows(N, d, noise, loops, kernel, monkeypatch):
    """k_mmar / k_mmar2 with M^T in an extra row block of the filter buffer (every k-tile full, the mean tile column read
    with permuted columns) vs the C oracle and vs the shared-memory round-trip kernels (k_mma / k_mma2) on the same inputs."""
    rng = np.random.default_rng(277 + N)
    mod = oracle_model(N, d=d, loops=loops)
    T, P = 50, 29
    x, _ = synth_traj(mod, T, rng, noise, p_nan=0.15)
    x[0] = np.nan if N % 2 else x[0]                     # odd N: first frame missing
    ss, thetas = random_profiles(rng, P, T, len(loops), 6)
    err = np.broadcast_to(np.asarray(noise, dtype=float), (d,))
    s2, Cind = ko.noise_to_s2_cind(err)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
    eng = engine_for(mod)
    traj = eng.trajectory(x, err)
    plan = traj.describe_plan(P)
    assert plan.split()[0] == kernel and "mean-in-extra-rows" in plan
    got = eng.logl_st(traj, ss, thetas)
    assert rel_err(got, want) < TOL
    assert np.array_equal(got, eng.logl_states(traj, states))
    monkeypatch.setenv("BILDK_MMAR_MX", "0")             # the older kernels on the same inputs
    assert traj.describe_plan(P).split()[0] in ("mma", "mma2")
    assert rel_err(eng.logl_st(traj, ss, thetas), got) < 1e-12


FOUR_WARP_CASES = [
    # N, d, noise, loops                      k_mmar2 with FOUR warps per filter (GT = 8: complementary row pairs {0,7} {1,6} {2,5} {3,4})
    (57, 3, 0.3, (None, [(0, -1)])),          # r=1
    (58, 2, [0.1, 0.4], (None, [(0, -1)])),   # r=2, d*=2
    (60, 4, 0.4, (None, [(0, -1)])),          # r=4, d=4: every spare row carries a mean column
    (60, 3, 0.3, (None, [(0, -1)], [(5, 30), (12, 44, 0.5)])),   # r=4, 3 states: three propagators resident -> fewer filters per CTA
    (61, 3, 0.3, (None, [(0, -1)])),          # r=5: mean in extra rows
    (63, 1, 0.3, (None, [(0, -1)])),          # r=7, d=1
    (64, 3, 0.3, (None, [(0, -1)])),          # r=8
    (64, 3, [0.2, 0.2, 0.5], (None, [(0, -1)])),                 # ... anisotropic error
]


@pytest.mark.parametrize("fpc", [0, 1, 2], ids=["default", "one-filter-per-cta", "two-filters"])
@pytest.mark.parametrize("N,d,noise,loops", FOUR_WARP_CASES)
def test_register_chained_four_warp_kernel(N, d, noise, loops, fpc, monkeypatch):
    """k_mmar2 at GT = 8 (N = 57..64): the tile rows of a filter split over four warps, vs the C oracle and vs k_mmac (one warp
    per tile column, T through shared memory) on the same inputs."""
    rng = np.random.default_rng(477 + N)
    mod = oracle_model(N, d=d, loops=loops)
    T, P = 40, 19
    x, _ = synth_traj(mod, T, rng, noise, p_nan=0.15)
    x[0] = np.nan if N % 2 else x[0]                     # odd N: first frame missing
    ss, thetas = random_profiles(rng, P, T, len(loops), 6)
    err = np.broadcast_to(np.asarray(noise, dtype=float), (d,))
    s2, Cind = ko.noise_to_s2_cind(err)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
    eng = engine_for(mod)
    traj = eng.trajectory(x, err)
    if fpc:
        monkeypatch.setenv("BILDK_FPC2", str(fpc))
    plan = traj.describe_plan(P)
    assert plan.split()[0] == "mmar2" and "four-warps-per-filter" in plan and ("mean-in-extra-rows" in plan) == (N > 60)
    got = eng.logl_st(traj, ss, thetas)
    assert rel_err(got, want) < TOL
    assert np.array_equal(got, eng.logl_states(traj, states))
    monkeypatch.setenv("BILDK_MMAR2", "0")               # one CTA per filter, one warp per tile column
    assert traj.describe_plan(P).split()[0] == "mmac"
    assert rel_err(eng.logl_st(traj, ss, thetas), got) < 1e-12


FIVE_WARP_CASES = [
    # N, d, noise, loops                      k_mmar2 with FIVE warps per filter (GT = 9: row pairs {0,8} {1,7} {2,6} {3,5} and the middle row {4})
    (65, 3, 0.3, (None, [(0, -1)])),          # r=1
    (68, 2, [0.1, 0.4], (None, [(0, -1)])),   # r=4, d*=2
    (68, 3, 0.3, (None, [(0, -1)], [(5, 30), (12, 44, 0.5)])),   # r=4, 3 states
    (69, 3, 0.3, (None, [(0, -1)])),          # r=5: mean in extra rows
    (72, 4, 0.4, (None, [(0, -1)])),          # r=8, d=4
]


@pytest.mark.parametrize("fpc", [0, 1], ids=["default", "one-filter-per-cta"])
@pytest.mark.parametrize("N,d,noise,loops", FIVE_WARP_CASES)
def test_register_chained_five_warp_kernel(N, d, noise, loops, fpc, monkeypatch):
    """k_mmar2 at GT = 9 (N = 65..72): four complementary row pairs and the middle row on five warps, vs the C oracle and vs
    k_mmact / k_mmac on the same inputs."""
    rng = np.random.default_rng(677 + N)
    mod = oracle_model(N, d=d, loops=loops)
    T, P = 36, 17
    x, _ = synth_traj(mod, T, rng, noise, p_nan=0.15)
    x[0] = np.nan if N % 2 else x[0]                     # odd N: first frame missing
    ss, thetas = random_profiles(rng, P, T, len(loops), 6)
    err = np.broadcast_to(np.asarray(noise, dtype=float), (d,))
    s2, Cind = ko.noise_to_s2_cind(err)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
    eng = engine_for(mod)
    traj = eng.trajectory(x, err)
    if fpc:
        monkeypatch.setenv("BILDK_FPC2", str(fpc))
    plan = traj.describe_plan(P)
    assert plan.split()[0] == "mmar2" and "five-warps-per-filter" in plan and ("mean-in-extra-rows" in plan) == (N > 68)
    got = eng.logl_st(traj, ss, thetas)
    assert rel_err(got, want) < TOL
    assert np.array_equal(got, eng.logl_states(traj, states))
    monkeypatch.setenv("BILDK_MMAR2", "0")
    assert traj.describe_plan(P).split()[0] in ("mmact", "mmac")
    assert rel_err(eng.logl_st(traj, ss, thetas), got) < 1e-12


EIGHT_WARP_CASES = [
    # N, d, noise, loops                      k_mmar8 (GT = 10..13): one filter per CTA, eight warps, one resident propagator swapped by TMA
    (73, 3, 0.3, (None, [(0, -1)])),          # GT=10, r=1
    (80, 2, [0.1, 0.4], (None, [(0, -1)], [(10, 50)])),    # GT=10, r=8, 3 states, d*=2
    (84, 3, 0.3, (None, [(0, -1)])),          # GT=11, r=4
    (87, 1, 0.3, (None, [(0, -1)])),          # GT=11, r=7, d=1
    (89, 3, 0.3, (None, [(0, -1)])),          # GT=12, r=1
    (96, 4, 0.4, (None, [(0, -1)])),          # GT=12, r=8, d=4
    (97, 3, 0.3, (None, [(0, -1)])),          # GT=13, r=1
    (100, 3, 0.3, (None, [(0, -1)])),         # r=4 (BASELINE configs[2])
    (100, 2, [0.2, 0.4], (None, [(0, -1)], [(10, 50)])),   # 3 states (propagator swaps between three), d*=2
    (101, 3, 0.3, (None, [(0, -1)])),         # r=5: mean in extra rows
    (104, 4, 0.4, (None, [(0, -1)])),         # r=8, d=4
]


@pytest.mark.parametrize("N,d,noise,loops", EIGHT_WARP_CASES)
def test_register_chained_eight_warp_kernel(N, d, noise, loops, monkeypatch):
    """k_mmar8 vs the C oracle and vs k_mmact / k_mmac (T through shared memory) on the same inputs; profiles with many state
    switches exercise the TMA swap of the single resident propagator."""
    rng = np.random.default_rng(577 + N)
    mod = oracle_model(N, d=d, loops=loops)
    T, P = 30, 11
    x, _ = synth_traj(mod, T, rng, noise, p_nan=0.15)
    x[0] = np.nan if N % 2 else x[0]                     # odd N: first frame missing
    ss, thetas = random_profiles(rng, P, T, len(loops), 8)
    err = np.broadcast_to(np.asarray(noise, dtype=float), (d,))
    s2, Cind = ko.noise_to_s2_cind(err)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
    eng = engine_for(mod)
    traj = eng.trajectory(x, err)
    plan = traj.describe_plan(P)
    assert plan.split()[0] == "mmar8" and ("mean-in-extra-rows" in plan) == ((N - 1) % 8 + 1 > 4)
    got = eng.logl_st(traj, ss, thetas)
    assert rel_err(got, want) < TOL
    assert np.array_equal(got, eng.logl_states(traj, states))
    monkeypatch.setenv("BILDK_MMAR8", "0")
    assert traj.describe_plan(P).split()[0] in ("mmact", "mmac")
    assert rel_err(eng.logl_st(traj, ss, thetas), got) < 1e-12


@pytest.mark.parametrize("N,nz", [
    (20, {3: -1.0, 14: 1.0}),            # two interior monomers (k_mma)
    (20, {7: 0.5, 8: -2.0}),             # neighbours in one tile, unequal weights
    (26, {0: -1.0, 25: 1.0}),            # GT=4, mean columns in padding
    (50, {10: -1.0, 40: 1.0}),           # k_mma GT=7
    (60, {5: 1.0, 58: -1.0}),            # k_mmac
    (100, {49: -1.0, 50: 1.0}),          # k_mmac, both non-zeros in one tile column
    (130, {1: -1.0, 120: 1.0}),          # k_mmag
    (20, {4: 1.0}),                      # one non-zero: absolute position of one monomer (tile kernel, sparse path)
    (20, {2: -1.0, 9: 0.5, 17: 0.5}),    # three non-zeros
    (24, {0: -1.0, 5: 1.0, 11: -1.0, 23: 1.0}),   # four non-zeros
])
def test_sparse_measurement_vectors(N, nz):
    """The measurement vector is arbitrary in the API (models.py:193-196); kernels specialise on its sparsity."""
    rng = np.random.default_rng(N + len(nz))
    w = np.zeros(N)
    for i, v in nz.items():
        w[i] = v
    mod = oracle_model(N, d=3, w=w)
    T, P = 30, 6
    x, _ = synth_traj(mod, T, rng, 0.25, p_nan=0.1)
    ss, thetas = random_profiles(rng, P, T, 2, 4)
    s2, Cind = ko.noise_to_s2_cind([0.25] * 3)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
    eng = engine_for(mod)
    got = eng.logl_st(eng.trajectory(x, [0.25] * 3), ss, thetas)
    assert rel_err(got, want) < TOL


def test_dense_measurement_vector():
    rng = np.random.default_rng(5)
    N, d, T, P = 20, 3, 60, 21
    w = rng.normal(size=N)
    w -= w.mean()
    mod = oracle_model(N, d=d, w=w)
    x, _ = synth_traj(mod, T, rng, 0.2, p_nan=0.1)
    ss, thetas = random_profiles(rng, P, T, 2, 5)
    s2, Cind = ko.noise_to_s2_cind([0.2] * d)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
    eng = engine_for(mod)
    got = eng.logl_st(eng.trajectory(x, [0.2] * d), ss, thetas)
    assert rel_err(got, want) < TOL


@pytest.mark.parametrize("N", [15, 18, 19, 25, 40, 100])   # 15, 40: k_mmar / k_mmar2 with the mean in extra rows; 18, 25: k_mmarb (border means); 19: k_mmar; 100: k_mmar8
def test_external_force_mean_offset(N):
    """G != 0 (pyx:209): constant force on the chain ends."""
    rng = np.random.default_rng(6)
    d, T, P = 2, 40, 10
    F = np.zeros((N, d))
    F[0, 0], F[-1, 0] = -0.7, 0.7
    mod = oracle_model(N, d=d, F=F)
    assert np.any(mod["Gs"] != 0)
    x, _ = synth_traj(mod, T, rng, 0.2)
    ss, thetas = random_profiles(rng, P, T, 2, 4)
    s2, Cind = ko.noise_to_s2_cind([0.2] * d)
    states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
    eng = engine_for(mod)
    got = eng.logl_st(eng.trajectory(x, [0.2] * d), ss, thetas)
    assert rel_err(got, want) < TOL


def test_every_polymer_size():
    """Every N from 2 to 116 once (kernel selection has a dozen classes by tile grid and N mod 8): a short trajectory with a
    missing frame, a handful of profiles, d = 3, against the C oracle; the plan must be one of the register-chained kernels
    wherever DESIGN.md section 4 says so."""
    expect = {}
    for N in range(2, 117):
        GT, r = (N + 7) // 8, N - 8 * ((N + 7) // 8 - 1)
        expect[N] = "mmar" if GT <= 4 else "mmar2" if GT <= 9 else "mmar8" if GT <= 13 else None
    rng = np.random.default_rng(4242)
    s2, Cind = ko.noise_to_s2_cind([0.3] * 3)
    worst = 0.0
    for N in range(2, 117):
        mod = oracle_model(N, d=3)
        T, P = 7, 4
        x, _ = synth_traj(mod, T, rng, 0.3)
        x[3] = np.nan
        ss, thetas = random_profiles(rng, P, T, 2, 3)
        states = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
        want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, states)
        eng = engine_for(mod)
        traj = eng.trajectory(x, [0.3] * 3)
        plan = traj.describe_plan(P).split()[0]
        if expect[N] is not None:
            assert plan == expect[N], (N, plan)
        got = eng.logl_st(traj, ss, thetas)
        err = rel_err(got, want)
        assert err < TOL, (N, plan, err)
        worst = max(worst, err)
    assert worst < 1e-11


def test_no_valid_frames_and_single_frame():
    mod = oracle_model(12, d=2)
    eng = engine_for(mod)
    x = np.full((9, 2), np.nan)
    traj = eng.trajectory(x, [0.3, 0.3])
    assert np.array_equal(eng.logl_states(traj, np.zeros((3, 9), dtype=int)), np.zeros(3))   # _py.py:121: sum of nothing
    x1 = np.array([[0.4, -0.1]])
    traj1 = eng.trajectory(x1, [0.3, 0.3])
    s2, Cind = ko.noise_to_s2_cind([0.3, 0.3])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x1, s2, Cind, np.array([[1]]))
    assert rel_err(eng.logl_states(traj1, np.array([[1]])), want) < TOL


@pytest.mark.parametrize("N", [20, 25, 16, 40, 50, 100])   # k_mmar, k_mmarb, k_mmar MX, k_mmar2 MX, k_mmar2, k_mmar8
def test_multi_trajectory_batch(N):
    """Fused multi-trajectory launch (bildk_logl_runs_multi: trajectories of different lengths and batch sizes in one grid,
    CTA -> trajectory map) against the C oracle, for every kernel family the dataset driver can meet."""
    rng = np.random.default_rng(11 + N)
    d = 3
    mod = oracle_model(N, d=d)
    eng = engine_for(mod)
    s2, Cind = ko.noise_to_s2_cind([0.3] * d)
    trajs, wants, starts_all, states_all, offsets = [], [], [], [], [0]
    from bild_b200.engine import st_to_runs
    for i, (T, P) in enumerate([(40, 7), (75, 1), (40, 12), (90, 5)]):
        x, _ = synth_traj(mod, T, rng, 0.3, p_nan=0.1 * i)
        ss, thetas = random_profiles(rng, P, T, 2, 6)
        st = np.array([ko.st2states(s, th, T) for s, th in zip(ss, thetas)])
        wants.append(ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, st))
        trajs.append(eng.trajectory(x, [0.3] * d))
        a, b = st_to_runs(ss, thetas, T)
        starts_all.append(a); states_all.append(b); offsets.append(offsets[-1] + P)
    got = eng.logl_runs_multi(trajs, offsets, np.concatenate(starts_all), np.concatenate(states_all))
    assert rel_err(got, np.concatenate(wants)) < TOL


def test_error_behaviour():
    mod = oracle_model(8, d=1)
    eng = engine_for(mod)
    traj = eng.trajectory(np.arange(5.0), [0.5])
    with pytest.raises(ValueError):
        eng.logl_states(traj, np.array([[0, 1, 2, 0, 0]]))          # state out of range
    with pytest.raises(ValueError):
        eng.logl_states(traj, np.zeros((1, 4), dtype=int))           # length mismatch
    with pytest.raises(ValueError):
        eng.trajectory(np.arange(5.0), [np.nan])
    assert eng.logl_states(traj, np.zeros((0, 5), dtype=int)).shape == (0,)


def test_round_trip_properties_full_size():
    """BASELINE config 2 size (N=20, T=500, P=4096): properties that need no oracle."""
    rng = np.random.default_rng(2)
    N, d, T, P = 20, 3, 500, 4096
    mod = oracle_model(N, d=d)
    x, _ = synth_traj(mod, T, rng, 0.3, p_nan=0.0)
    ss, thetas = random_profiles(rng, P, T, 2, 10)
    eng = engine_for(mod)
    traj = eng.trajectory(x, [0.3] * d)
    a = eng.logl_st(traj, ss, thetas)
    assert np.all(np.isfinite(a))
    # (1) permutation equivariance + run-to-run determinism (bitwise)
    perm = rng.permutation(P)
    b = eng.logl_st(traj, ss[perm], thetas[perm])
    assert np.array_equal(a[perm], b)
    # (2) batch-split invariance (bitwise)
    c = np.concatenate([eng.logl_st(traj, ss[:1000], thetas[:1000]), eng.logl_st(traj, ss[1000:], thetas[1000:])])
    assert np.array_equal(a, c)
    # (3) a sample of the batch against the oracle
    idx = rng.choice(P, 24, replace=False)
    s2, Cind = ko.noise_to_s2_cind([0.3] * d)
    st = np.array([ko.st2states(ss[i], thetas[i], T) for i in idx])
    want = ko.logl_c(*(mod[k] for k in MODEL_KEYS), x, s2, Cind, st)
    assert rel_err(a[idx], want) < TOL
    # (4) spatial dimensions with equal error are exchangeable: permuting columns leaves logL unchanged up to summation order
    traj_p = eng.trajectory(x[:, [2, 0, 1]], [0.3] * d)
    e = eng.logl_st(traj_p, ss[:256], thetas[:256])
    assert rel_err(e, a[:256]) < 1e-12


FULL_SIZE = [
    # name, N, T, P, p_nan, n_check                                      BASELINE.json sizes, checked on a random sample
    ("north-star N=50 T=1000", 50, 1000, 16384, 0.0, 256),                # k_mma2
    ("configs[2] N=100 T=1000 10% NaN", 100, 1000, 16384, 0.10, 256),     # k_mmar8
    ("sweep N=200 T=100", 200, 100, 1024, 0.0, 256),                      # k_mmag2
]


@pytest.mark.parametrize("name,N,T,P,p_nan,n_check", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_full_size_sampled_parity(name, N, T, P, p_nan, n_check):
    """The batch of the BASELINE configuration at FULL size on the GPU; `n_check` randomly chosen filters of it against the
    C oracle (all host cores), 1e-9 relative; plus the size-independent properties (bitwise permutation equivariance)."""
    rng = np.random.default_rng(50 + N)
    mod = oracle_model(N, d=3)
    x, _ = synth_traj(mod, T, rng, 0.3, p_nan=p_nan)
    ss, thetas = random_profiles(rng, P, T, 2, 10)
    eng = engine_for(mod)
    traj = eng.trajectory(x, [0.3] * 3)
    got = eng.logl_st(traj, ss, thetas)
    assert np.all(np.isfinite(got))
    idx = np.sort(rng.choice(P, n_check, replace=False))
    s2, Cind = ko.noise_to_s2_cind([0.3] * 3)
    st = np.array([ko.st2states(ss[i], thetas[i], T) for i in idx])
    want = logl_c_parallel(mod, x, s2, Cind, st)
    assert rel_err(got[idx], want) < TOL
    # the checked filters evaluated alone, in another order, in a small batch (other CTA / wave geometry): same bits
    perm = rng.permutation(n_check)
    again = eng.logl_st(traj, ss[idx][perm], thetas[idx][perm])
    assert np.array_equal(again, got[idx][perm])


def test_in_library_profile_coding_matches_numpy():
    """bildk_logl_st codes (s, theta) with numpy's arithmetic: same logL bits as coding with st_to_runs first."""
    from bild_b200.engine import st_to_runs
    rng = np.random.default_rng(9)
    N, d, T, P = 20, 3, 101, 512
    mod = oracle_model(N, d=d)
    x, _ = synth_traj(mod, T, rng, 0.3)
    ss, thetas = random_profiles(rng, P, T, 2, 10)
    # adversarial rows: exact ties on frame boundaries, zero-length intervals, cumsum reaching 1.0 early
    ss[0] = 0; ss[0, :4] = [0.25, 0.5, 0.25, 0.0]
    ss[1] = 0; ss[1, :3] = [0.5, 0.0, 0.5]
    ss[2] = 0; ss[2, :3] = [0.999, 0.0005, 0.0005]
    ss[3] = 0; ss[3, :2] = [1.0, 0.0]
    ss[4] = 0; ss[4, :11] = np.full(11, 1 / 11)
    ss[5] = 0; ss[5, :3] = [1e-17, 1.0 - 1e-16, 1e-16]
    eng = engine_for(mod)
    traj = eng.trajectory(x, [0.3] * d)
    a = eng.logl_st(traj, ss, thetas)
    starts, rst = st_to_runs(ss, thetas, T)
    b = eng.logl_runs(traj, starts, rst)
    assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        eng.logl_st(traj, ss, thetas + 2)                       # state out of range
    bad = ss.copy(); bad[7, 0] = np.nan
    with pytest.raises(ValueError):
        eng.logl_st(traj, bad, thetas)
