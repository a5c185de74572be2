"""
The drop-in boundary, EXECUTED: the unmodified reference package (staged by oracle/stage_reference.py into
baseline/_ref_b200/, git-ignored, travels to the GPU box) with `integration/MSRouse_logL.py` - the ctypes binding of
libbild_b200.so - in the reference's own native plugin slot (/root/reference/bild/cython_imports.py:3-7).  Nothing of
bild_b200's Python layer is involved: reference models.py / amis.py / core.py -> stub -> C ABI -> CUDA kernels.
"""
import os

import numpy as np
import pytest

import stage_reference as sr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.isdir(os.path.join(ROOT, "baseline", "_ref_b200", "bild"))
needs_staged = pytest.mark.skipif(not STAGED, reason="baseline/_ref_b200 not staged (run __graft_entry__.build() in the build container)")


@pytest.fixture(scope="module")
def ref_b200():
    bild = sr.import_reference("_ref_b200")
    yield bild
    import sys
    for name in [m for m in sys.modules if m == "bild" or m.startswith("bild.")]:
        del sys.modules[name]


@needs_staged
def test_stub_sits_in_the_plugin_slot(ref_b200):
    """CPU: the reference resolves its likelihood to the stub (no warning-and-fallback to the Python twin), and the staged
    tree is byte-identical to the reference sources apart from bin/."""
    fn = ref_b200.models.MSRouse_logL
    assert fn.__module__ == "bild.bin.MSRouse_logL" and fn.__code__.co_filename.endswith(os.path.join("bin", "MSRouse_logL.py"))
    assert open(os.path.join(ROOT, "baseline", "_ref_b200", "bild", "bin", "MSRouse_logL.py")).read() == \
        open(os.path.join(ROOT, "integration", "MSRouse_logL.py")).read()
    if os.path.isdir("/root/reference/bild"):
        for name in ("models.py", "amis.py", "core.py", "cython_imports.py", "choicesampler.py", "postproc.py", "util.py"):
            assert open(os.path.join("/root/reference/bild", name), "rb").read() == \
                open(os.path.join(ROOT, "baseline", "_ref_b200", "bild", name), "rb").read(), name


@needs_staged
def test_stub_has_no_cpu_fallback(ref_b200):
    from bild_b200 import _lib
    if _lib.load().bildk_device_count() > 0:
        pytest.skip("a device is present")
    import noctiluca as nl
    traj = nl.Trajectory(np.array([1, 2, np.nan, 4]), localization_error=[0.5])
    model = ref_b200.models.MultiStateRouse(20, 1, 5, d=1)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        model.logL(ref_b200.Loopingprofile([1, 1, 0, 0]), traj)


@needs_staged
@pytest.mark.gpu
def test_reference_fixture_through_the_stub(ref_b200):
    """/root/reference/tests/test_bild.py:125-148 with the engine behind the unmodified reference model class."""
    import noctiluca as nl
    traj = nl.Trajectory(np.array([1, 2, np.nan, 4]), localization_error=[0.5])
    profile = ref_b200.Loopingprofile([1, 1, 0, 0])
    model = ref_b200.models.MultiStateRouse(20, 1, 5, d=1)
    v = model.logL(profile, traj)
    assert -100 < v < 0
    g = np.load(os.path.join(ROOT, "tests", "golden", "logl_fixture_test_bild.npz"))
    assert abs(v - g["logL_cy"][0]) < 1e-9 * abs(g["logL_cy"][0])
    model2 = ref_b200.models.MultiStateRouse(20, 1, 5, d=1, localization_error=0.5)
    assert model2.logL(profile, traj) == v
    traj.localization_error = None
    with pytest.raises(ValueError):
        model.logL(profile, traj)
    # editing the model between calls takes effect (the .pyx re-reads the dynamics on every call, pyx:150-160)
    traj.localization_error = np.array([0.5])
    model.models[1].k = 2.0
    model.models[1].update_dynamics()
    v2 = model.logL(profile, traj)
    assert v2 != v
    fresh = ref_b200.models.MultiStateRouse(20, 1, 5, d=1)
    fresh.models[1].k = 2.0
    fresh.models[1].update_dynamics()
    assert fresh.logL(profile, traj) == v2


@needs_staged
@pytest.mark.gpu
@pytest.mark.parametrize("batched", [False, True])
def test_reference_bild_sample_through_the_stub(ref_b200, golden_dir, batched):
    """BASELINE.json configs[0] run by the REFERENCE's own core.py / amis.py / choicesampler.py with the engine in the
    plugin slot: one launch per profile (unmodified call pattern, amis.py:735-739) or, with the batched hook of
    INTEGRATION.md, one launch per batch.  Reproduces the run recorded from the reference with its own .pyx."""
    import noctiluca as nl
    from test_amis_host import check_c1_evidence
    stub = ref_b200.models.MSRouse_logL.__globals__
    runs = np.load(os.path.join(golden_dir, "sample_runs.npz"))
    replay = np.load(os.path.join(golden_dir, "sample_c1_replay.npz"))
    original = stub["install_batched_hook"](ref_b200) if batched else None
    try:
        model = ref_b200.models.MultiStateRouse(20, 1, 5, d=3, localization_error=0.3)
        traj = nl.Trajectory(runs["c1_x"], localization_error=[0.3] * 3)
        from bild_b200 import _lib
        n0 = _lib.load().bildk_launch_count()      # same libbild_b200.so image as the stub's CDLL handle
        np.random.seed(1234)
        res = ref_b200.sample(traj, model)
        launched = _lib.load().bildk_launch_count() - n0
    finally:
        if original is not None:
            ref_b200.amis.FixedkSampler.logL = original
    assert np.array_equal(res.k, runs["c1_k"])
    assert np.array_equal(res.log["k"], runs["c1_logk"])
    check_c1_evidence(res, runs, replay)
    n_eval = int(sum(len(s["logLs"]) for smp in res.samplers for s in smp.samples))
    assert launched >= (int(runs["c1_n_logl_batches"]) if batched else n_eval)
