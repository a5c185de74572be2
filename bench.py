#!/usr/bin/env python
"""
bench.py - throughput of the BILD profile-likelihood hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|n50|...] [--impl reference]

A "step" is one AMIS-iteration-shaped pass of the hot path: the batched multi-state-Rouse Kalman
log-likelihood of P sampled profiles on one trajectory (+ the all-gather of logL across ranks when
N > 1) + the AMIS weight normalisation.  Default workload = BASELINE.json configs[1]:
4096 profiles x 1 trajectory, N=20 monomers, T=500 frames, d=3, on 1 B200.  Multi-GPU runs are weak
scaling: every rank evaluates its own batch of P profiles (global batch N*P), one NCCL all-gather of
the logL vector per step, identical deterministic weight reduction on every rank.

Prints ONE JSON line (rank 0).  `value` = frame-steps/s with inputs resident in HBM (CUDA events);
`e2e` = the same through the host-buffer API (numpy in, numpy out: host-side run-length coding,
H2D, kernel, D2H inside the timed region); `roofline` = FP64 roofline of the filter kernel;
`cpu_baseline` = the reference's own Cython implementation (oracle/_ref, compiled from the reference
.pyx) on all host cores, same profiles, with the parity of the GPU results against it.
`--impl reference` times only that CPU implementation, in the same JSON format.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

import numpy as np

# The contract is ONE JSON line on stdout.  Native libraries (NCCL prints its version banner to fd 1) must not
# add lines: keep a private copy of the real stdout for the result and point fd 1 at stderr for everybody else.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: N, T, P (per GPU), p_nan          -- BASELINE.json configs / SURVEY.md section 8(d)
    "c2": dict(N=20, T=500, P=4096, p_nan=0.0, desc="configs[1]: 4096 profiles x 1 trajectory, N=20, T=500"),
    "c3": dict(N=100, T=1000, P=16384, p_nan=0.10, desc="configs[2]: 16384 profiles, N=100, T=1000, 10% NaN"),
    "n50": dict(N=50, T=1000, P=16384, p_nan=0.0, desc="north-star target shape: N=50, T=1000"),
    "n10": dict(N=10, T=1000, P=65536, p_nan=0.0, desc="sweep point N=10, T=1000, 64k profiles"),
    "n25": dict(N=25, T=1000, P=65536, p_nan=0.0, desc="sweep point N=25, T=1000, 64k profiles"),
    "n200": dict(N=200, T=100, P=1024, p_nan=0.0, desc="sweep point N=200, T=100, 1024 profiles"),
    "n150": dict(N=150, T=100, P=2048, p_nan=0.0, desc="N=150, T=100, 2048 profiles"),
    "n20big": dict(N=20, T=500, P=65536, p_nan=0.0, desc="N=20, T=500, 64k profiles"),
}
D_SPATIAL, DIFF, KSPRING, LOC_ERR, KMAX = 3, 1.0, 5.0, 0.3, 10
SEED = 685441950   # /root/reference/tests/test_bild.py:9


# ------------------------------------------------------------------------------------------------ synthetic inputs
def make_inputs(wl, rank):
    """Model, trajectory and an AMIS-like profile batch; all FP64, synthetic."""
    from bild_b200.models import MultiStateRouse
    from bild_b200.util import Loopingprofile
    N, T, P = wl["N"], wl["T"], wl["P"]
    model = MultiStateRouse(N, DIFF, KSPRING, d=D_SPATIAL, localization_error=LOC_ERR)
    np.random.seed(SEED)
    dwell = max(2, T // 5)
    truth = (np.cumsum(np.random.rand(T) < 1.0 / dwell) % 2).astype(int)
    traj = model.trajectory_from_loopingprofile(Loopingprofile(truth), missing_frames=wl["p_nan"] or None)
    rng = np.random.default_rng(SEED + 1 + rank)
    K1 = KMAX + 1
    ks = rng.integers(0, KMAX + 1, size=P)
    ss = np.zeros((P, K1))
    for k in range(KMAX + 1):
        idx = np.nonzero(ks == k)[0]
        if len(idx):
            ss[idx, :k + 1] = rng.dirichlet(np.ones(k + 1), size=len(idx))
    thetas = (rng.integers(0, 2, size=(P, 1)) + np.arange(K1)[None, :]) % 2   # 2 states: CFC = alternate
    return model, traj, ss, thetas


def flops_per_eval(N, d, dstar, T, V):
    """Algorithmic FP64 flop per log-likelihood evaluation."""
    f_prop = dstar * (4 * N ** 3 + N ** 2) + 2 * N ** 2 * d
    f_upd = dstar * (4 * N ** 2 + 3 * N) + d * (4 * N + 8)
    return (T - 1) * f_prop + V * f_upd


# ------------------------------------------------------------------------------------------------ CPU reference arm
_W = {}


def _cpu_worker(chunk):
    fn, model, traj, states = _W["fn"], _W["model"], _W["traj"], _W["states"]
    from bild_b200.util import Loopingprofile
    return [fn(model, Loopingprofile(states[i]), traj) for i in chunk]


def cpu_reference(model, traj, states, n_sample, repeats=1):
    """
    Time the reference CPU implementation on `n_sample` profiles using every host core
    (multiprocessing fork pool, one BLAS thread per worker, warm-up pass excluded).
    kind "reference" = oracle/_ref (the reference's MSRouse_logL.pyx compiled as is); if that build is
    absent, kind "port" = the C restatement oracle/kalman_oracle.c.
    """
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kalman_oracle as ko
    fn = ko.ref_cython()
    kind = "reference"
    if fn is None:
        kind = "port"
        arrs = ko.model_arrays(model.models)
        s2, cind = ko.noise_to_s2_cind(model._get_noise(traj))

        def fn(model_, profile, traj_):
            return ko.logl_c(*arrs, model_.measurement, traj_[:], s2, cind, profile[:])
    cores = os.cpu_count() or 1
    n_sample = min(n_sample, len(states))
    _W.update(fn=fn, model=model, traj=traj, states=states)
    chunks = [list(range(i, n_sample, cores)) for i in range(cores)]
    chunks = [c for c in chunks if c]
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        pool.map(_cpu_worker, [c[:1] for c in chunks])          # warm-up: imports, caches
        best, vals = None, None
        for _ in range(repeats):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, chunks)
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, vals = dt, res
    out = np.empty(n_sample)
    for c, r in zip(chunks, vals):
        out[c] = r
    return dict(seconds=best, n=n_sample, cores=len(chunks), kind=kind, logL=out)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock, power and clock-event (throttle) reasons sampled DURING the timed region: NVML in-process
    (a few hundred samples per second); falls back to polling nvidia-smi when NVML is unavailable."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.nvml_error = None
        self._nv = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _init_nvml(self):
        """Synchronous (NVML start-up can take longer than the whole timed region)."""
        import pynvml as nv
        nv.nvmlInit()
        # CUDA_VISIBLE_DEVICES may renumber devices; resolve through the UUID of the CUDA device when possible
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            for i in range(nv.nvmlDeviceGetCount()):
                hi = nv.nvmlDeviceGetHandleByIndex(i)
                u = nv.nvmlDeviceGetUUID(hi)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    h = hi
                    break
        except Exception:
            pass
        self._nv, self._h = nv, h
        self._mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        self._sample_nvml()            # fails here, not in the thread, if a query is unsupported

    def _sample_nvml(self):
        nv, h = self._nv, self._h
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        self.rows.append([time.perf_counter(), str(sm), str(self._mx), str(pw)] +
                         ["Active" if r & bits[n] else "Not Active" for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])

    def _run_nvml(self):
        while not self.stop.is_set():
            self._sample_nvml()
            self.stop.wait(0.004)

    def _run(self):
        if self._nv is not None:
            try:
                self._run_nvml()
                return
            except Exception as err:  # noqa: BLE001
                self.nvml_error = repr(err)
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([time.perf_counter()] + [c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self.stop.wait(0.02)

    def __enter__(self):
        try:
            self._init_nvml()
        except Exception as err:  # noqa: BLE001
            self._nv, self.nvml_error = None, repr(err)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self, t0=None, t1=None):
        """Median SM clock / reasons over the samples taken inside [t0, t1] (all samples if none fell inside)."""
        rows = [r[1:] for r in self.rows if t0 is None or t0 <= r[0] <= t1] or [r[1:] for r in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "error": self.nvml_error}
        self_rows = rows
        sm = sorted(float(r[0]) for r in self_rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self_rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self_rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self_rows), "reasons": reasons, "samples": len(self_rows)}


# ------------------------------------------------------------------------------------------------ arms
def run_reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    model, traj, ss, thetas = make_inputs(wl, 0)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kalman_oracle as ko
    T = wl["T"]
    # bounded sample per step: about 1.5 s of wall on all cores, estimated from a short probe
    n_probe = min(32, wl["P"])
    st_probe = np.array([ko.st2states(ss[i], thetas[i], T) for i in range(n_probe)])
    probe = cpu_reference(model, traj, st_probe, n_probe)
    per_eval_core = probe["seconds"] * probe["cores"] / max(1, probe["n"]) if probe["n"] >= probe["cores"] else probe["seconds"]
    n_sample = int(min(wl["P"], max(probe["cores"], 1.5 * probe["cores"] / max(per_eval_core, 1e-9))))
    states = np.array([ko.st2states(ss[i], thetas[i], T) for i in range(n_sample)])
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference(model, traj, states, n_sample)
        if i >= args.warmup:
            times.append(r["seconds"])
    tot = float(np.sum(times))
    fs = n_sample * (T - 1) * args.steps / tot
    line = {
        "impl": "reference", "metric": "profile_logL_frame_steps_per_sec", "value": fs, "unit": "frame-steps/s",
        "evals_per_s": n_sample * args.steps / tot,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "N": wl["N"], "T": T, "d": D_SPATIAL, "profiles_per_step": n_sample,
                   "note": "CPU arm: rank 0 only, all host cores, bounded sample of the batch per step"},
        "cpu_baseline": {"value": fs, "unit": "frame-steps/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": f"{n_sample} of {wl['P']} profiles per step, multiprocessing fork pool, 1 BLAS thread per worker"},
        "e2e": {"value": fs, "unit": "frame-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_gpu_arm(args, wl, rank, world, local_rank):
    N, T, P = wl["N"], wl["T"], wl["P"]
    model, traj, ss, thetas = make_inputs(wl, rank)
    V = traj.count_valid_frames()

    # ---- CPU baseline first (fork pool must precede CUDA initialisation), rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import kalman_oracle as ko
        probe_states = np.array([ko.st2states(ss[i], thetas[i], T) for i in range(min(16, P))])
        probe = cpu_reference(model, traj, probe_states, len(probe_states))
        per_eval = probe["seconds"] * min(probe["cores"], probe["n"]) / probe["n"]     # core-seconds per eval
        n_sample = int(min(P, max(16, 20.0 / max(per_eval, 1e-9))))                      # ~20 s of CPU work
        states = np.array([ko.st2states(ss[i], thetas[i], T) for i in range(n_sample)])
        cpu = cpu_reference(model, traj, states, n_sample)

    import torch
    import torch.distributed as dist
    from bild_b200 import _lib
    from bild_b200.engine import st_to_runs
    import ctypes

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model.device = local_rank
    eng = model.engine
    th = model._handle(traj)
    lib = _lib.load()

    # ---- FP64 peak, measured here (no FP64 entry in MEASURED_PEAKS.json)
    dfma, dmma = ctypes.c_double(), ctypes.c_double()
    _lib.check(lib.bildk_measure_fp64_peak(local_rank, ctypes.byref(dfma), ctypes.byref(dmma)))
    peak_tf = max(dfma.value, dmma.value)

    starts, rstates = st_to_runs(ss, thetas, T)
    K1 = starts.shape[1]
    d_starts = torch.from_numpy(starts).to(dev)
    d_states = torch.from_numpy(rstates).to(dev)
    d_out = torch.empty(P, dtype=torch.float64, device=dev)
    d_all = torch.empty(P * world, dtype=torch.float64, device=dev)
    g = torch.Generator(device="cpu").manual_seed(5)
    d_logdelta = (torch.randn(P * world, generator=g, dtype=torch.float64) - 40.0).to(dev)   # synthetic proposal terms
    d_curlp = (torch.randn(P * world, generator=g, dtype=torch.float64) - 40.0).to(dev)
    d_logw = torch.empty(P * world, dtype=torch.float64, device=dev)
    d_stats = torch.empty(4, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device():
        eng.logl_runs_device(th, P, K1, d_starts.data_ptr(), d_states.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_out)
            src = d_all
        else:
            src = d_out
        _lib.check(lib.bildk_amis_weights_device(P * world, ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(d_logdelta.data_ptr()),
                                                 ctypes.c_void_p(d_curlp.data_ptr()), float(np.log(7.0)),
                                                 ctypes.c_void_p(d_logw.data_ptr()), ctypes.c_void_p(d_stats.data_ptr()),
                                                 ctypes.c_void_p(stream.cuda_stream)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(fn, n):
        evs = []
        for _ in range(n):
            flush.zero_()                                   # L2 flush between timed iterations (untimed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    with ClockSampler(local_rank) as clk:       # started before the warm-up so that nvidia-smi is warm; only samples
        for _ in range(max(args.warmup, 3)):    # inside the timed window are reported
            step_device()
        barrier()
        launches0 = lib.bildk_launch_count()
        wall0 = time.perf_counter()
        ms = timed_steps(step_device, args.steps)
        barrier()
        wall = time.perf_counter() - wall0
        launches = lib.bildk_launch_count() - launches0
        # kernel-only duration of the filter kernel for the roofline (same stream, CUDA events, L2 flushed)
        kms = timed_steps(lambda: eng.logl_runs_device(th, P, K1, d_starts.data_ptr(), d_states.data_ptr(), d_out.data_ptr(),
                                                       stream.cuda_stream), max(3, min(args.steps, 10)))
        clocks = clk.summary(wall0, time.perf_counter())
    t_total = torch.tensor([float(np.sum(ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    total_ms = float(t_total.item())
    gpu_logl = d_out.cpu().numpy().copy()

    # ---- end to end through the host-buffer API (numpy in -> numpy out), AMIS weights on the host side
    def step_e2e():
        ll = model.logL_st_batch(ss, thetas, traj)
        return ll

    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ll_host = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())
    assert np.array_equal(ll_host, gpu_logl), "host-buffer and device-resident paths disagree"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frame_steps = P * world * (T - 1)
    value = frame_steps * args.steps / (total_ms * 1e-3)
    kavg_ms = float(np.mean(kms))
    fl = flops_per_eval(N, D_SPATIAL, 1, T, V) * P
    ach_tf = fl / (kavg_ms * 1e-3) * 1e-12
    line = {
        "metric": "profile_logL_frame_steps_per_sec", "value": value, "unit": "frame-steps/s",
        "evals_per_s": P * world * args.steps / (total_ms * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "N": N, "T": T, "d": D_SPATIAL, "states": 2, "profiles_per_gpu": P,
                   "global_profiles": P * world, "valid_frames": V, "kmax": KMAX,
                   "step": "logL kernel" + (" + NCCL all-gather of logL" if world > 1 else "") + " + AMIS weight reduction",
                   "l2": "flushed between timed steps (256 MiB memset, untimed)", "plan": th.describe_plan(P),
                   "parallelism": f"profiles sharded, {world} rank(s)"},
        "e2e": {"value": frame_steps * args.steps / e2e_s, "unit": "frame-steps/s",
                "h2d_bytes_per_step": int(starts.nbytes + rstates.nbytes) * world, "d2h_bytes_per_step": int(P * 8) * world,
                "api": "MultiStateRouse.logL_st_batch(ss, thetas, traj): numpy in/out, pageable host buffers"},
        "gpu_launches": int(launches),
        "wall_s_timed_region": wall,
        "clocks": clocks,
        # "tensor": the bounding unit is the FP64 tensor pipe (DMMA m8n8k4) - the peak below is ITS measured rate, not bf16
        "roofline": {"bound": "tensor", "pipe": "fp64 DMMA m8n8k4 (no tcgen05 kind for f64)", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                     "traffic": _ncu_traffic(args.workload if not args.profiles else None), "kernel": th.describe_plan(P).split(" FPC")[0].split(" WPC")[0], "kernel_ms": kavg_ms,
                     "flop_per_launch": fl, "flop_model": "4N^3 d* per frame-step + lower-order terms (SURVEY.md 8d); the DMMA kernels execute "
                                                          "3N^3 (symmetric output) on 8x8 tiles - achieved counts ALGORITHMIC flops only",
                     "algorithmic_hbm_bytes_per_launch": int(starts.nbytes + rstates.nbytes + traj[:].nbytes + P * 8),
                     "peak_source": f"measured in this run: DFMA {dfma.value:.2f}, DMMA {dmma.value:.2f} TFLOP/s (bildk_measure_fp64_peak)",
                     "hbm_streaming_model": {"bytes_per_frame_step": 16 * N * N,
                                             "achieved_gbs": 16 * N * N * P * (T - 1) / (kavg_ms * 1e-3) * 1e-9,
                                             "peak_gbs": _hbm_peak(), "note": "state is on-chip; what streaming C from HBM would need"}},
    }
    if cpu is not None:
        cfs = cpu["n"] * (T - 1) / cpu["seconds"]
        rel = float(np.max(np.abs(gpu_logl[:cpu["n"]] - cpu["logL"]) / np.maximum(1.0, np.abs(cpu["logL"]))))
        line["cpu_baseline"] = {"value": cfs, "unit": "frame-steps/s", "cores": cpu["cores"], "kind": cpu["kind"],
                                "sample": f"first {cpu['n']} of {P} profiles, multiprocessing fork pool, 1 BLAS thread per worker",
                                "evals_per_s": cpu["n"] / cpu["seconds"]}
        line["parity"] = {"max_rel_err_vs_cpu": rel, "n": cpu["n"], "gate": 1e-9, "ok": rel < 1e-9}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def _ncu_traffic(workload):
    """DRAM bytes of one launch from the committed ncu capture of this workload (None if there is none)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[workload]["bytes"]
    except Exception:
        return None


def _hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0   # fallback stated in B200_PROFILING.md


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--profiles", type=int, default=0, help="override profiles per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.profiles:
        wl["P"] = args.profiles
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
    else:
        run_gpu_arm(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
