#!/usr/bin/env python
"""Benchmark of the bundle-adjustment hot path (BASELINE.json metric: LM iterations/s and
residual+Jacobian observations/s on the 24-camera x 1 M-point rig, 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --steps K --warmup W    # CPU path (scipy least_squares)

A step is ONE trust-region (LM) outer iteration over the whole observation set:
linearise -> reduced camera system -> Cholesky -> back-substitution -> trial cost ->
on-device accept/reject (reference equivalent: one pass of scipy trf.py:465-578).
N > 1 shards the SAME problem by point (strong scaling, NCCL all-reduce of the reduced
camera system); launch with torchrun (RANK/LOCAL_RANK/WORLD_SIZE in the env).
Prints one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "LM_iters_per_sec"
UNIT = "iter/s"
FP64_PEAK_FALLBACK_TFLOPS = 37.1   # highest FP64 rate measured on this pool (profiles/r01_fp64_peak.txt)
HBM_FALLBACK_GBS = 6650.0          # B200_PROFILING.md fallback


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.01)
        except Exception as e:  # NVML missing: report that, never fail the bench
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def pinned(a):
    """Copy a numpy array into page-locked host memory (returns a numpy view of it)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy(), t


def schur_flops(k_hist_counts, n_obs):
    """Algorithmic FP64 flop of one Schur launch: per point with k views
    2*363*k(k+1)/2 (S update, lower triangle) + 2*99*k (W V^-1) + 2*33*k (rhs), plus
    ~300 flop/observation of model + Jacobian."""
    k = np.arange(k_hist_counts.size, dtype=np.float64)
    per_point = 726.0 * k * (k + 1) / 2 + 198.0 * k + 66.0 * k
    return float(np.sum(per_point * k_hist_counts) + 300.0 * n_obs)


def cpu_reference_run(rig, n_points, p_vis, steps, warmup, n_full):
    """The reference CPU path (oracle.bundle_adjust == PySBA.bundleAdjust through scipy
    least_squares: TRF + 3-point FD Jacobian + LSMR) on a bounded sample of the workload.
    Times `steps` outer iterations after `warmup` via the iteration callback."""
    from lasercalib_b200.synth import make_rig
    from oracle import pysba_oracle as O
    pb = make_rig(rig, n_points, seed=0, variant="volume", p_vis=p_vis)
    stamps = []

    class Stop(Exception):
        pass

    def cb(intermediate_result):
        stamps.append(time.perf_counter())
        if len(stamps) >= warmup + steps + 1:
            raise StopIteration

    t0 = time.perf_counter()
    try:
        O.bundle_adjust(pb["cams0"], pb["pts0"], pb["points_2d"], pb["camera_ind"], pb["point_ind"],
                        ftol=None, xtol=None, gtol=1e-15, max_nfev=10 * (warmup + steps + 2),
                        callback=cb)
    except StopIteration:
        pass
    if len(stamps) >= warmup + 2:
        done = len(stamps) - 1 - warmup
        dt = stamps[-1] - stamps[warmup]
    else:   # terminated early: fall back to the whole call
        done = max(1, len(stamps))
        dt = time.perf_counter() - t0
    it_s_sample = done / dt
    n_sample = pb["n_obs"]
    try:
        from threadpoolctl import threadpool_info
        blas_threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        blas_threads = os.cpu_count()
    return dict(value=it_s_sample * n_sample / n_full, unit=UNIT, cores=blas_threads, kind="port",
                sample="%s rig, %d points (%d obs): %d outer iterations in %.1f s = %.4f iter/s; "
                       "scaled linearly in observations to %d obs" %
                       (rig, pb["n_points"], n_sample, done, dt, it_s_sample, n_full),
                host_cpus=os.cpu_count(), sample_iters_per_s=it_s_sample, sample_obs=n_sample)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rig", default="ring24")
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--pvis", type=float, default=1.0)
    ap.add_argument("--cpu-points", type=int, default=12000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    K, W = args.steps, max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    workload = "%s: %d cameras x %d laser points, p_vis=%.2f, volume variant, seed 0 (BASELINE.json configs[2])" \
        % (args.rig, {"ring24": 24}.get(args.rig, 0), args.points, args.pvis)

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        from lasercalib_b200.synth import make_rig
        full = make_rig(args.rig, args.points, seed=0, variant="volume", p_vis=args.pvis)
        n_full = full["n_obs"]
        del full
        cb = cpu_reference_run(args.rig, args.cpu_points, args.pvis, K, min(W, 3), n_full)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
                "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "n_obs": n_full},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ this engine
    import torch
    from lasercalib_b200 import dist as D
    from lasercalib_b200._cabi import Engine
    from lasercalib_b200.pySBA import PySBA
    from lasercalib_b200.synth import make_rig
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference)")
    rank, ws, local = D.init_from_env()
    if ws == 1:
        local = 0
    torch.cuda.set_device(local)
    import torch.distributed as tdist

    def barrier():
        torch.cuda.synchronize()
        if ws > 1:
            tdist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if ws == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    pb = make_rig(args.rig, args.points, seed=0, variant="volume", p_vis=args.pvis)
    C, P, N = pb["n_cams"], pb["n_points"], pb["n_obs"]
    k_hist = np.bincount(np.bincount(pb["point_ind"], minlength=P))
    sh = D.shard_problem(pb["pts0"], pb["points_2d"], pb["camera_ind"], pb["point_ind"], None, rank, ws)

    # ---- device-resident run (value) ----
    eng = Engine(local)
    eng.set_problem(pb["cams0"], sh["pts"], sh["points_2d"], sh["camera_ind"], sh["point_ind"],
                    pt_offset=sh["pt_offset"])
    if ws > 1:
        D.connect_engine(eng)
    # Exactly K outer iterations of the real trajectory: bundleAdjust(1e-4) from the standard
    # perturbed x0 terminates after a handful of iterations, so the solve is repeated from x0
    # (parameters re-uploaded outside the device-timed region) until K iterations are done.
    tol = dict(ftol=1e-4, xtol=1e-8, gtol=1e-8, max_nfev=200)

    def run_steps(k, profile=False):
        done, ms, launches, last = 0, 0.0, 0, None
        while done < k:
            eng.set_params(pb["cams0"], sh["pts"])
            r, _ = eng.solve(max_iterations=k - done, profile=profile, **tol)
            if r.iterations <= 0:
                raise RuntimeError("solver made no iteration")
            done += int(r.iterations)
            ms += r.solve_ms
            launches += int(r.gpu_launches)
            last = r
        return done, ms, launches, last

    if W > 0:
        run_steps(W)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    steps_done, dev_ms, launches, res = run_steps(K)
    barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(dev_ms)
    ms_per_step = total_ms / max(1, steps_done)
    value = 1e3 / ms_per_step
    final_cost, nfev = res.cost, int(res.nfev)

    # ---- per-kernel times of the same K steps (CUDA events around every launch) ----
    barrier()
    prof = {}
    pdone = 0
    while pdone < K:
        eng.set_params(pb["cams0"], sh["pts"])
        r, _ = eng.solve(max_iterations=K - pdone, profile=True, **tol)
        pdone += int(r.iterations)
        for kname, v in eng.profile().items():
            a = prof.setdefault(kname, dict(launches=0, total_ms=0.0))
            a["launches"] += v["launches"]
            a["total_ms"] += v["total_ms"]
    barrier()
    # ---- residual + Jacobian blocks materialised (M1), device resident ----
    ms_jac = max_over_ranks(eng.time_device(1, 10))
    ms_res = max_over_ranks(eng.time_device(0, 10))
    peaks, peak_src = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
    fp64_peak = FP64_PEAK_FALLBACK_TFLOPS
    n_loc, p_loc = sh["point_ind"].size, sh["pts"].shape[0]
    k_loc = np.bincount(np.bincount(sh["point_ind"] - sh["pt_offset"], minlength=p_loc))
    roof = None
    if "schur" in prof:
        # dense rigs: the Schur complement runs on the FP64 tensor path (k_schur_mma) and the
        # camera blocks U in k_cam_normal; their time is counted together against the same
        # algorithmic flop count the single DFMA kernel (sparse rigs) is measured with
        mma = "cam_normal" in prof
        ms_schur = prof["schur"]["total_ms"] / prof["schur"]["launches"]
        ms_u = prof["cam_normal"]["total_ms"] / prof["cam_normal"]["launches"] if mma else 0.0
        ms_y = prof["make_Y"]["total_ms"] / prof["make_Y"]["launches"] if "make_Y" in prof else 0.0
        ms_u += ms_y          # helper passes of the tensor path: Y to HBM + camera blocks
        fl = schur_flops(k_loc, n_loc)
        kern_total = sum(v["total_ms"] for v in prof.values())
        # bound: the dense path runs on the tensor pipe (FP64 DMMA); its peak is the FP64 DMMA rate
        # measured on this pool (MEASURED_PEAKS.json only has the bf16 tensor rate), see peak_source
        roof = {"kernel": "k_schur_mma (+ k_make_Y + k_cam_normal)" if mma else "k_schur",
                "bound": "tensor" if mma else "fp64", "precision": "fp64",
                "achieved": fl / (ms_schur + ms_u) * 1e-9,
                "peak": fp64_peak, "unit": "TFLOP/s", "frac": fl / (ms_schur + ms_u) * 1e-9 / fp64_peak,
                "traffic": None, "flop_per_launch": fl, "ms_per_launch": ms_schur + ms_u,
                "ms_k_schur": ms_schur, "ms_helpers": ms_u, "ms_k_make_Y": ms_y,
                "frac_k_schur_alone": fl / ms_schur * 1e-9 / fp64_peak,
                "share_of_step": (prof["schur"]["total_ms"] + (prof["cam_normal"]["total_ms"] if mma else 0.0)
                                  + (prof["make_Y"]["total_ms"] if "make_Y" in prof else 0.0)) / kern_total,
                "peak_source": "FP64 DMMA (mma.sync m8n8k4) / DFMA microbenchmarks on this pool "
                               "(tools/fp64_peak.cu, profiles/r01_fp64_peak.txt): 37.1 TFLOP/s"}
    jac_bytes = 264.0 * n_loc + 24.0 * p_loc
    roof_m1 = {"kernel": "k_jacobian_blocks", "bound": "hbm", "achieved": jac_bytes / ms_jac * 1e-6,
               "peak": hbm_peak, "unit": "GB/s", "frac": jac_bytes / ms_jac * 1e-6 / hbm_peak,
               "traffic": None, "bytes_per_launch": jac_bytes, "ms_per_launch": ms_jac,
               "obs_per_s": N / ms_jac * 1e3, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)"}
    eng.close()
    try:   # dram bytes per launch from the committed `ncu --set full` capture of this workload
        with open(os.path.join(REPO, "profiles", "r01_traffic.json")) as f:
            tr = json.load(f)
        if ws == 1 and tr.get("n_obs") == N:
            t = tr["dram_bytes_per_launch"]
            if roof is not None:
                roof["traffic"] = ((t.get("k_schur_mma", 0) + t.get("k_cam_normal", 0) + t.get("k_make_Y", 0))
                                   or t.get("void k_schur<0>"))
            roof_m1["traffic"] = t.get("k_jacobian_blocks")
    except Exception:
        pass

    # ---- end to end through the public API with host buffers (e2e) ----
    cams_h, _k0 = pinned(pb["cams0"])
    pts_h, _k1 = pinned(pb["pts0"])
    p2_h, _k2 = pinned(pb["points_2d"])
    ci_h, _k3 = pinned(pb["camera_ind"])
    pi_h, _k4 = pinned(pb["point_ind"])
    sba = PySBA(cams_h.copy(), pts_h.copy(), p2_h, ci_h, pi_h)
    sba.bundleAdjust(1e-4, verbose=0, max_iterations=1, max_nfev=200)     # warm the path
    barrier()
    e2e_iters, e2e_calls = 0, 0
    t0 = time.perf_counter()
    while e2e_iters < K:
        sba = PySBA(cams_h, pts_h, p2_h, ci_h, pi_h)          # fresh object: full ingest per call
        r2 = sba.bundleAdjust(1e-4, verbose=0, max_iterations=K - e2e_iters, max_nfev=200)
        _ = float(r2.cost) + float(sba.cameraArray[0, 0])
        e2e_iters += int(r2["nit"])
        e2e_calls += 1
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = (C * 11 + sh["pts"].size) * 8 + n_loc * (16 + 8 + 8)
    d2h = (C * 11 + sh["pts"].size) * 8
    e2e = {"value": e2e_iters / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": h2d * e2e_calls / max(1, e2e_iters),
           "d2h_bytes_per_step": d2h * e2e_calls / max(1, e2e_iters), "seconds_total": e2e_s,
           "steps": e2e_iters, "calls": e2e_calls,
           "note": "PySBA(...).bundleAdjust(1e-4) on pinned host numpy arrays, repeated from x0 until K "
                   "iterations: every call = ingest (H2D, validate, narrow, CSR/masks) + its "
                   "iterations + parameters D2H; wall clock, max over ranks"}

    # ---- secondary workload (N = 1 only): the same rig at 50 % Bernoulli visibility ----
    secondary = None
    if ws == 1 and args.pvis == 1.0:
        pb2 = make_rig(args.rig, args.points, seed=0, variant="volume", p_vis=0.5)
        e2 = Engine(local)
        e2.set_problem(pb2["cams0"], pb2["pts0"], pb2["points_2d"], pb2["camera_ind"], pb2["point_ind"])
        done2, ms2 = 0, 0.0
        for phase in ("warm", "timed"):
            done2, ms2 = 0, 0.0
            while done2 < max(4, K // 2):
                e2.set_params(pb2["cams0"], pb2["pts0"])
                r, _ = e2.solve(max_iterations=max(4, K // 2) - done2, **tol)
                done2 += int(r.iterations)
                ms2 += r.solve_ms
        secondary = {"workload": "same rig, p_vis=0.50 (%d points kept, %d obs, %.1f views/point)"
                                 % (pb2["n_points"], pb2["n_obs"], pb2["n_obs"] / pb2["n_points"]),
                     "value": done2 / ms2 * 1e3, "unit": UNIT, "ms_per_step": ms2 / done2,
                     "steps": done2}
        e2.close()
        del pb2

    if rank != 0:
        return 0
    cpu_baseline = None
    if ws == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_reference_run(args.rig, args.cpu_points, args.pvis, 3, 1, N)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": steps_done,
            "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "n_cams": C, "n_points": P, "n_obs": N,
                       "mean_views_per_point": N / P, "parallelism": "point-shard x%d" % ws,
                       "l2": "inputs (%.0f MB of observations) exceed the 126 MB L2; no explicit flush"
                             % (N * 28 / 1e6),
                       "solver": "scipy-TRF semantics, exact Schur/Cholesky Gauss-Newton direction"},
            "resjac_obs_per_s": N / ms_jac * 1e3, "residual_obs_per_s": N / ms_res * 1e3,
            "roofline": roof, "roofline_m1": roof_m1,
            "hbm_frac_iteration": None, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
            "kernels_ms_per_step": {k: v["total_ms"] / max(1, steps_done) for k, v in prof.items()},
            "final_cost": final_cost, "nfev": nfev, "secondary": secondary}
    # algorithmic HBM bytes of one iteration (SURVEY 8d B_M2, 4 streaming passes + Schur)
    b_m2 = 4 * 24.0 * N + (4 * 24 + 24) * P
    line["hbm_frac_iteration"] = b_m2 / (ms_per_step * 1e-3) / 1e9 / hbm_peak
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
