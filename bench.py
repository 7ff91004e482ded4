#!/usr/bin/env python
"""bench.py — the hot-path benchmark contract.

One "step" = what a user of the reference does for BASELINE.json's headline metric:

    fit_model = model.fit(dataset)           # Gram build + factorisation + information solve
    nll       = -model.log_likelihood(dataset)   # a second, independent Gram build + factorisation

on the configs[2] workload (exact GP, N = 65 536, fp64, 3-D features U[0,10]^3,
SquaredExponential(1,1) + IndependentNoise(0.1), y = sin x0 + 0.1 cos 10 x0), exactly as the
reference sequences it (gp.hpp:285-294 and gp.hpp:443-451: two Gram builds, two factorisations).

  value  = useful fp64 FLOP/s of the whole job, 2 * N^3/3 per step (the two factorisations; the
           O(N^2) Gram/solve work is not counted) / device time, inputs resident in HBM;
  e2e    = the same metric through the host-pointer C ABI (ab_gp_fit + ab_gp_nll): features and
           targets are copied host->device and information/nll device->host inside the timed region;
  roofline      = the trailing-update DGEMM/DSYRK kernel (gemm_nt_tma_kernel: TMA + mbarrier fed DMMA), FP64
                  tensor bound: one isolated 8192^3 launch timed live with CUDA events (2*8192^3 flops /
                  launch time) vs the cuBLAS DGEMM rate measured live on the same GPU (MEASURED_PEAKS.json
                  has no fp64 figure); `phase` repeats the ratio for the whole factorisation phase
                  (N^3/3 flops / CUDA-event time, panel kernels included); `traffic` is the ncu DRAM
                  byte count of that launch for this build (profiles/r02b_ncu_summary.md);
  roofline_gram = the Gram kernel at configs[1] (N = 32 768, SE + Matern52, full symmetric store),
                  HBM bound: (8 N^2 + 8 N D) bytes / launch time vs MEASURED_PEAKS.json hbm_gbs;
  cpu_baseline  = the reference's own Eigen path (oracle/_ref, or the C port) on the host cores on
                  a bounded sample of the same workload (smaller N), same metric.

  configs       = device timings of BASELINE configs[0], [3], [4] (sinc N = 1000, LOO-CV N = 32 768, sparse GP
                  N = 2^20 / M = 4096) measured once after the timed region (N = 1), and
                  strong_scaling_anchor = one N = 131 072 fit on this single GPU (the matrix is 137 GB).

`--impl reference` times the reference CPU implementation alone (rank 0 only); after its steps it adds one
pass at N = 8192 and the fitted c * N^3 model, so that the same-config ratio can be extrapolated from the record.
Multi-GPU (--gpus N under torchrun, N > 1): BASELINE configs[2]'s multi-GPU leg — the same step at
N = 131 072 on ONE matrix sharded over the N ranks: Gram generated in block-column-cyclic layout,
right-looking Cholesky with NCCL panel broadcasts over NVLink (ab_dist_gp_fit), block substitution
solves.  Total work is fixed for N = 2/4/8 ("strong"); the metric (whole-job FLOP/s) is comparable
with the 1-GPU line.  `--dist-n` changes the size; `--replicas` instead runs N independent N=65 536
replicas (one hyper-parameter evaluation per GPU, as the tuner's finite-difference gradient does).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SE, NOISE, SUM, M52 = 1, 6, 7, 4
OPS_FIT = [SE, NOISE, SUM]
PARAMS_FIT = [1.0, 1.0, 0.1, 0.0, 0.0, 0.0]
OPS_GRAM = [SE, M52, SUM]
PARAMS_GRAM = [2.0, 1.5, 3.0, 0.7, 0.0, 0.0]


def make_data(n, dim=3, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 10.0, size=(n, dim))
    y = np.sin(x[:, 0]) + 0.1 * np.cos(10.0 * x[:, 0])
    return x, y


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                     "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in out.stdout.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples),
                "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# reference arm (CPU)
# ---------------------------------------------------------------------------------------------

def cpu_step(n, kind_pref="reference"):
    """One fit + log_likelihood of the reference on the host at size n.  Returns (seconds, kind)."""
    from oracle.oracle import Ref, Restate

    x, y = make_data(n)
    if kind_pref == "reference" and Ref.available():
        t0 = time.perf_counter()
        Ref.gp_fit(6, [1.0, 1.0, 0.1], x, y, nthreads=os.cpu_count() or 1)
        Ref.gp_nll(6, [1.0, 1.0, 0.1], x, y)
        return time.perf_counter() - t0, "reference"
    t0 = time.perf_counter()
    Restate.gp_fit(OPS_FIT, PARAMS_FIT, x, y)
    Restate.gp_nll(OPS_FIT, PARAMS_FIT, x, y)
    return time.perf_counter() - t0, "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_n
    times = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        dt, kind = cpu_step(n)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    value = 2.0 * n ** 3 / 3.0 / t * 1e-12
    cores = os.cpu_count() or 1
    # one larger pass: the reference's unblocked LDLT is O(N^3) with a rate that FALLS with N (cache), so
    # the N = 4096 rate flatters it; c from the larger size is the honest extrapolation constant
    extrap = None
    if not args.no_cpu and args.ref_n2 > n:
        dt2, _ = cpu_step(args.ref_n2)
        c1, c2 = t / n ** 3, dt2 / args.ref_n2 ** 3
        extrap = {"n": [n, args.ref_n2], "seconds_per_step": [t, dt2], "c_seconds_per_n3": [c1, c2],
                  "extrapolated_seconds_per_step_at_n65536": c2 * 65536.0 ** 3,
                  "note": "fit + log_likelihood = 2 factorisations; t = c * N^3, c taken at the larger size"}
    line = {
        "impl": "reference", "metric": "gp_fit_plus_nll_fp64_tflops", "value": value,
        "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"exact GP fit + log_likelihood, N={n} (bounded CPU sample of the "
                               "N=65536 workload), 3-D U[0,10]^3, SE(1,1)+IndependentNoise(0.1)",
                   "flops_per_step": "2*N^3/3"},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": cores if kind == "reference" else 1,
                         "kind": kind,
                         "sample": f"N={n}; Gram build threaded over {cores} cores, Eigen LDLT is "
                                   "single-threaded by construction"},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "extrapolation": extrap,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# device arm
# ---------------------------------------------------------------------------------------------

def bench_dataset(n, seed):
    """benchmarks/bench_utils.h:76-85: x ~ U[0,10], y = sin x + 0.1 cos 10x."""
    x = np.random.default_rng(seed).uniform(0.0, 10.0, size=n)
    return x, np.sin(x) + 0.1 * np.cos(10.0 * x)


def _guard(fn):
    try:
        return fn()
    except Exception as exc:  # an extra must never cost the headline line
        return {"error": repr(exc)[:300]}


def other_configs(h, capi):
    """Device timings (library CUDA events / wall clock of the host-pointer call) of BASELINE configs[0],
    [3] and [4], one pass each; parity for them is in tests/test_gpu_fullsize.py."""
    MARGINAL = capi.MARGINAL
    h.trim()

    def sinc():
        rng = np.random.default_rng(0)
        x = rng.uniform(-10.0, 23.0, size=1000)
        y = np.sqrt(2.0) * x + 3.14159 + 10.0 * np.sinc((x - 3.0) / np.pi) + rng.normal(size=1000)
        ops, pp = [SE, NOISE, SUM], [3.5, 5.7, 1.0, 0.0, 0.0, 0.0]
        t = np.linspace(-20.0, 33.0, 161)
        best = {}
        for _ in range(3):
            t0 = time.perf_counter()
            f, info = h.gp_fit(ops, pp, x, y)
            fit_ms = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            h.gp_predict(f, ops, pp, x, info, t, MARGINAL)
            pred_ms = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            nll = h.gp_nll(ops, pp, x, y)
            nll_ms = (time.perf_counter() - t0) * 1e3
            f.free()
            for k, v in (("fit_wall_ms", fit_ms), ("predict_marginal_wall_ms", pred_ms), ("nll_wall_ms", nll_ms)):
                best[k] = min(best.get(k, 1e30), v)
        best["nll"] = nll
        return best

    def loo():
        n = 32768
        x, y = bench_dataset(n, 27)
        f, info = h.gp_fit(OPS_FIT, PARAMS_FIT, x, y)
        out = {"n": n, "fit_factor_ms": h.timings()["factor_ms"]}
        _, off8, idx8 = capi.group_indexers(x.astype(np.int64) % 8)
        _, off1, idx1 = capi.group_indexers(np.arange(n, dtype=np.int64))
        for name, off, idx in (("grouped8", off8, idx8), ("loo", off1, idx1)):
            t0 = time.perf_counter()
            _, _, _, score = h.gp_cv(f, y, info, off, idx, MARGINAL, want_score=True)
            out[f"{name}_wall_ms"] = (time.perf_counter() - t0) * 1e3
            out[f"{name}_device_ms"] = h.timings()["total_ms"]
            out[f"{name}_score"] = score
        out["loo_TFLOPs"] = n ** 3 / 3.0 / out["loo_device_ms"] * 1e-9  # explicit triangular inverse: N^3/3
        f.free()
        h.trim()
        return out

    def sparse():
        n, m = 1 << 20, 4096
        x, y = bench_dataset(n, 0)
        u = np.linspace(x.min(), x.max(), m)
        out = {"n": n, "m": m}
        for name, keys in (("fitc", np.arange(n, dtype=np.int64)),
                           ("pitc1024", (x * (n / 10.0 / 1024.0)).astype(np.int64))):
            _, off, idx = capi.group_indexers(keys)
            t0 = time.perf_counter()
            f, _, ll = h.sparse_fit(OPS_FIT, PARAMS_FIT, x, y, u, off, idx)
            wall = (time.perf_counter() - t0) * 1e3
            tf = h.timings()
            t0 = time.perf_counter()
            f.predict(OPS_FIT, PARAMS_FIT, np.linspace(0.0, 10.0, 512), MARGINAL)
            pred = (time.perf_counter() - t0) * 1e3
            f.free()
            h.trim()
            out[name] = {"fit_wall_ms": wall, "fit_device_ms": tf["total_ms"], "factor_ms": tf["factor_ms"],
                         "h2d_ms": tf["h2d_ms"], "predict_marginal_p512_wall_ms": pred, "ll": ll,
                         # CholQR2: 2 x (SYRK + TRSM) of (N+M) M^2 each + P = L_u^-1 K_uf (N M^2)
                         "TFLOPs_5NM2": 5.0 * n * m * m / tf["total_ms"] * 1e-9}
        return out

    return {"0_sinc_n1000_p161": _guard(sinc), "3_loo_cv_n32768": _guard(loo),
            "4_sparse_n1048576_m4096": _guard(sparse)}


def strong_scaling_anchor(h, capi, n):
    """One fit of the multi-GPU leg's matrix (N = 131 072, 137 GB) on THIS single GPU: the 1-GPU point of
    the strong-scaling curve whose other points are the --gpus 2/4/8 lines."""
    def run():
        import torch
        torch.cuda.empty_cache()
        h.trim()
        free_b, _ = torch.cuda.mem_get_info()
        need = 8.0 * n * (n + 16) + 2e9
        if need > free_b:
            return {"n": n, "skipped": f"needs {need / 1e9:.0f} GB, {free_b / 1e9:.0f} GB free"}
        x, y = make_data(n, seed=0)
        fd, yd = h.upload_features(x), h.upload(y)
        f, info = h.gp_fit_d(OPS_FIT, PARAMS_FIT, fd, yd)
        t = h.timings()
        f.free(); info.free(); fd.free(); yd.free()
        h.trim()
        return {"n": n, "fit_device_ms": t["total_ms"], "gram_ms": t["gram_ms"], "factor_ms": t["factor_ms"],
                "solve_ms": t["solve_ms"], "factor_TFLOPs": n ** 3 / 3.0 / t["factor_ms"] * 1e-9,
                "fit_TFLOPs": n ** 3 / 3.0 / t["total_ms"] * 1e-9,
                "note": "same matrix as the --gpus 2/4/8 lines (their value counts 2 factorisations per step: "
                        "compare factor_TFLOPs / fit_TFLOPs with value)"}
    return _guard(run)


def run_device(args):
    import torch
    import torch.distributed as dist

    from albatross_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: albatross_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.n
    stream = torch.cuda.current_stream()
    h = capi.Handle(local_rank, stream=stream.cuda_stream)
    x, y = make_data(n, seed=rank)
    peaks, peak_src = measured_peaks()

    # ---- live fp64 peak probe: cuBLAS DGEMM through torch (the roofline denominator) ----------
    probe_n = 8192
    a = torch.randn(probe_n, probe_n, dtype=torch.float64, device="cuda")
    b = torch.randn(probe_n, probe_n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fp64_peak = 2.0 * probe_n ** 3 / (best * 1e-3) * 1e-12
    del a, b
    torch.cuda.empty_cache()

    # ---- device-resident arm -------------------------------------------------------------------
    fd = h.upload_features(x)
    yd = h.upload(y)

    def step_device():
        f, info = h.gp_fit_d(OPS_FIT, PARAMS_FIT, fd, yd)
        t_fit = h.timings()
        nll = h.gp_nll_d(OPS_FIT, PARAMS_FIT, fd, yd)
        t_nll = h.timings()
        f.free()
        info.free()
        return nll, t_fit, t_nll

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    h.reset_counters()
    barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(stream)
    phases = []
    nll = None
    for _ in range(args.steps):
        nll, t_fit, t_nll = step_device()
        phases.append((t_fit, t_nll))
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = h.timings()["kernel_launches"]
    sampler.stop_flag = True
    sampler.join()
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    flops_step = 2.0 * n ** 3 / 3.0
    value = world * flops_step * args.steps / (dev_ms * 1e-3) * 1e-12

    # ---- end-to-end arm: host buffers through the C ABI ---------------------------------------
    xp = torch.from_numpy(x).pin_memory().numpy()
    yp = torch.from_numpy(y).pin_memory().numpy()
    f, info = h.gp_fit(OPS_FIT, PARAMS_FIT, xp, yp)  # warm
    f.free()
    barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(stream)
    e2e_steps = max(1, args.steps)
    for _ in range(e2e_steps):
        f, info = h.gp_fit(OPS_FIT, PARAMS_FIT, xp, yp)
        nll_e2e = h.gp_nll(OPS_FIT, PARAMS_FIT, xp, yp)
        f.free()
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * flops_step * e2e_steps / (e2e_ms * 1e-3) * 1e-12
    h2d = 2 * (x.nbytes + y.nbytes)  # fit and nll each upload features + targets
    d2h = y.nbytes + 8               # information vector + nll scalar

    # ---- roofline of the dominant kernel (factorisation GEMM) -----------------------------------
    factor_ms = float(np.mean([p[0]["factor_ms"] + p[1]["factor_ms"] for p in phases]))
    gram_ms_fit = float(np.mean([p[0]["gram_ms"] + p[1]["gram_ms"] for p in phases]))
    solve_ms = float(np.mean([p[0]["solve_ms"] + p[1]["solve_ms"] + p[1]["reduce_ms"] for p in phases]))
    achieved_phase = flops_step / (factor_ms * 1e-3) * 1e-12
    # the dominant kernel alone: one 8192^3 NT launch (the shape of a trailing update) timed with the
    # library's CUDA events on its own stream
    f = None
    h.trim()
    gA, gB, gC = h.alloc(probe_n, probe_n), h.alloc(probe_n, probe_n), h.alloc(probe_n, probe_n)
    kernel_ms = 1e30
    for _ in range(4):
        h.gemm(gA, gB, gC, alpha=-1.0, beta=1.0, trans_a=False, trans_b=True, lower=False)
        kernel_ms = min(kernel_ms, h.timings()["factor_ms"])
    gA.free(); gB.free(); gC.free()
    h.trim()
    achieved = 2.0 * probe_n ** 3 / (kernel_ms * 1e-3) * 1e-12
    roofline = {"bound": "tensor",
                "kernel": "ab::gemm_nt_tma_kernel (TMA + mbarrier fed DMMA m8n8k4: the DSYRK / DGEMM of every "
                          "trailing update and TRSM), one 8192^3 launch",
                "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                "peak_source": f"cuBLAS DGEMM {probe_n}^3 via torch.matmul, measured live in this run "
                               "(MEASURED_PEAKS.json carries no fp64 figure)",
                "launch_ms": kernel_ms,
                "traffic": 1.2336e10,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full "
                                  "of this build (profiles/r02p_ncu_gemm_tma_grouped.md): 11.80 GB + 0.54 GB = "
                                  "6 % of DRAM peak (46.92 + 0.54 GB before the grouped tile order, "
                                  "profiles/r02b_ncu_summary.md); algorithmic minimum 3 * 8 * 8192^2 = 1.6 GB "
                                  "(tile re-reads beyond the 126 MB L2: tensor-bound, not memory-bound)",
                "phase": {"what": "whole factorisation phase of the step (2 x N^3/3 flops / CUDA-event time, "
                                  "panel kernels and TRSMs included)",
                          "achieved": achieved_phase, "frac": achieved_phase / fp64_peak}}

    # ---- Gram roofline at configs[1] -----------------------------------------------------------
    gram_line = None
    if rank == 0:
        f = None
        h.trim()
        ng = args.gram_n
        xg, _ = make_data(ng, seed=1)
        fg = h.upload_features(xg)
        best = 1e30
        flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
        for i in range(5):
            flush.fill_(float(i))  # > L2 write between timed launches
            K = h.gram_sym_d(OPS_GRAM, PARAMS_GRAM, fg)
            ms = h.timings()["gram_ms"]
            if i > 0:
                best = min(best, ms)
            K.free()
        gbytes = (8.0 * ng * ng + 8.0 * ng * 3) * 1e-9
        hbm = float(peaks["hbm_gbs"])
        gram_line = {"bound": "hbm", "kernel": "ab::gram_kernel<3,sym> SE+Matern52 full symmetric",
                     "n": ng, "achieved": gbytes / (best * 1e-3), "peak": hbm, "unit": "GB/s",
                     "frac": gbytes / (best * 1e-3) / hbm, "peak_source": peak_src,
                     "ms": best,
                     "algorithmic_bytes": gbytes * 1e9,
                     # not measured in this run: the ncu --set full capture of this build's kernel at
                     # n = 32768 (profiles/r02b_ncu_summary.md): 8.530 GB written + 0.060 GB read
                     "traffic": (8.590310e9 if ng == 32768 else None),
                     "traffic_source": ("ncu capture of this build (profiles/r02b_ncu_summary.md), not a live "
                                        "counter") if ng == 32768 else None}
        del flush
        fg.free()
        h.trim()

    # ---- once-only extras (rank 0, N = 1): the other BASELINE configs and the strong-scaling anchor --
    extras, anchor = None, None
    if rank == 0 and world == 1 and not args.no_extras:
        extras = other_configs(h, capi)
        anchor = strong_scaling_anchor(h, capi, args.dist_n)

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        dt, kind = cpu_step(args.ref_n)
        cores = os.cpu_count() or 1
        cpu = {"value": 2.0 * args.ref_n ** 3 / 3.0 / dt * 1e-12, "unit": "TFLOP/s",
               "cores": cores if kind == "reference" else 1, "kind": kind, "seconds": dt,
               "sample": f"fit + log_likelihood at N={args.ref_n} (bounded sample of the N={n} "
                         "workload; the reference's LDLT is O(N^3) single-threaded, "
                         f"extrapolated N={n}: {dt * (n / args.ref_n) ** 3:.0f} s)"}

    if rank == 0:
        line = {
            "metric": "gp_fit_plus_nll_fp64_tflops", "value": value, "unit": "TFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"exact GP fit + log_likelihood, N={n}, 3-D U[0,10]^3, "
                                   "SE(1,1)+IndependentNoise(0.1) (BASELINE configs[2]); "
                                   "two Gram builds + two factorisations per step, as the reference",
                       "flops_per_step": "2*N^3/3", "parallelism": f"replicas x{world}",
                       "l2_policy": "inputs (32 GiB matrix) larger than L2"},
            "nll": nll,
            "phase_ms": {"gram": gram_ms_fit, "factor": factor_ms, "solve_reduce": solve_ms},
            "e2e": {"value": e2e_value, "unit": "TFLOP/s", "ms_per_step": e2e_ms / e2e_steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "nll_matches_device_arm": bool(abs(nll_e2e - nll) <= 1e-9 * abs(nll))
                    if world == 1 else None},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roofline,
            "roofline_gram": gram_line,
            "cpu_baseline": cpu,
            "configs": extras,
            "strong_scaling_anchor": anchor,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def gather_ranks(value):
    """Per-rank list of a python float (max-over-ranks is the contract's reduction; the list shows balance)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    t = torch.zeros(world, dtype=torch.float64, device="cuda")
    t[dist.get_rank()] = float(value)
    dist.all_reduce(t)
    return [round(float(v), 3) for v in t.cpu()]


def dist_configs(h, capi, abd, rank, world):
    """Multi-GPU legs of BASELINE configs[3] and [4], one pass each after the timed region:
    LOO-CV N = 32 768 with L replicated by ONE ncclBroadcast and the folds sharded; sparse GP N = 2^20,
    M = 4096 with the observation groups sharded."""
    import torch
    import torch.distributed as dist

    MARGINAL = capi.MARGINAL
    h.trim()

    def loo():
        n = 32768
        x, y = bench_dataset(n, 27)
        out = {"n": n}
        f = info = None
        if rank == 0:
            f, info = h.gp_fit(OPS_FIT, PARAMS_FIT, x, y)
            out["fit_factor_ms_rank0"] = h.timings()["factor_ms"]
        info_t = torch.zeros(n, dtype=torch.float64, device="cuda")
        if rank == 0:
            info_t.copy_(torch.from_numpy(info))
        dist.broadcast(info_t, src=0)
        info = info_t.cpu().numpy()
        dist.barrier()
        t0 = time.perf_counter()
        f = h.dist_factor_broadcast(f, root=0)
        out["L_broadcast_wall_ms"] = abd.max_over_ranks((time.perf_counter() - t0) * 1e3)
        out["L_broadcast_GBps"] = 8.0 * n * (n + 16) / (out["L_broadcast_wall_ms"] * 1e-3) * 1e-9
        _, off8, idx8 = capi.group_indexers(x.astype(np.int64) % 8)
        _, off1, idx1 = capi.group_indexers(np.arange(n, dtype=np.int64))
        for name, off, idx in (("grouped8", off8, idx8), ("loo", off1, idx1)):
            dist.barrier()
            t0 = time.perf_counter()
            _, _, score = h.dist_gp_cv(f, y, info, off, idx, MARGINAL, want_score=True)
            out[f"{name}_wall_ms"] = abd.max_over_ranks((time.perf_counter() - t0) * 1e3)
            out[f"{name}_score"] = score
        f.free()
        h.trim()
        return out

    def sparse():
        n, m = 1 << 20, 4096
        x, y = bench_dataset(n, 0)
        u = np.linspace(x.min(), x.max(), m)
        out = {"n": n, "m": m}
        for name, keys in (("fitc", np.arange(n, dtype=np.int64)),
                           ("pitc1024", (x * (n / 10.0 / 1024.0)).astype(np.int64))):
            _, off, idx = capi.group_indexers(keys)
            xl, yl, _, lo, li = abd.shard_sparse_inputs(x, y, None, off, idx, rank, world)
            dist.barrier()
            t0 = time.perf_counter()
            f, _, ll = h.sparse_fit(OPS_FIT, PARAMS_FIT, xl, yl, u, lo, li)
            out[name] = {"fit_wall_ms": abd.max_over_ranks((time.perf_counter() - t0) * 1e3),
                         "fit_device_ms": abd.max_over_ranks(h.timings()["total_ms"]), "ll": ll}
            f.free()
            h.trim()
        return out

    return {"3_loo_cv_n32768": _guard(loo), "4_sparse_n1048576_m4096": _guard(sparse)}


def run_device_dist(args):
    """N > 1: one matrix of size --dist-n sharded over the ranks (ab_dist_gp_fit)."""
    import torch
    import torch.distributed as dist

    from albatross_b200 import capi
    from albatross_b200 import dist as abd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: albatross_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()
    h = capi.Handle(local_rank, stream=stream.cuda_stream)
    abd.bootstrap(h)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    n = args.dist_n
    # preflight (rank-symmetric, before any collective of the library): the block-column-cyclic shard, the
    # two packed panel buffers and 6 GB of slack must fit the free HBM of every rank, else halve N
    nb_eff = args.nb or (512 if world >= 4 else 1024)  # the library's default block-column width
    while n > 8192:
        nblk = (n + nb_eff - 1) // nb_eff
        need = 8.0 * n * ((nblk + world - 1) // world) * nb_eff + 2 * 8.0 * n * nb_eff + 6e9
        free_b, _ = torch.cuda.mem_get_info()
        if abd.max_over_ranks(1.0 if need > free_b else 0.0) == 0.0:
            break
        if rank == 0:
            print(f"bench.py: N={n} needs {need / 1e9:.0f} GB per rank, {free_b / 1e9:.0f} GB free: "
                  f"halving N", file=sys.stderr)
        n //= 2
    x, y = make_data(n, seed=0)  # identical on every rank: features are replicated, K is sharded
    xp = torch.from_numpy(x).pin_memory().numpy()
    yp = torch.from_numpy(y).pin_memory().numpy()

    # live fp64 peak probe (roofline denominator), as in the 1-GPU arm
    probe_n = 8192
    a = torch.randn(probe_n, probe_n, dtype=torch.float64, device="cuda")
    b = torch.randn(probe_n, probe_n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fp64_peak = 2.0 * probe_n ** 3 / (best * 1e-3) * 1e-12
    del a, b
    torch.cuda.empty_cache()

    def step():
        # model.fit(dataset): Gram + factorisation + information; then -log_likelihood: a second,
        # independent Gram + factorisation (gp.hpp:443-451), as the reference sequences it
        f, info, _ = h.dist_gp_fit(OPS_FIT, PARAMS_FIT, xp, yp, nb=args.nb)
        t_fit = h.timings()
        f.free()
        f, _, nll = h.dist_gp_fit(OPS_FIT, PARAMS_FIT, xp, yp, nb=args.nb, want_information=False)
        t_nll = h.timings()
        f.free()
        return nll, info, t_fit, t_nll

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    h.reset_counters()
    barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(stream)
    phases = []
    nll = None
    t_nll = None
    for _ in range(args.steps):
        nll, info, t_fit, t_nll = step()
        phases.append((t_fit, t_nll))
    e1.record(stream)
    barrier()
    dev_ms = abd.max_over_ranks(e0.elapsed_time(e1))
    launches = h.timings()["kernel_launches"]
    sampler.stop_flag = True
    sampler.join()
    flops_step = 2.0 * n ** 3 / 3.0
    value = flops_step * args.steps / (dev_ms * 1e-3) * 1e-12
    factor_ms = abd.max_over_ranks(
        float(np.mean([p[0]["factor_ms"] + p[1]["factor_ms"] for p in phases])))
    gram_ms = abd.max_over_ranks(float(np.mean([p[0]["gram_ms"] + p[1]["gram_ms"] for p in phases])))
    solve_ms = abd.max_over_ranks(
        float(np.mean([p[0]["solve_ms"] + p[1]["solve_ms"] for p in phases])))
    # every step of this leg IS end to end: ab_dist_gp_fit takes HOST pointers, so the timed region above
    # already contains the host->device copy of features / targets and the device->host read of the
    # information vector and the NLL in every step (bytes declared below); e2e repeats that figure
    e2e_ms = dev_ms / args.steps
    nll_e2e = nll
    wait_ms, panel_ms, nsteps = h.dist_fit_breakdown()
    breakdown = {"what": "last factorisation of the timed region, per rank, CUDA events on the update stream",
                 "steps": nsteps, "factor_ms": gather_ranks(t_nll["factor_ms"]),
                 "wait_for_panel_ms": gather_ranks(wait_ms), "panel_chain_ms_overlapped": gather_ranks(panel_ms),
                 "solve_ms": gather_ranks(t_nll["solve_ms"]), "gram_ms": gather_ranks(t_nll["gram_ms"])}
    achieved = flops_step / (factor_ms * 1e-3) * 1e-12
    line = None
    if rank == 0:
        line = {
            "metric": "gp_fit_plus_nll_fp64_tflops", "value": value, "unit": "TFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"exact GP fit + log_likelihood, N={n}, 3-D U[0,10]^3, "
                                   "SE(1,1)+IndependentNoise(0.1) (BASELINE configs[2], multi-GPU "
                                   "leg): one matrix, block-column-cyclic over the ranks, NCCL "
                                   "panel broadcasts; two Gram builds + two factorisations per step",
                       "flops_per_step": "2*N^3/3", "parallelism": f"block-cyclic 1x{world}, "
                                                                   f"nb={nb_eff}, decoupled panel pipeline, bulk updates in panel pairs (depth {2 * nb_eff})",
                       "l2_policy": "inputs (matrix shard >= 17 GB) larger than L2",
                       "scaling_note": "total work fixed for N_gpus = 2/4/8; the 1-GPU line is "
                                       "N=65536 (the same metric, FLOP/s, on 1/8 of the flops)"},
            "nll": nll,
            "phase_ms": {"gram": gram_ms, "factor": factor_ms, "solve": solve_ms},
            "e2e": {"value": flops_step / (e2e_ms * 1e-3) * 1e-12, "unit": "TFLOP/s",
                    "ms_per_step": e2e_ms, "h2d_bytes_per_step": 2 * (x.nbytes + y.nbytes),
                    "d2h_bytes_per_step": y.nbytes + 8,
                    "nll_matches_device_arm": bool(abs(nll_e2e - nll) <= 1e-12 * abs(nll)),
                    "note": "the timed steps themselves (host-pointer C ABI, copies inside every step)"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor",
                         "kernel": "ab::gemm_nt_tma_kernel (TMA + mbarrier fed DMMA trailing updates) inside "
                                   "ab_dist_gp_fit",
                         "achieved": achieved, "peak": fp64_peak * world, "unit": "TFLOP/s",
                         "frac": achieved / (fp64_peak * world),
                         "peak_source": f"{world} x cuBLAS DGEMM {probe_n}^3 measured live on rank 0",
                         "traffic": None,
                         "note": "achieved = 2*N^3/3 / max-over-ranks CUDA-event time of the two "
                                 "factorisation phases (panel kernels, packing and NCCL waits "
                                 "included)"},
            "cpu_baseline": None,
            "breakdown": breakdown,
            "configs_dist": None,
        }
    # once-only extras AFTER the headline numbers are final.  They contain collectives: a watchdog prints the
    # line without them and ends the process if a rank fails to reach one (never lose the headline to an extra)
    def bail():
        if rank == 0:
            line["configs_dist"] = {"error": "timed out (watchdog)"}
            print(json.dumps(line), flush=True)
        os._exit(0)

    if not args.no_extras:
        dog = threading.Timer(240.0, bail)
        dog.daemon = True
        dog.start()
        extras = dist_configs(h, capi, abd, rank, world)
        dog.cancel()
        if rank == 0:
            line["configs_dist"] = extras
    if rank == 0:
        print(json.dumps(line), flush=True)
    h.dist_finalize()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="device", choices=["device", "reference"])
    ap.add_argument("--n", type=int, default=65536, help="training-set size of the exact-GP step")
    ap.add_argument("--gram-n", type=int, default=32768)
    ap.add_argument("--ref-n", type=int, default=4096,
                    help="size of the bounded CPU sample (fit+ll is ~9.5e-11*N^3 s per pass)")
    ap.add_argument("--ref-n2", type=int, default=8192,
                    help="size of the one extra reference pass used for the c*N^3 extrapolation")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the once-only extras (other configs, N=131072 anchor)")
    ap.add_argument("--dist-n", type=int, default=131072,
                    help="matrix size of the multi-GPU (block-cyclic) step")
    ap.add_argument("--nb", type=int, default=0, help="block-column width (0 = library default)")
    ap.add_argument("--replicas", action="store_true",
                    help="N > 1: independent N=65536 replicas instead of one sharded matrix")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args)
    elif world > 1 and not args.replicas:
        run_device_dist(args)
    else:
        run_device(args)


if __name__ == "__main__":
    main()
