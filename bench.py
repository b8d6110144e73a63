#!/usr/bin/env python
"""Benchmark of the sfa-numpy modal beamforming hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[2], the config the headline metric is quoted on): rigid sphere,
order N=8, 64 microphones (grid_equal_angle(3)), 1024 frequency bins, 2562 look directions,
938 STFT frames (10 s at 48 kHz, NFFT 2048, hop 512), Tikhonov-regularised plane-wave
decomposition.  One "step" = the whole hot path over that batch: SH matrices at the mics and
look directions (stage 1), radial filters + fused regularised inverse (stage 2), and the per-bin
steering contraction (stage 3) producing F*L*T beam outputs.

Prints ONE JSON line (rank 0).  `value` = beam outputs/s with inputs resident in HBM (CUDA
events, max over ranks); `e2e` = the same metric through the host-buffer C ABI call
(sfa_pwd_host_run: pinned NumPy in, NumPy out, H2D + D2H inside the timed region).
With N > 1 every rank processes its own batch of F bins (frequency sharding, no data-path
collective): weak scaling.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1.  The CPU arm (rank 0: `--impl reference` and `cpu_baseline`) is
# meant to use every host core, and OpenBLAS fixes its pool size when NumPy is imported.
if os.environ.get("RANK", "0") == "0":
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

# stdout carries exactly ONE JSON line: everything else that writes to file descriptor 1 (NCCL prints its
# version banner there) is sent to stderr; the line itself goes to the saved descriptor.
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(name="cfg3: rigid sphere N=8, 64 mics (grid_equal_angle(3)), F=1024 bins, L=2562 look dirs, "
                "T=938 frames, Tikhonov a0=100",
           N=8, F=1024, L=2562, T=938, r=0.042, a0=100.0, method="Tikh", setup="rigid")


DTYPE_NOTE = {"c64": "c64 (3xTF32 on tcgen05, FP32 accumulate; stages 1-2 in f64)",
              "c64_fast": "c64 (TF32 hi*hi + BF16 cross terms on tcgen05, FP32 accumulate; stages 1-2 in f64)"}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """Polls nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] if power else sm
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core, so lift the
    BLAS thread limit at run time.  Returns the thread count in effect."""
    n = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
        info = threadpoolctl.threadpool_info()
        got = max([i.get("num_threads", 1) for i in info if i.get("user_api") == "blas"] or [1])
        return got
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", n))


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs next to GPU `index` (NVML's ideal affinity), so that the pinned host
    buffers of the e2e path are first-touched on the GPU's own NUMA node.  Returns the previous CPU set."""
    try:
        import pynvml
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return prev
    except Exception:
        return None


def make_inputs(O, cfg, rank):
    azi, colat, w = O.grid_equal_angle(3)                 # 64 microphones
    azl, col, _ = O.fibonacci_grid(cfg["L"])
    k = O.wavenumbers(cfg["F"])
    X = O.synth_spectra(cfg["F"], len(azi), cfg["T"], seed=rank)
    return azi, colat, w, azl, col, k, X


def run_reference(args):
    """CPU arm: the oracle port of the reference path (the reference is Python/NumPy/SciPy and
    cannot travel to the GPU box; oracle/ is its bit-exact restatement) on all host cores, on a
    bounded sample of the same workload."""
    from oracle import sfa_oracle as O
    cfg = CFG
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = use_all_host_threads()
    azi, colat, w, azl, col, k, _ = make_inputs(O, dict(cfg, F=1, T=1), 0)
    k = O.wavenumbers(cfg["F"])
    fs = 16                                               # bins per step (sample)
    X = O.synth_spectra(fs, len(azi), cfg["T"], seed=0).astype(np.complex128)
    t12s, t3s, times = [], [], []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        Y_p = O.sht_matrix(cfg["N"], azi, colat, w)
        Y_q = O.sht_matrix(cfg["N"], azl, col)
        bn = O.spherical_pw(cfg["N"], k, cfg["r"], cfg["setup"])
        dn, _ = O.regularize(1 / bn, cfg["a0"], cfg["method"])
        t1 = time.perf_counter()
        q = O.pwd_factored(Y_q, O.repeat_n_m(dn[:fs]), Y_p, X)
        t2 = time.perf_counter()
        if it >= args.warmup:
            t12s.append(t1 - t0)
            t3s.append(t2 - t1)
            times.append(t2 - t0)
    # whole-job throughput: stages 1-2 once per batch of F bins + stage 3 scaled from the sampled bins
    t_full = sum(t12s) / len(t12s) + (sum(t3s) / len(t3s)) * cfg["F"] / fs
    val = cfg["F"] * cfg["L"] * cfg["T"] / t_full
    t = sum(times)
    sample = ("per step: stages 1-2 for all %d bins (SciPy, 1 thread) + stage 3 on %d of %d bins (full L, T; "
              "complex128 NumPy/OpenBLAS factored form, all cores); value = F*L*T / (t_stage12 + t_stage3 * %d/%d)"
              % (cfg["F"], fs, cfg["F"], cfg["F"], fs))
    line = {"impl": "reference", "metric": "beam outputs/sec", "value": val, "unit": "outputs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": cfg["name"], "sample": sample},
            "cpu_baseline": {"value": val, "unit": "outputs/s", "cores": threads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": "outputs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def cpu_baseline(O, cfg, azi, colat, w, azl, col, k):
    """Oracle port on the host cores, bounded sample (~10-20 s)."""
    fs = 16
    threads = use_all_host_threads()
    X = O.synth_spectra(fs, len(azi), cfg["T"], seed=0).astype(np.complex128)
    t_stage12 = time.perf_counter()
    Y_p = O.sht_matrix(cfg["N"], azi, colat, w)
    Y_q = O.sht_matrix(cfg["N"], azl, col)
    bn = O.spherical_pw(cfg["N"], k, cfg["r"], cfg["setup"])
    dn, _ = O.regularize(1 / bn, cfg["a0"], cfg["method"])
    t_stage12 = time.perf_counter() - t_stage12
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        O.pwd_factored(Y_q, O.repeat_n_m(dn[:fs]), Y_p, X)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    outputs = fs * cfg["L"] * cfg["T"]
    # whole-path equivalent: stage 1/2 once per step of F bins + stage 3 scaled to F bins
    t_full = t_stage12 + best * cfg["F"] / fs
    return {"value": cfg["F"] * cfg["L"] * cfg["T"] / t_full, "unit": "outputs/s", "cores": threads,
            "kind": "port",
            "sample": "stage 3 on %d of %d bins (full L, T; complex128 NumPy/OpenBLAS factored form, best of 3, "
                      "%.3f s) scaled to %d bins + stage 1/2 once (%.3f s, SciPy, single thread)"
                      % (fs, cfg["F"], best, cfg["F"], t_stage12),
            "stage3_outputs_per_s": outputs / best, "stage12_s": t_stage12}


def parity_check(O, cfg, azi, colat, w, azl, col, k, Xh, out, bins):
    """Compare `bins` of the beam tensor the timed loop just wrote with the CPU oracle (complex128
    factored form).  Returns the worst normwise error per bin, max|q - ref| / max|ref|."""
    Y_p = O.sht_matrix(cfg["N"], azi, colat, w)
    Y_q = O.sht_matrix(cfg["N"], azl, col)
    bn = O.spherical_pw(cfg["N"], k, cfg["r"], cfg["setup"])
    dn, _ = O.regularize(1 / bn, cfg["a0"], cfg["method"])
    ref = O.pwd_factored(Y_q, O.repeat_n_m(dn)[bins], Y_p, Xh[bins].astype(np.complex128))
    got = out[bins].cpu().numpy()
    errs = [float(np.max(np.abs(got[i] - ref[i])) / np.max(np.abs(ref[i]))) for i in range(len(bins))]
    return {"bins": [int(b) for b in bins], "max_normwise_err": max(errs), "tolerance": 1e-5,
            "oracle": "oracle.sfa_oracle.pwd_factored (complex128 NumPy) on the same X, after the timed loop"}


def tensor_bound_configs(torch, micarray, O, dev, precision):
    """Stage 3 of the two tensor-bound BASELINE configs (cfg 4: N=16, 578 mics; cfg 5: N=32, 1089 mics) at
    full size -- the in-kernel K-loop GEMM (cgemm3x_kloop_sm100.cu) -- next to a cuBLAS TF32 GEMM timed
    on the same device.  `frac` = algorithmic real flops * passes / time / TF32 peak, passes = 2 (hybrid:
    one tf32 pass + one bf16 pass of the same duration) or 3 (3xTF32)."""
    A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
    a = torch.randn((8192, 8192), device=dev)
    b = torch.randn((8192, 8192), device=dev)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    best = 1e9
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = old
    tf32_peak = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    del a, b
    passes = 2 if precision == "c64_fast" else 3
    rows = []
    for name, N, grid, F, L, T, r, setup in (("cfg4", 16, "gauss", 1024, 2562, 64, 0.1, "card"),
                                             ("cfg5", 32, 1089, 4096, 16384, 32, 0.2, "rigid")):
        az, co, w = O.grid_gauss(N) if grid == "gauss" else O.fibonacci_grid(grid)
        azl, col, _ = O.fibonacci_grid(L)
        Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
        _, dn, _ = R.radial_filters(N, O.wavenumbers(F) + 0.1, r, setup, 100.0, "Tikh")
        Q, M = Y_p.shape
        X = torch.randn((F, Q, T), dtype=torch.complex64, device=dev)
        plan = S.PwdPlan(Y_q, dn, Y_p, T, precision=precision)
        out = S.empty_beams(F, L, T, torch.complex64, dev)
        for _ in range(2):
            plan(X, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            plan(X, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        flops = 8.0 * F * T * M * (Q + L)
        rows.append({"config": "%s: N=%d, %d mics, F=%d, L=%d, T=%d (factored form, K-loop GEMM)" % (name, N, Q, F, L, T),
                     "ms": ms, "outputs_per_s": F * L * T / (ms * 1e-3), "alg_tflops": flops / (ms * 1e-3) / 1e12,
                     "tensor_passes": passes, "cublas_tf32_tflops": tf32_peak,
                     "frac_of_tf32_peak": flops * passes / (ms * 1e-3) / 1e12 / tf32_peak})
        del X, out, plan, Y_p, Y_q, dn
        torch.cuda.empty_cache()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-stage12", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--precision", default="c64_fast", choices=["c64", "c64_fast"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from oracle import sfa_oracle as O            # synthetic-input generator + cpu_baseline only
    import sfa_numpy_b200 as micarray          # loads the prebuilt libsfa_b200.so; raises if it is missing
    from sfa_numpy_b200 import _lib
    import ctypes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    cfg = CFG
    A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering

    azi, colat, w, azl, col, k, Xh = make_inputs(O, cfg, rank)
    Q = len(azi)
    F, L, T, N = cfg["F"], cfg["L"], cfg["T"], cfg["N"]
    d_azi, d_col, d_w = (torch.from_numpy(a).to(dev) for a in (azi, colat, w))
    d_azl, d_coll = torch.from_numpy(azl).to(dev), torch.from_numpy(col).to(dev)
    d_k = torch.from_numpy(k).to(dev)
    X = torch.from_numpy(Xh).to(dev)
    out = S.empty_beams(F, L, T, torch.complex64, dev)
    plan = {}

    def step():
        Y_p = A.sht_matrix(N, d_azi, d_col, d_w)
        Y_q = A.sht_matrix(N, d_azl, d_coll)
        _, dn, _ = R.radial_filters(N, d_k, cfg["r"], cfg["setup"], cfg["a0"], cfg["method"], check="defer")
        if "p" not in plan:
            plan["p"] = S.PwdPlan(Y_q, dn, Y_p, T, precision=args.precision)
        p = plan["p"]
        p.Y_look, p.Y_mic, p.dn = Y_q, Y_p, dn
        return p(X, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.sfa_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    R.check_deferred()
    launches = lib.sfa_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * F * L * T * args.steps / (ms * 1e-3)
    # the output of the LAST timed step against the oracle: a fast step that computes something else is not a result
    check_bins = [0, 74, F // 2 + 5, F - 1]
    pcheck = parity_check(O, cfg, azi, colat, w, azl, col, k, Xh, out, check_bins)
    if not (pcheck["max_normwise_err"] <= pcheck["tolerance"]):
        raise SystemExit("bench: parity check failed: %r" % (pcheck,))

    # ---- dominant kernel: average launch duration with CUDA events on its launching stream ----
    lib.sfa_prof_enable(1)
    for _ in range(max(2, min(args.steps, 5))):
        step()
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_longlong()
    lib.sfa_prof_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.sfa_prof_enable(0)
    gemm_ms_per_step = tot.value / max(2, min(args.steps, 5))
    # algorithmic bytes of stage 3 in complex64 I/O = spectra read + beams written + the two SH
    # matrices + the radial filters.  The steering matrices A_f are this implementation's own intermediate
    # and do NOT count.  Algorithmic flops = 8 F L Q T (steering form, Q < M).
    M_modes, C_d = (N + 1) ** 2, N + 1
    alg_bytes = 8.0 * F * T * (Q + L) + 8.0 * (Q + L) * M_modes + 8.0 * F * C_d
    alg_flops = 8.0 * F * L * Q * T
    passes = 2 if args.precision == "c64_fast" else 3
    peaks, peak_src = load_peaks()
    # tensor-side denominator: measured dense bf16 throughput / 2 = TF32 rate (sustained: the kernel runs
    # inside a long step under the power cap); a tf32 pass and the bf16 pass of the hybrid mode take the same time
    tf32_peak = 0.5 * peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    t_hbm_ms = alg_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3
    t_tensor_ms = alg_flops * passes / (tf32_peak * 1e12) * 1e3
    achieved = alg_bytes / (gemm_ms_per_step * 1e-3) / 1e9
    n_prof = max(2, min(args.steps, 5))
    launches_per_step = cnt.value / n_prof
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (74 bins) from the committed ncu
    # capture profiles/r1_i_cgemm3x_hybrid_2stage_ring2.txt; bench.py itself never runs under ncu.
    NCU_TRAFFIC_74_BINS = 1.377538e9 + 0.242853e9
    step_ms = ms / args.steps
    roofline = {"kernel": "stage-3 steering kernel, %s (tcgen05)" % ("hybrid TF32+BF16" if args.precision == "c64_fast" else "3xTF32"),
                # `bound` / `frac` follow SURVEY 8(d): max(bytes / measured HBM, flops * passes / measured tensor) / t
                "bound": "hbm" if t_hbm_ms >= t_tensor_ms else "tensor",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"],
                "frac_hbm_kernel": t_hbm_ms / gemm_ms_per_step, "frac_hbm_step": t_hbm_ms / step_ms,
                "frac_tensor_kernel": t_tensor_ms / gemm_ms_per_step, "frac_tensor_step": t_tensor_ms / step_ms,
                "frac_max_rule_kernel": max(t_hbm_ms, t_tensor_ms) / gemm_ms_per_step,
                "frac_max_rule_step": max(t_hbm_ms, t_tensor_ms) / step_ms,
                "frac_step": t_hbm_ms / step_ms,
                "t_hbm_ms": t_hbm_ms, "t_tensor_ms": t_tensor_ms, "tensor_passes": passes,
                "tensor_peak_tflops": tf32_peak,
                "tensor_peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 rate)",
                "traffic": NCU_TRAFFIC_74_BINS if args.precision == "c64_fast" else None,
                "traffic_source": "ncu --set full, one 74-bin launch (profiles/r1_i_cgemm3x_hybrid_2stage_ring2.txt)",
                "alg_bytes_per_launch": alg_bytes / launches_per_step if launches_per_step else None,
                "kernel_ms_per_launch": gemm_ms_per_step / launches_per_step if launches_per_step else None,
                "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback",
                "launches_per_step": launches_per_step,
                "kernel_ms_per_step": gemm_ms_per_step, "kernel_share_of_step": gemm_ms_per_step / (ms / args.steps),
                "alg_bytes_per_step": alg_bytes, "alg_tflops_achieved": alg_flops / (gemm_ms_per_step * 1e-3) / 1e12}

    line = {"metric": "beam outputs/sec", "value": value, "unit": "outputs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_NOTE[args.precision],
            "data": "synthetic",
            "config": {"workload": cfg["name"], "bins_per_gpu": F, "outputs_per_step_per_gpu": F * L * T,
                       "l2": "inputs+outputs (20 GB) far larger than L2; no flush needed",
                       "parallelism": "frequency-sharded x%d, no collective" % world},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "parity_check": pcheck}

    # ---- the collective of the multi-GPU path (north_star: "only an all-gather of the output beam maps") ----
    # Not inside `value` (each rank's beams stay resident); timed separately with CUDA events, max over ranks:
    #  (a) beam power map mean_t|q|^2 [F, L] f32 per rank (one HBM-read-bound kernel) + NCCL all-gather;
    #  (b) NCCL all-gather of a bounded slab of the complex64 beam tensor itself (32 bins per rank).
    try:
        from sfa_numpy_b200 import dist as sdist

        def timed(fn, n=3):
            fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b) / n], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        pm_ms = timed(lambda: S.power_map(out))
        coll = {"power_map_kernel_ms": pm_ms, "power_map_read_GBps": 8.0 * F * L * T / (pm_ms * 1e-3) / 1e9}
        if world > 1:
            coll["power_map_allgather_ms"] = timed(lambda: sdist.gather_power_map(out, world * F))
            coll["power_map_allgather_bytes_per_rank"] = 4 * F * L
            nb = 32
            slab = out[:nb].contiguous()
            ag_ms = timed(lambda: sdist.all_gather_slabs(slab, world * nb))
            coll["beam_slab_allgather"] = {"bins_per_rank": nb, "bytes_per_rank": int(slab.numel() * 8), "ms": ag_ms,
                                           "ingest_GBps_per_gpu": (world - 1) * slab.numel() * 8 / (ag_ms * 1e-3) / 1e9}
            del slab
        line["collective"] = coll
    except Exception as e:
        line["collective"] = {"error": repr(e)}

    # ---- e2e through the host-buffer C ABI: every rank pushes its own batch through its own PCIe link ----
    if not args.no_e2e:
        host, err = None, None
        prev_cpus = bind_to_gpu_numa_node(local) if world > 1 else None
        try:
            torch.cuda.synchronize()
            Yq_h = A.sht_matrix(N, d_azl, d_coll).cpu().numpy()
            Yp_h = A.sht_matrix(N, d_azi, d_col, d_w).cpu().numpy()
            dn_h = R.radial_filters(N, d_k, cfg["r"], cfg["setup"], cfg["a0"], cfg["method"])[1].cpu().numpy()
            del out
            torch.cuda.empty_cache()
            host = S.PwdHost(Yq_h, dn_h, Yp_h, T, precision=args.precision, device=local, chunk_bins=0)
            Xp = host.pinned_empty((F, Q, T))
            Xp[...] = Xh
            outp = host.pinned_empty((F, L, T))
            host(Xp, out=outp)                          # warm-up
        except Exception as e:                           # e.g. not enough pinned host memory
            err = repr(e)
        ok = torch.tensor([0 if err else 1], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)    # all ranks take the same branch (no barrier mismatch)
        if int(ok.item()) == 1:
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                host(Xp, out=outp)
            dt = (time.perf_counter() - t0) / args.e2e_steps
            td = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dt = float(td.item())
            line["e2e"] = {"value": world * F * L * T / dt, "unit": "outputs/s",
                           "h2d_bytes_per_step": int(Xp.nbytes) * world, "d2h_bytes_per_step": int(outp.nbytes) * world,
                           "ms_per_step": dt * 1e3,
                           "api": "sfa_pwd_host_run (C ABI, pinned host buffers, chunked H2D/compute/D2H pipeline), "
                                  "one call per rank, wall clock, max over ranks",
                           "steps": args.e2e_steps}
        else:
            line["e2e"] = {"value": None, "unit": "outputs/s", "error": err or "another rank failed to set up"}
        if host is not None:
            host.close()
        if prev_cpus:
            os.sched_setaffinity(0, prev_cpus)           # the CPU baseline below uses every core again
    if rank == 0 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(O, cfg, azi, colat, w, azl, col, k)
    if rank == 0 and not args.no_stage12:
        # second half of the BASELINE metric: SH / Bessel evals per second (stages 1 and 2)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_stage12
            rows = bench_stage12.run(cpu=not args.no_cpu, hbm_gbs=peaks["hbm_gbs"])
            line["stage12"] = [{k2: (round(v, 6) if isinstance(v, float) and abs(v) < 1e3 else v)
                                for k2, v in r.items()} for r in rows]
        except Exception as e:
            line["stage12"] = {"error": repr(e)}
    if rank == 0 and not args.no_other_configs:
        try:
            torch.cuda.empty_cache()
            line["tensor_bound_configs"] = tensor_bound_configs(torch, micarray, O, dev, args.precision)
        except Exception as e:
            line["tensor_bound_configs"] = {"error": repr(e)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()                 # rank 0 may still be timing its CPU / single-GPU extras
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
