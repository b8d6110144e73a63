#!/usr/bin/env python
"""Benchmark of the sfa-numpy modal beamforming hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[2], the config the headline metric is quoted on): rigid sphere,
order N=8, 64 microphones (grid_equal_angle(3)), 1024 frequency bins, 2562 look directions,
938 STFT frames (10 s at 48 kHz, NFFT 2048, hop 512), Tikhonov-regularised plane-wave
decomposition.  One "step" = the whole hot path over that batch: SH matrices at the mics and
look directions (stage 1), radial filters + fused regularised inverse (stage 2), and the per-bin
steering contraction (stage 3) producing F*L*T beam outputs -- `pipeline.ModalBeamformer.run()`,
one CUDA-graph launch.

Prints ONE JSON line (rank 0).  `value` = beam outputs/s with inputs resident in HBM (CUDA
events, max over ranks); `e2e` = the same metric through the host-buffer C ABI call
(sfa_pwd_host_run: pinned NumPy in, NumPy out, H2D + D2H inside the timed region).

N > 1 (torchrun, one rank per GPU): STRONG scaling of the same problem.  The F bins are sharded over
the ranks (stage 1-2 replicated), every rank computes its slab, and the beam power maps
mean_t|q|^2 [F, L] are all-gathered over NCCL inside the timed step (`value`).  The line also carries
the same step without any collective and with the all-gather of the complex beam tensor itself
(chunk-overlapped, `dist.OverlappedPwd`), and the three variants for BASELINE config 5.
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


_EMIT_LOCK = threading.Lock()
_EMITTED = [False]


def emit(line):
    """Write THE JSON line (once: the watchdog and the normal path may race)."""
    with _EMIT_LOCK:
        if _EMITTED[0]:
            return
        _EMITTED[0] = True
        os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())


def start_watchdog(seconds, line, rank):
    """Safety net for the sections after the headline measurement (multi-GPU variants, e2e, extra rows): if they
    have not finished after `seconds`, rank 0 prints the line with what it has, and every rank leaves.  A hang in
    an optional section must not cost the measured headline."""
    def fire():
        if rank == 0:
            for _ in range(5):                       # the main thread may be adding a key right now
                try:
                    snap = dict(line)
                    snap["watchdog"] = ("sections after the headline measurement did not finish within %d s; "
                                        "line emitted as is" % seconds)
                    json.dumps(snap)
                    break
                except Exception:
                    time.sleep(0.05)
            emit(snap)
        os._exit(0)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(name="cfg3: rigid sphere N=8, 64 mics (grid_equal_angle(3)), F=1024 bins, L=2562 look dirs, "
                "T=938 frames, Tikhonov a0=100",
           N=8, F=1024, L=2562, T=938, Q=64, r=0.042, a0=100.0, method="Tikh", setup="rigid")


DTYPE_NOTE = {"c64": "c64 (3xTF32 on tcgen05, FP32 accumulate; stages 1-2 in f64)",
              "c64_fast": "c64 (TF32 hi*hi + BF16 cross terms on tcgen05, FP32 accumulate; stages 1-2 in f64)"}


def config_dict(cfg, world):
    """The `config` object of the JSON line -- the SAME dict for the GPU arm and the reference (CPU) arm."""
    par = ("single GPU" if world == 1 else
           "one problem, F bins sharded over %d GPUs (strong scaling), stage 1-2 replicated, "
           "NCCL all-gather of the beam power maps inside the timed step" % world)
    return {"workload": cfg["name"], "N": cfg["N"], "Q": cfg["Q"], "F": cfg["F"], "L": cfg["L"], "T": cfg["T"],
            "outputs_per_step": cfg["F"] * cfg["L"] * cfg["T"],
            "l2": "inputs+outputs per step (20 GB) far larger than L2; no flush needed",
            "parallelism": par}


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
                 "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


def make_geometry(O, cfg):
    azi, colat, w = O.grid_equal_angle(3)                 # 64 microphones
    azl, col, _ = O.fibonacci_grid(cfg["L"])
    k = O.wavenumbers(cfg["F"])
    return azi, colat, w, azl, col, k


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_stage12(O, cfg, azi, colat, w, azl, col, k):
    t0 = time.perf_counter()
    Y_p = O.sht_matrix(cfg["N"], azi, colat, w)
    Y_q = O.sht_matrix(cfg["N"], azl, col)
    bn = O.spherical_pw(cfg["N"], k, cfg["r"], cfg["setup"])
    dn, _ = O.regularize(1 / bn, cfg["a0"], cfg["method"])
    return Y_p, Y_q, dn, time.perf_counter() - t0


def cpu_sample_text(cfg, fs):
    return ("per step: stages 1-2 for all %d bins (SciPy, 1 thread) + stage 3 on %d of %d bins (full L, T; "
            "complex128 NumPy/OpenBLAS factored form, all cores); value = F*L*T / (t_stage12 + t_stage3 * %d/%d)"
            % (cfg["F"], fs, cfg["F"], cfg["F"], fs))


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
    azi, colat, w, azl, col, k = make_geometry(O, cfg)
    fs = 64                                               # bins per step (sample)
    X = O.synth_spectra(fs, len(azi), cfg["T"], seed=0).astype(np.complex128)
    t12s, t3s, times = [], [], []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        Y_p, Y_q, dn, t12 = cpu_stage12(O, cfg, azi, colat, w, azl, col, k)
        t1 = time.perf_counter()
        O.pwd_factored(Y_q, O.repeat_n_m(dn[:fs]), Y_p, X)
        t2 = time.perf_counter()
        if it >= args.warmup:
            t12s.append(t1 - t0)
            t3s.append(t2 - t1)
            times.append(t2 - t0)
    # whole-job throughput: stages 1-2 once per batch of F bins + stage 3 scaled from the sampled bins
    t_full = sum(t12s) / len(t12s) + (sum(t3s) / len(t3s)) * cfg["F"] / fs
    val = cfg["F"] * cfg["L"] * cfg["T"] / t_full
    t = sum(times)
    sample = cpu_sample_text(cfg, fs)
    line = {"impl": "reference", "metric": "beam outputs/sec", "value": val, "unit": "outputs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / len(times), "higher_is_better": True,
            "scaling": "weak" if args.gpus == 1 else "strong",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": config_dict(cfg, args.gpus),
            "cpu_baseline": {"value": val, "unit": "outputs/s", "cores": threads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": "outputs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def cpu_baseline(O, cfg, azi, colat, w, azl, col, k):
    """Oracle port on the host cores, bounded sample (~10-20 s): the factored form (what a careful NumPy user
    would write) and the reference's LITERAL chain (diagonal_mode_mat + three np.matmul,
    doc/examples/modal_beamforming_rigid_spherical_array.py:36-39) on the same bins."""
    fs = 32
    threads = use_all_host_threads()
    X = O.synth_spectra(fs, len(azi), cfg["T"], seed=0).astype(np.complex128)
    Y_p, Y_q, dn, t_stage12 = cpu_stage12(O, cfg, azi, colat, w, azl, col, k)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        O.pwd_factored(Y_q, O.repeat_n_m(dn[:fs]), Y_p, X)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    outputs = fs * cfg["L"] * cfg["T"]
    # whole-path equivalent: stage 1/2 once per step of F bins + stage 3 scaled to F bins
    t_full = t_stage12 + best * cfg["F"] / fs
    row = {"value": cfg["F"] * cfg["L"] * cfg["T"] / t_full, "unit": "outputs/s", "cores": threads,
           "kind": "port",
           "sample": "stage 3 on %d of %d bins (full L, T; complex128 NumPy/OpenBLAS factored form, best of 3, "
                     "%.3f s) scaled to %d bins + stage 1/2 once (%.3f s, SciPy, single thread)"
                     % (fs, cfg["F"], best, cfg["F"], t_stage12),
           "stage3_outputs_per_s": outputs / best, "stage12_s": t_stage12}
    try:
        fl = 4
        D = O.diagonal_mode_mat(dn[:fl])
        lit = None
        for _ in range(2):
            t0 = time.perf_counter()
            O.pwd_literal(Y_q, D, Y_p, X[:fl])
            dt = time.perf_counter() - t0
            lit = dt if lit is None else min(lit, dt)
        t_lit = t_stage12 + lit * cfg["F"] / fl
        row["literal"] = {"value": cfg["F"] * cfg["L"] * cfg["T"] / t_lit, "unit": "outputs/s",
                          "sample": "the reference's literal chain (dense [F, M, M] diagonal stack + 3 np.matmul) on %d "
                                    "of %d bins, full L and T, best of 2 (%.3f s), scaled + stage 1/2 once"
                                    % (fl, cfg["F"], lit)}
    except Exception as e:
        row["literal"] = {"error": repr(e)}
    return row


def parity_check(O, cfg, azi, colat, w, azl, col, k, Xh_bins, got, bins):
    """Compare `bins` of the beam tensor the timed loop just wrote with the CPU oracle (complex128
    factored form).  Returns the worst normwise error per bin, max|q - ref| / max|ref|."""
    Y_p, Y_q, dn, _ = cpu_stage12(O, cfg, azi, colat, w, azl, col, k)
    ref = O.pwd_factored(Y_q, O.repeat_n_m(dn)[bins], Y_p, Xh_bins.astype(np.complex128))
    errs = [float(np.max(np.abs(got[i] - ref[i])) / np.max(np.abs(ref[i]))) for i in range(len(bins))]
    pm_ref = np.mean(np.abs(ref) ** 2, axis=-1)
    return {"bins": [int(b) for b in bins], "max_normwise_err": max(errs), "tolerance": 1e-5,
            "oracle": "oracle.sfa_oracle.pwd_factored (complex128 NumPy) on the same X, after the timed loop"}, pm_ref


# ------------------------------------------------------------------------------------------- GPU helpers
def gpu_timer(torch, dist, world, dev):
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n=3, warm=1):
        """ms per call: CUDA events around n calls, barrier + synchronize on both sides, max over ranks."""
        for _ in range(warm):
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
    return barrier, timed


def measure_tf32_peak(torch, dev):
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
    return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12


def measure_fp64_peak(torch, dev):
    a = torch.randn((4096, 4096), device=dev, dtype=torch.float64)
    b = torch.randn((4096, 4096), device=dev, dtype=torch.float64)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * 4096 ** 3 / (best * 1e-3) / 1e12


CFG5 = ("cfg5", 32, 1089, 4096, 16384, 32, 0.2, "rigid")
CFG4 = ("cfg4", 16, "gauss", 1024, 2562, 64, 0.1, "card")


def build_config_operators(micarray, O, spec):
    A, R = micarray.modal.angular, micarray.modal.radial
    name, N, grid, F, L, T, r, setup = spec
    az, co, w = O.grid_gauss(N) if grid == "gauss" else O.fibonacci_grid(grid)
    azl, col, _ = O.fibonacci_grid(L)
    Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
    _, dn, _ = R.radial_filters(N, O.wavenumbers(F) + 0.1, r, setup, 100.0, "Tikh")
    return Y_p, Y_q, dn


def tensor_bound_configs(torch, micarray, O, dev, precision, peaks):
    """Stage 3 of the two tensor-bound BASELINE configs (cfg 4: N=16, 578 mics; cfg 5: N=32, 1089 mics) at
    full size -- the in-kernel K-loop GEMM (cgemm3x_kloop_sm100.cu) -- next to a cuBLAS TF32 GEMM timed
    on the same device and to MEASURED_PEAKS' sustained bf16 rate / 2.  `frac` = algorithmic real flops *
    passes / time / peak, passes = 2 (hybrid: one tf32 pass + one bf16 pass of the same duration) or 3 (3xTF32)."""
    S = micarray.modal.steering
    tf32_peak = measure_tf32_peak(torch, dev)
    tf32_sust = 0.5 * peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    passes = 2 if precision == "c64_fast" else 3
    rows = []
    for spec in (CFG4, CFG5):
        name, N, grid, F, L, T, r, setup = spec
        Y_p, Y_q, dn = build_config_operators(micarray, O, spec)
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
                     "frac_of_tf32_peak": flops * passes / (ms * 1e-3) / 1e12 / tf32_peak,
                     "frac_of_measured_bf16_sustained_half": flops * passes / (ms * 1e-3) / 1e12 / tf32_sust})
        del X, out, plan, Y_p, Y_q, dn
        torch.cuda.empty_cache()
    return rows


def small_configs(torch, micarray, O, dev, precision, hbm_gbs):
    """Stage 3 of BASELINE configs 1 and 2 (T = 256): HBM-bound, small -- time per call through a CUDA graph
    of the plan call, with the algorithmic bytes 8 F T (Q + L) against the measured HBM peak."""
    A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
    rows = []
    for name in ("cfg1", "cfg1 (Gauss grid)", "cfg2"):
        if name.startswith("cfg1"):
            N, F, L, T = 4, 512, 360, 256
            az, co, w = O.grid_equal_angle(N) if name == "cfg1" else O.grid_gauss(N)
            azl, col = np.linspace(0, 2 * np.pi, L, endpoint=False), np.full(L, np.pi / 2)
            Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
            _, dn, _ = R.radial_filters(N, O.wavenumbers(F), 0.042, "rigid", 100.0, "softclip")
        else:
            N, F, L, T, Qc = 16, 1024, 360, 256, 32
            pol = np.linspace(0, 2 * np.pi, Qc, endpoint=False)
            Y_p = A.cht_matrix(N, pol, np.full(Qc, 1.0 / Qc))
            Y_q = A.cht_matrix(N, np.linspace(0, 2 * np.pi, L, endpoint=False))
            _, dn, _ = R.radial_filters(N, O.wavenumbers(F) + 0.1, 0.1, "open", 3000.0, "softclip", kind="circular_pw")
        Q = Y_p.shape[0]
        X = torch.randn((F, Q, T), dtype=torch.complex64, device=dev)
        plan = S.PwdPlan(Y_q, dn, Y_p, T, precision=precision)
        out = S.empty_beams(F, L, T, torch.complex64, dev)
        for _ in range(3):
            plan(X, out=out)
        torch.cuda.synchronize()
        g = None
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                plan(X, out=out)
        except Exception:
            g = None
            torch.cuda.synchronize()
        fn = (lambda: g.replay()) if g is not None else (lambda: plan(X, out=out))
        fn()
        torch.cuda.synchronize()
        n = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        nbytes = 8.0 * F * T * (Q + L)
        rows.append({"config": "%s: N=%d, Q=%d, F=%d, L=%d, T=%d" % (name, N, Q, F, L, T), "form_used": int(plan.shape.form),
                     "graph": g is not None, "ms": ms, "outputs_per_s": F * L * T / (ms * 1e-3),
                     "alg_GBps": nbytes / (ms * 1e-3) / 1e9, "frac_hbm": nbytes / (ms * 1e-3) / 1e9 / hbm_gbs,
                     "note": "0.5 GB (cfg 1) / 0.8 GB (cfg 2) of algorithmic traffic per call in 0.15-0.25 ms"})
        del g, X, out, plan
    return rows


def c128_rows(torch, micarray, O, dev, cfg, geometry):
    """The complex128 path (FP64 on the CUDA cores: the only way to reference-class beams) on cfg 1 and on a
    16-bin slab of cfg 3, against a cuBLAS DGEMM timed on the same device."""
    A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
    peak = measure_fp64_peak(torch, dev)
    rows = []
    specs = []
    N, F, L, T = 4, 512, 360, 256
    az, co, w = O.grid_equal_angle(N)
    azl, col = np.linspace(0, 2 * np.pi, L, endpoint=False), np.full(L, np.pi / 2)
    specs.append(("cfg1", N, az, co, w, azl, col, O.wavenumbers(F), 0.042, "softclip", T))
    azi, colat, w3, azl3, col3, k3 = geometry
    specs.append(("cfg3 slab of 16 bins", cfg["N"], azi, colat, w3, azl3, col3, k3[100:116], cfg["r"], cfg["method"], cfg["T"]))
    for name, N, az, co, w, azl, col, k, r, method, T in specs:
        Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
        _, dn, _ = R.radial_filters(N, k, r, "rigid", 100.0, method)
        F, Q, L, M = len(k), Y_p.shape[0], Y_q.shape[0], Y_p.shape[1]
        X = torch.randn((F, Q, T), dtype=torch.complex128, device=dev)
        plan = S.PwdPlan(Y_q, dn, Y_p, T, precision="c128")
        out = S.empty_beams(F, L, T, torch.complex128, dev)
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
        form = int(plan.shape.form)
        flops = 8.0 * F * L * Q * T if form == 1 else 8.0 * F * T * M * (Q + L)
        rows.append({"config": "%s: N=%d, Q=%d, F=%d, L=%d, T=%d, complex128" % (name, N, Q, F, L, T), "form_used": form,
                     "ms": ms, "outputs_per_s": F * L * T / (ms * 1e-3), "alg_fp64_tflops": flops / (ms * 1e-3) / 1e12,
                     "cublas_dgemm_tflops": peak, "frac_of_dgemm": flops / (ms * 1e-3) / 1e12 / peak,
                     "alg_GBps": 16.0 * F * T * (Q + L) / (ms * 1e-3) / 1e9})
        del X, out, plan
    return rows


def fft_rows(torch, micarray, dev, out, hbm_gbs):
    """Time/frequency ends: irfft + fftshift along the bin axis of a column slab of the beam
    tensor (n = 2048, the headline STFT length) and the STFT front end producing X[F, Q, T]."""
    from sfa_numpy_b200 import fft as sfft
    rows = []
    F, L, T = out.shape
    lc = min(L, 256)
    q = out[:, :lc, :].contiguous()
    n = 2 * F
    C = lc * T

    def t(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms = t(lambda: sfft.irfft(q, n=n, shift=True))
    nbytes = 8.0 * F * C + 4.0 * n * C
    rows.append({"stage": "irfft+fftshift along bins", "config": "F=%d -> n=%d, %d columns (look dirs x frames), complex64" % (F, n, C),
                 "ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "frac_hbm": nbytes / (ms * 1e-3) / 1e9 / hbm_gbs,
                 "outputs_per_s": F * C / (ms * 1e-3)})
    S_len = 2048 + 512 * (T - 1)
    x = torch.randn((64, S_len), device=dev)
    win = torch.hann_window(2048, device=dev)
    ms = t(lambda: sfft.stft(x, 2048, 512, window=win, bins=F))
    nbytes = 4.0 * 64 * S_len + 8.0 * F * 64 * T
    rows.append({"stage": "stft", "config": "64 mics x %d samples (10 s at 48 kHz), nfft 2048, hop 512 -> X[%d, 64, %d]" % (S_len, F, T),
                 "ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "frac_hbm": nbytes / (ms * 1e-3) / 1e9 / hbm_gbs})
    return rows


def pcie_probe(torch, dev, host_arrays):
    """Raw cudaMemcpyAsync rates on the SAME pinned buffers the e2e call uses (torch views of them), 1 GiB pieces."""
    res = {}
    try:
        Xp, outp = host_arrays
        n = min(outp.size, 1 << 27)                      # 1 GiB of complex64
        h = torch.from_numpy(outp.reshape(-1)[:n])
        d = torch.empty(n, dtype=torch.complex64, device=dev)
        for name, fn in (("d2h_GBps", lambda: h.copy_(d, non_blocking=True)), ("h2d_GBps", lambda: d.copy_(h, non_blocking=True))):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            res[name] = 3 * n * 8 / (time.perf_counter() - t0) / 1e9
        res["pinned_seen_by_torch"] = bool(h.is_pinned())
    except Exception as e:
        res["error"] = repr(e)
    return res


# ------------------------------------------------------------------------------------------- GPU arm
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
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--pieces", type=int, default=4, help="bin blocks per rank of the overlapped complex all-gather")
    ap.add_argument("--precision", default="c64_fast", choices=["c64", "c64_fast"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from oracle import sfa_oracle as O            # synthetic-input generator + cpu_baseline + parity check only
    import sfa_numpy_b200 as micarray          # loads the prebuilt libsfa_b200.so; raises if it is missing
    from sfa_numpy_b200 import _lib
    from sfa_numpy_b200 import dist as sdist
    from sfa_numpy_b200.pipeline import ModalBeamformer
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
    barrier, timed = gpu_timer(torch, dist, world, dev)
    peaks, peak_src = load_peaks()

    geometry = make_geometry(O, cfg)
    azi, colat, w, azl, col, k = geometry
    Q, F, L, T, N = cfg["Q"], cfg["F"], cfg["L"], cfg["T"], cfg["N"]
    f0, f1 = sdist.shard_bins(F, world, rank)
    nb = f1 - f0
    # ONE problem: the spectra of all F bins come from one seed; every rank keeps its slab
    Xh = O.synth_spectra(F, Q, T, seed=0)[f0:f1]
    bf = ModalBeamformer(N, (azi, colat, w), (azl, col), k, cfg["r"], cfg["setup"], cfg["a0"], cfg["method"], T,
                         precision=args.precision, bins=(f0, f1), power=(world > 1), graph=not args.no_graph, device=dev)
    bf.X.copy_(torch.from_numpy(Xh))
    pm_full = torch.empty((F, L), dtype=torch.float32, device=dev) if world > 1 else None
    even = (F % world == 0)

    def gather_maps():
        if even:
            dist.all_gather_into_tensor(pm_full, bf.power)
        else:
            pm_full.copy_(sdist.all_gather_slabs(bf.power, F))

    def step():
        bf.run()                                   # stages 1 + 2 + 3 of this rank's slab: one graph launch
        if world > 1:
            gather_maps()                          # the collective of the path: beam power maps of all bins on every rank

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    R.check_deferred()
    launches = bf.launches_per_run * args.steps
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    step_ms = ms / args.steps
    value = F * L * T / (step_ms * 1e-3)
    # ---- dominant kernel: average launch duration with CUDA events on its launching stream (eager steps:
    #      the library brackets the kernel with events when sfa_prof_enable(1); graph replays carry no events) ----
    n_prof = max(2, min(args.steps, 20))      # right after the timed loop: same sustained power / clock state
    bf.run_eager()
    torch.cuda.synchronize()
    lib.sfa_prof_enable(1)
    for _ in range(n_prof):
        bf.run_eager()
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_longlong()
    lib.sfa_prof_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.sfa_prof_enable(0)
    gemm_ms_per_step = tot.value / n_prof
    eager_ms = timed(lambda: bf.run_eager(), n=5)
    # the output of the LAST timed step against the oracle: a fast step that computes something else is not a result
    check_local = sorted(set([0, min(74, nb - 1), nb // 2 + 5 if nb // 2 + 5 < nb else nb - 1, nb - 1]))
    got = bf.out[check_local].cpu().numpy()
    pcheck, pm_ref = parity_check(O, cfg, azi, colat, w, azl, col, k, Xh[check_local], got, [f0 + b for b in check_local])
    if world > 1:
        pm_got = pm_full[[f0 + b for b in check_local]].cpu().numpy()
        pcheck["power_map_max_rel_err"] = float(np.max(np.abs(pm_got - pm_ref) / np.max(pm_ref, axis=1, keepdims=True)))
        # every rank checks its own bins of the gathered map; the worst case is reported
        tt = torch.tensor([pcheck["max_normwise_err"], pcheck["power_map_max_rel_err"]], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        pcheck["max_normwise_err"], pcheck["power_map_max_rel_err"] = float(tt[0].item()), float(tt[1].item())
        # ... and one foreign bin of the gathered map against this rank's oracle would need that rank's X: the
        # gathered map is instead compared across ranks (identical on all of them)
        chk = pm_full.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        pcheck["gathered_map_identical_on_all_ranks"] = bool(float(lo.item()) == float(hi.item()))
    bad = not (pcheck["max_normwise_err"] <= pcheck["tolerance"]) or pcheck.get("power_map_max_rel_err", 0.0) > 1e-4 \
        or pcheck.get("gathered_map_identical_on_all_ranks", True) is False
    if bad:
        raise SystemExit("bench: parity check failed: %r" % (pcheck,))

    # algorithmic bytes of stage 3 in complex64 I/O = spectra read + beams written + the two SH
    # matrices + the radial filters.  The steering matrices A_f are this implementation's own intermediate
    # and do NOT count.  Algorithmic flops = 8 F L Q T (steering form, Q < M).  Per rank: its nb bins.
    M_modes, C_d = (N + 1) ** 2, N + 1
    alg_bytes = 8.0 * nb * T * (Q + L) + 8.0 * (Q + L) * M_modes + 8.0 * nb * C_d
    alg_flops = 8.0 * nb * L * Q * T
    passes = 2 if args.precision == "c64_fast" else 3
    # tensor-side denominator: measured dense bf16 throughput / 2 = TF32 rate (sustained: the kernel runs
    # inside a long step under the power cap); a tf32 pass and the bf16 pass of the hybrid mode take the same time
    tf32_peak = 0.5 * peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    t_hbm_ms = alg_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3
    t_tensor_ms = alg_flops * passes / (tf32_peak * 1e12) * 1e3
    achieved = alg_bytes / (gemm_ms_per_step * 1e-3) / 1e9
    launches_per_step = cnt.value / n_prof
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the fused kernel over 148 bins from the committed
    # ncu capture (profiles/r5_c_fused_final.txt), scaled to this rank's bins; bench.py never runs under ncu.
    NCU_TRAFFIC_148_BINS = 2.846004e9 + 0.086852e9
    roofline = {"kernel": "pwd_fused_kernel: stage-3 steering build + GEMM in one persistent launch, %s (tcgen05)"
                          % ("hybrid TF32+BF16" if args.precision == "c64_fast" else "3xTF32"),
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
                "traffic": NCU_TRAFFIC_148_BINS * nb / 148.0 / max(launches_per_step, 1.0) if args.precision == "c64_fast" else None,
                "traffic_source": "ncu --set full, one 148-bin launch (profiles/r5_c_fused_final.txt), scaled to the launch's bins",
                "alg_bytes_per_launch": alg_bytes / launches_per_step if launches_per_step else None,
                "kernel_ms_per_launch": gemm_ms_per_step / launches_per_step if launches_per_step else None,
                "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback",
                "launches_per_step": launches_per_step,
                "kernel_ms_per_step": gemm_ms_per_step, "kernel_share_of_step": gemm_ms_per_step / step_ms,
                "eager_ms_per_step": eager_ms,
                "alg_bytes_per_step": alg_bytes, "alg_tflops_achieved": alg_flops / (gemm_ms_per_step * 1e-3) / 1e12}

    line = {"metric": "beam outputs/sec", "value": value, "unit": "outputs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
            "dtype": DTYPE_NOTE[args.precision], "data": "synthetic",
            "config": config_dict(cfg, world),
            "step_api": "sfa_numpy_b200.pipeline.ModalBeamformer.run(): stages 1-3 as ONE CUDA-graph launch (%s)%s"
                        % ("graph" if bf.graph is not None else "eager: " + str(bf.graph_error),
                           "" if world == 1 else " + torch.distributed.all_gather_into_tensor of the [F/N, L] power maps"),
            "bins_per_gpu": nb,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "parity_check": pcheck}

    wd = start_watchdog(900, line, rank)      # everything below is explanation; the headline above is safe
    # ---- multi-GPU: the three readings of "sharded over the GPUs with an all-gather of the beam maps" ----
    if world > 1:
        try:
            line["scaling_variants"] = scaling_variants(torch, dist, sdist, S, bf, timed, world, rank, dev, args, cfg,
                                                        step_ms, gather_maps, Xh, f0)
        except Exception as e:
            line["scaling_variants"] = {"error": repr(e)}
            barrier()
    else:
        pm_ms = timed(lambda: S.power_map(bf.out))
        line["collective"] = {"power_map_kernel_ms": pm_ms, "power_map_read_GBps": 8.0 * nb * L * T / (pm_ms * 1e-3) / 1e9,
                              "note": "single GPU: no collective; at N > 1 the map comes out of the fused kernel's epilogue"}
        try:
            # the beam-map product alone: stage 3 with the power map accumulated in the epilogue and NO beam stores
            # (the beams exist in tensor memory only) -- the tensor-issue side of the kernel without its HBM side
            pm_t = torch.empty((nb, L), dtype=torch.float32, device=dev)
            mo_ms = timed(lambda: bf.plan(bf.X, power=pm_t, beams=False), n=5)
            line["power_map_only"] = {"ms": mo_ms, "outputs_per_s": F * L * T / (mo_ms * 1e-3),
                                      "api": "PwdPlan(X, power=..., beams=False): stage 3 only, beams never written",
                                      "alg_tflops": 8.0 * nb * L * Q * T / (mo_ms * 1e-3) / 1e12}
        except Exception as e:
            line["power_map_only"] = {"error": repr(e)}
            torch.cuda.synchronize()

    # ---- e2e through the host-buffer C ABI: every rank pushes ITS SLAB of the problem through its own PCIe link.
    #      The timed region covers stages 1-2 on the device, the hand-over of the operators to the host context
    #      (sfa_pwd_host_set_operators_dev) and sfa_pwd_host_run (H2D of X, kernels, D2H of the beams). ----
    if not args.no_e2e:
        host, err, fft_out = None, None, None
        prev_cpus = bind_to_gpu_numa_node(local) if world > 1 else None
        try:
            torch.cuda.synchronize()
            Yq_d, Yp_d, dn_d = bf.Y_look, bf.Y_mic, bf.dn[f0:f1].contiguous()
            Yq_h, Yp_h, dn_h = Yq_d.cpu().numpy(), Yp_d.cpu().numpy(), dn_d.cpu().numpy()
            if rank == 0 and not args.no_other_configs:
                try:
                    fft_out = fft_rows(torch, micarray, dev, bf.out, peaks["hbm_gbs"])
                except Exception as e:
                    fft_out = {"error": repr(e)}
                    torch.cuda.synchronize()
            d_geo = [torch.from_numpy(a).to(dev) for a in (azi, colat, w, azl, col, k)]
            del bf
            torch.cuda.empty_cache()
            host = S.PwdHost(Yq_h, dn_h, Yp_h, T, precision=args.precision, device=local, chunk_bins=0)
            Xp = host.pinned_empty((nb, Q, T))
            Xp[...] = Xh
            outp = host.pinned_empty((nb, L, T))

            def e2e_step():
                Y_p = A.sht_matrix(N, d_geo[0], d_geo[1], d_geo[2])
                Y_q = A.sht_matrix(N, d_geo[3], d_geo[4])
                _, dn, _ = R.radial_filters(N, d_geo[5], cfg["r"], cfg["setup"], cfg["a0"], cfg["method"], check="defer")
                host.set_operators(Y_q, dn[f0:f1].contiguous(), Y_p)
                host(Xp, out=outp)

            e2e_step()                                   # warm-up
        except Exception as e:                           # e.g. not enough pinned host memory
            err = repr(e)
        ok = torch.tensor([0 if err else 1], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)    # all ranks take the same branch (no barrier mismatch)
        if int(ok.item()) == 1:
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e_step()
            dt = (time.perf_counter() - t0) / args.e2e_steps
            td = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dt = float(td.item())
            tb = torch.tensor([float(Xp.nbytes), float(outp.nbytes)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tb)
            R.check_deferred()
            e2e = {"value": F * L * T / dt, "unit": "outputs/s",
                   "h2d_bytes_per_step": int(tb[0].item()), "d2h_bytes_per_step": int(tb[1].item()),
                   "ms_per_step": dt * 1e3,
                   "api": "stage 1-2 on the device (angular.sht_matrix x2, radial.radial_filters) + PwdHost.set_operators + "
                          "sfa_pwd_host_run (C ABI: pinned host spectra in, host beams out, chunked H2D/compute/D2H "
                          "pipeline); one call per rank on its slab of the bins, wall clock, max over ranks",
                   "steps": args.e2e_steps}
            # the ceiling of that number: raw copies on the same buffers
            probe = pcie_probe(torch, dev, (Xp, outp))
            if world > 1:
                tp = torch.tensor([probe.get("d2h_GBps", 0.0), probe.get("h2d_GBps", 0.0)], device=dev, dtype=torch.float64)
                dist.all_reduce(tp)
                probe = {"d2h_GBps_sum_over_ranks": float(tp[0].item()), "h2d_GBps_sum_over_ranks": float(tp[1].item()),
                         "note": "every rank copies at the same time (the probe runs on all ranks between two barriers)"}
                d2h_probe = probe["d2h_GBps_sum_over_ranks"]
            else:
                d2h_probe = probe.get("d2h_GBps", 0.0)
            e2e["pcie_probe"] = probe
            e2e["d2h_GBps_achieved"] = float(tb[1].item()) / dt / 1e9
            e2e["frac_of_pcie"] = (e2e["d2h_GBps_achieved"] / d2h_probe) if d2h_probe else None
            e2e["limiter"] = "PCIe D2H of the complex beams (%.1f GB per step)" % (float(tb[1].item()) / 1e9)
            # the "beam map" reading of the metric: only the power maps [F, L] f32 come back (10 MB)
            try:
                pmh = np.empty((nb, L), dtype=np.float32)
                host.power_map(Xp, power=pmh)
                barrier()
                t0 = time.perf_counter()
                for _ in range(max(args.e2e_steps, 3)):
                    host.power_map(Xp, power=pmh)
                dtp = (time.perf_counter() - t0) / max(args.e2e_steps, 3)
                tdp = torch.tensor([dtp], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(tdp, op=dist.ReduceOp.MAX)
                dtp = float(tdp.item())
                e2e["e2e_power_map"] = {"value": F * L * T / dtp, "unit": "outputs/s", "ms_per_step": dtp * 1e3,
                                        "h2d_bytes_per_step": int(tb[0].item()), "d2h_bytes_per_step": 4 * F * L,
                                        "api": "sfa_pwd_host_run_power: spectra in, beam power maps out; map-only kernel runs (the beams are never written), H2D-bound"}
            except Exception as e:
                e2e["e2e_power_map"] = {"error": repr(e)}
            line["e2e"] = e2e
        else:
            line["e2e"] = {"value": None, "unit": "outputs/s", "error": err or "another rank failed to set up"}
        if host is not None:
            host.close()
        if prev_cpus:
            os.sched_setaffinity(0, prev_cpus)           # the CPU baseline below uses every core again
        if rank == 0 and fft_out is not None:
            line["fft"] = fft_out
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
    if rank == 0 and not args.no_stage12:
        # stages 1-2 of the headline config as they run inside the step: one graph launch over the four calls
        # (2 x sht_matrix, radial filter + zero replacement) instead of four launch-latency-bound API calls
        try:
            d_geo12 = [torch.from_numpy(a).to(dev) for a in (azi, colat, w, azl, col, k)]

            def s12():
                A.sht_matrix(N, d_geo12[0], d_geo12[1], d_geo12[2])
                A.sht_matrix(N, d_geo12[3], d_geo12[4])
                R.radial_filters(N, d_geo12[5], cfg["r"], cfg["setup"], cfg["a0"], cfg["method"], check="none")
            for _ in range(3):
                s12()
            torch.cuda.synchronize()
            def local_ms(fn, n):                     # rank-0-only section: no barriers in here
                fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    fn()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / n
            eager_us = local_ms(s12, 20) * 1e3
            g12 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g12):
                s12()
            graph_us = local_ms(lambda: g12.replay(), 50) * 1e3
            evals = (Q + L) * (N + 1) ** 2 + F * (N + 1)
            line["stage12_graph"] = {"config": "cfg3: sht_matrix at 64 mics + 2562 look dirs, spherical_pw + regularize at 1024 bins",
                                     "eager_us": eager_us, "graph_us": graph_us, "evals": evals,
                                     "evals_per_s_graph": evals / (graph_us * 1e-6),
                                     "note": "3.5 MB of output: latency-bound whatever is done; inside the step it is 1-2 % of the time"}
            del g12
        except Exception as e:
            line["stage12_graph"] = {"error": repr(e)}
            torch.cuda.synchronize()
    if rank == 0 and not args.no_other_configs and world == 1:
        for key, fn in (("tensor_bound_configs", lambda: tensor_bound_configs(torch, micarray, O, dev, args.precision, peaks)),
                        ("small_configs", lambda: small_configs(torch, micarray, O, dev, args.precision, peaks["hbm_gbs"])),
                        ("c128", lambda: c128_rows(torch, micarray, O, dev, cfg, geometry))):
            try:
                torch.cuda.empty_cache()
                line[key] = fn()
            except Exception as e:
                line[key] = {"error": repr(e)}
                torch.cuda.synchronize()
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()                 # rank 0 may still be timing its CPU / single-GPU extras
        wd.cancel()
        dist.destroy_process_group()
    wd.cancel()


def scaling_variants(torch, dist, sdist, S, bf, timed, world, rank, dev, args, cfg, step_ms, gather_maps, Xh, f0):
    """cfg 3 and cfg 5, each as ONE problem sharded over the ranks, three ways (ms per step, max over ranks):
      sharded_resident        -- every rank computes its slab, nothing is exchanged;
      allgather_power_map     -- + NCCL all-gather of the [F/N, L] float32 beam power maps (cfg 3: this is `value`);
      allgather_complex_beams -- the complex beam tensor [F, L, T] gathered on every rank, the bins dealt out
                                 block-cyclically so that the gather of block c overlaps the compute of block c + 1
                                 (`dist.OverlappedPwd`).  Every GPU ingests (N-1)/N of the output over NVLink.
                                 Two transports: NCCL all-gather per block on a side stream, and (`_p2p`) DMA pushes
                                 of the finished block into every peer's buffer (CUDA IPC, copy engines, no SMs)."""
    import sfa_numpy_b200 as micarray
    from oracle import sfa_oracle as O
    F, L, T, Q = cfg["F"], cfg["L"], cfg["T"], cfg["Q"]
    res = {"limiter_of_the_complex_gather": "ncclAllGather NVLink ingest per GPU: (N-1)/N of the beam tensor must enter every GPU"}
    outputs = float(F) * L * T
    row = {"outputs_per_step": outputs}
    ms_i = timed(lambda: bf.run(), n=5)
    row["sharded_resident"] = {"ms": ms_i, "outputs_per_s": outputs / (ms_i * 1e-3)}
    row["allgather_power_map"] = {"ms": step_ms, "outputs_per_s": outputs / (step_ms * 1e-3),
                                  "bytes_gathered_per_gpu": 4 * F * L * (world - 1) // world}
    pieces = args.pieces
    while F % (world * pieces) and pieces > 1:
        pieces -= 1
    if F % (world * pieces) == 0:
        Xfull = O.synth_spectra(F, Q, T, seed=0)
        nbk = bf.out.shape[0]
        for transport, key in (("nccl", "allgather_complex_beams"), ("p2p", "allgather_complex_beams_p2p")):
            try:
                op = sdist.OverlappedPwd(bf.Y_look, bf.dn, bf.Y_mic, T, pieces=pieces, precision=args.precision, device=dev,
                                         transport=transport)
                X_bc = torch.from_numpy(np.ascontiguousarray(Xfull[op.my_bins()])).to(dev)

                def step_ii():
                    bf._stage12(check="none")              # stages 1-2 as in the other variants (eager launches)
                    op(X_bc)
                ms_ii = timed(step_ii, n=3)
                # result check: this rank's slab of the gathered tensor against its own resident result
                same = bool(torch.equal(op.full[f0:f0 + nbk], bf.out))
                ingest = 8.0 * F * L * op.buf.shape[2] * (world - 1) / world
                row[key] = {"ms": ms_ii, "outputs_per_s": outputs / (ms_ii * 1e-3), "pieces": pieces, "transport": transport,
                            "ingest_bytes_per_gpu": ingest, "ingest_GBps_per_gpu": ingest / (ms_ii * 1e-3) / 1e9,
                            "matches_resident_slab": same}
                del op, X_bc
            except Exception as e:
                row[key] = {"error": repr(e)}
                torch.cuda.synchronize()
            torch.cuda.empty_cache()
        del Xfull
    res["cfg3"] = row
    # ---- cfg 5 (the config BASELINE.json names for 1/2/4/8 GPUs): tensor-bound, K-loop kernel ----
    try:
        name, N5, grid, F5, L5, T5, r5, setup5 = CFG5
        Y_p, Y_q, dn = build_config_operators(micarray, O, CFG5)
        Q5 = Y_p.shape[0]
        g0, g1 = sdist.shard_bins(F5, world, rank)
        gen = torch.Generator(device=dev).manual_seed(5)
        X5 = torch.randn((F5, Q5, T5), dtype=torch.complex64, device=dev, generator=gen)     # same on every rank
        plan = S.PwdPlan(Y_q, dn[g0:g1], Y_p, T5, precision=args.precision)
        out5 = S.empty_beams(g1 - g0, L5, T5, torch.complex64, dev)
        Xs = X5[g0:g1].contiguous()
        outputs5 = float(F5) * L5 * T5
        row5 = {"outputs_per_step": outputs5, "config": "cfg5: N=32, 1089 mics, F=4096, L=16384, T=32 (stage 3, factored form)"}
        ms_a = timed(lambda: plan(Xs, out=out5), n=3)
        row5["sharded_resident"] = {"ms": ms_a, "outputs_per_s": outputs5 / (ms_a * 1e-3)}
        pm5 = torch.empty((F5, L5), dtype=torch.float32, device=dev)

        def step_pm():
            plan(Xs, out=out5)
            loc = S.power_map(out5)
            if F5 % world == 0:
                dist.all_gather_into_tensor(pm5, loc)
            else:
                pm5.copy_(sdist.all_gather_slabs(loc, F5))
        ms_b = timed(step_pm, n=3)
        row5["allgather_power_map"] = {"ms": ms_b, "outputs_per_s": outputs5 / (ms_b * 1e-3),
                                       "bytes_gathered_per_gpu": 4 * F5 * L5 * (world - 1) // world}
        del plan, out5, pm5
        torch.cuda.empty_cache()
        pieces5 = 8
        while F5 % (world * pieces5) and pieces5 > 1:
            pieces5 -= 1
        for transport, key in (("nccl", "allgather_complex_beams"), ("p2p", "allgather_complex_beams_p2p")):
            try:
                op5 = sdist.OverlappedPwd(Y_q, dn, Y_p, T5, pieces=pieces5, precision=args.precision, device=dev,
                                          transport=transport)
                X5bc = X5[op5.my_bins()].contiguous()
                ms_c = timed(lambda: op5(X5bc), n=3)
                ingest5 = 8.0 * F5 * L5 * op5.buf.shape[2] * (world - 1) / world
                row5[key] = {"ms": ms_c, "outputs_per_s": outputs5 / (ms_c * 1e-3), "pieces": pieces5, "transport": transport,
                             "ingest_bytes_per_gpu": ingest5, "ingest_GBps_per_gpu": ingest5 / (ms_c * 1e-3) / 1e9}
                del op5, X5bc
            except Exception as e:
                row5[key] = {"error": repr(e)}
                torch.cuda.synchronize()
            torch.cuda.empty_cache()
        res["cfg5"] = row5
        del X5
        torch.cuda.empty_cache()
    except Exception as e:
        res["cfg5"] = {"error": repr(e)}
        torch.cuda.synchronize()
    return res


if __name__ == "__main__":
    main()
