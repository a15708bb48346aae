#!/usr/bin/env python
"""Benchmark of the deComP hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--legs a,b,...]

Primary workload (BASELINE.json configs[1]): batched FISTA Lasso, 100 000 problems per GPU, A (256, 1024),
alpha = 0.1, float64, tol = 0 (fixed iteration count).  One *step* is one FISTA iteration over the whole batch.
After W warm-up iterations (plus one untimed region, so that the launch pattern being timed has run before), the
region "K iterations" is timed ``--repeats`` times back to back on the same solve (CUDA events on the launching
stream between barriers, max over ranks) and the MEDIAN region is reported: ``value`` = problems x K / median.
``gpu_launches`` counts the kernel launches of one region, the roofline is per launch of the dominant kernel.

``e2e``: the same metric through ``decomp_b200.lasso.solve`` with ORDINARY (pageable) numpy arrays in and a numpy
array out -- H2D of y and A, set-up GEMMs, K iterations, D2H of x inside the timed region (median of 5 calls, max
over ranks); ``e2e_pinned`` is the same call on page-locked arrays.

Secondary (configs[2]): NMF multiplicative update, 4096 features, k = 256, float64; one step = one full sweep.
``secondary`` = 1 000 000 rows per GPU (weak scaling); ``secondary_strong`` = 1 000 000 rows in total, sharded
(what the ">= 7x at 8 GPUs" target is quoted on), both with the time spent in the all-reduces per sweep, a CPU
baseline at n = 65 536 and an end-to-end ``nmf.solve`` call on host arrays at n = 131 072.
``extra_configs``: configs[3] (dictionary-learning minibatch step, complex128, 10 % mask) and configs[4] (masked
NMF sweep and masked FISTA iteration, 1 000 000 rows per GPU x 1024, k = 128).
``parity_multi_gpu``: small sharded NMF / Lasso / dictionary-learning solves checked against the numpy oracle on
rank 0, outside every timed region.

``--impl reference`` times the reference's own numpy implementation on the host cores with all BLAS threads: the
unmodified reference from baseline/_ref when it is installed there (kind "reference"), else the pinned numpy port
oracle/decomp_oracle.py (kind "port").  Under torchrun only rank 0 works.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FISTA = dict(batch=100000, k=256, f=1024, alpha=0.1)
NMF = dict(n=1000000, f=4096, k=256)
C5 = dict(n=1000000, f=1024, k=128)
C4 = dict(n=32768, f=2048, k=512, minibatch=8192)
CPU_FISTA_BATCH = 24576         # bounded CPU sample of the cpu_baseline leg (problems)
CPU_NMF_ROWS = 65536
E2E_NMF_ROWS = 131072
CPU_REFERENCE_BUDGET = 1.4e7    # problem-iterations of the --impl reference run (~60 s on 16 cores)
L2_BYTES = 126 * 2 ** 20
ALL_LEGS = ('fista', 'tf32', 'e2e', 'nmf', 'nmf_strong', 'nmf_e2e', 'configs', 'parity', 'cpu')


# ------------------------------------------------------------------------------------------ helpers
class ClockSampler(object):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
            'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
            'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
            'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(s)}


def merge_clocks(a, b):
    if a is None:
        return b
    if b is None:
        return a
    out = dict(a)
    if b.get('sm_mhz') is not None and (a.get('sm_mhz') is None or b['sm_mhz'] < a['sm_mhz']):
        out['sm_mhz'] = b['sm_mhz']
    out['reasons'] = sorted(set(a.get('reasons', [])) | set(b.get('reasons', [])))
    out['samples'] = a.get('samples', 0) + b.get('samples', 0)
    return out


def median(v):
    s = sorted(v)
    return s[len(s) // 2] if len(s) % 2 else 0.5 * (s[len(s) // 2 - 1] + s[len(s) // 2])


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def tf32_peak():
    """Dense TF32 tensor rate (TFLOP/s) for a kernel timed inside a long step: half of the measured sustained bf16
    rate (kind::tf32 runs at half the kind::f16 rate; MEASURED_PEAKS.json has no TF32 entry), else the nominal 1100."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        if p.get('bf16_tflops_sustained'):
            return 0.5 * float(p['bf16_tflops_sustained']), ('half of MEASURED_PEAKS.json bf16_tflops_sustained (%.1f); '
                                                             'no TF32 entry there' % float(p['bf16_tflops_sustained']))
    return 1100.0, 'nominal dense TF32 (B200_PROFILING.md)'


def load_traffic():
    """dram bytes per launch of the dominant kernels, from the committed ncu --set full captures (profiles/)."""
    path = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


def fista_config(batch, world):
    """The config both arms print (the CPU arm's bounded sample is described in its cpu_baseline.sample)."""
    return {'workload': 'batched FISTA Lasso: %d problems per GPU, A (256,1024), alpha=0.1, float64, tol=0 '
                        '(BASELINE.json configs[1])' % batch,
            'l2': 'working set per iteration 5*B*k*8 = %.0f MB > %d MB L2 (inputs larger than L2, no flush)'
                  % (5.0 * batch * 256 * 8 / 1e6, L2_BYTES // 2 ** 20),
            'parallelism': 'sample axis sharded over %d GPU(s), no data-path collective' % world}


# ------------------------------------------------------------------------------------------ synthetic data
def fista_data_device(torch, B, k, f, seed, device):
    """A = randn(k, f); x_true = randn * rint(U) (~50 % sparse); y = x_true A + 0.1 randn."""
    g = torch.Generator(device=device)
    g.manual_seed(0)                          # A is the same on every rank
    A = torch.randn((k, f), dtype=torch.float64, device=device, generator=g)
    g.manual_seed(1000 + seed)
    y = torch.empty((B, f), dtype=torch.float64, device=device)
    chunk = 16384
    for r0 in range(0, B, chunk):
        r1 = min(B, r0 + chunk)
        xt = torch.randn((r1 - r0, k), dtype=torch.float64, device=device, generator=g)
        xt *= torch.round(torch.rand((r1 - r0, k), dtype=torch.float64, device=device, generator=g))
        y[r0:r1] = xt @ A
        y[r0:r1] += 0.1 * torch.randn((r1 - r0, f), dtype=torch.float64, device=device, generator=g)
    return y, A


def fista_data_host(np, B, k, f, seed):
    rng = np.random.RandomState(seed)
    A = np.random.RandomState(0).randn(k, f)
    y = np.empty((B, f))
    chunk = 16384
    for r0 in range(0, B, chunk):
        r1 = min(B, r0 + chunk)
        xt = rng.randn(r1 - r0, k) * np.rint(rng.uniform(size=(r1 - r0, k)))
        y[r0:r1] = xt.dot(A) + 0.1 * rng.randn(r1 - r0, f)
    return y, A


def nmf_data_device(torch, n, f, k, seed, device, masked=False):
    """Ct, Dt = max(N(0,1), 0); Y = Ct Dt + 0.1 N(0,1); D0 = max(Dt + 0.3 N, 0.1) replicated;
    mask = (U > 0.1) in y's dtype."""
    g = torch.Generator(device=device)
    g.manual_seed(0)
    Dt = torch.randn((k, f), dtype=torch.float64, device=device, generator=g).clamp_(min=0.0)
    D0 = (Dt + 0.3 * torch.randn((k, f), dtype=torch.float64, device=device, generator=g)).clamp_(min=0.1)
    g.manual_seed(2000 + seed)
    y = torch.empty((n, f), dtype=torch.float64, device=device)
    mask = torch.empty((n, f), dtype=torch.float64, device=device) if masked else None
    chunk = 32768
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        ct = torch.randn((r1 - r0, k), dtype=torch.float64, device=device, generator=g).clamp_(min=0.0)
        y[r0:r1] = ct @ Dt
        y[r0:r1] += 0.1 * torch.randn((r1 - r0, f), dtype=torch.float64, device=device, generator=g)
        if masked:
            mask[r0:r1] = (torch.rand((r1 - r0, f), dtype=torch.float64, device=device, generator=g) > 0.1).double()
    return y, D0, mask


def nmf_data_host(np, n, f, k, seed):
    rng = np.random.RandomState(seed)
    Dt = np.maximum(rng.randn(k, f), 0.0)
    D0 = np.maximum(Dt + 0.3 * rng.randn(k, f), 0.1)
    y = np.empty((n, f))
    chunk = 16384
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        y[r0:r1] = np.maximum(rng.randn(r1 - r0, k), 0.0).dot(Dt) + 0.1 * rng.randn(r1 - r0, f)
    return y, D0


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_fista(np, arm, steps, warmup, batch):
    """decomp.lasso.solve(method='fista', tol=0) on the host cores; returns (problem-iterations/s, seconds)."""
    k, f, alpha = FISTA['k'], FISTA['f'], FISTA['alpha']
    y, A = fista_data_host(np, batch, k, f, 0)
    if warmup > 0:
        arm.lasso(y[:1024], A, alpha, tol=0.0, method='fista', maxiter=max(1, min(warmup, 5)))
    t0 = time.perf_counter()
    arm.lasso(y, A, alpha, tol=0.0, method='fista', maxiter=steps)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt


def cpu_nmf(np, arm, sweeps, rows):
    """decomp.nmf.solve(method='mu', 'l2', tol=0) on the host cores; returns (row-iterations/s, seconds)."""
    y, D0 = nmf_data_host(np, rows, NMF['f'], NMF['k'], 0)
    t0 = time.perf_counter()
    arm.nmf(y, D0, tol=0.0, maxiter=sweeps + 1)
    dt = time.perf_counter() - t0
    return rows * sweeps / dt, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import numpy as np
    from oracle import cpu_arm
    cores = cpu_arm.set_blas_threads()
    arm = cpu_arm.load()
    steps = max(1, args.steps)
    batch = int(min(args.batch, max(4096, CPU_REFERENCE_BUDGET // steps)))
    value, dt = cpu_fista(np, arm, steps, args.warmup, batch)
    line = {
        'impl': 'reference',
        'metric': 'batched_fista_problem_iterations_per_second', 'value': value, 'unit': 'problem-iters/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup, 'ms_per_step': dt / steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': fista_config(args.batch, args.gpus),
        'cpu_baseline': {'value': value, 'unit': 'problem-iters/s', 'cores': cores, 'kind': arm.kind,
                         'sample': '%d problems x %d FISTA iterations in one lasso.solve call (%.1f s, set-up GEMMs '
                                   'included), %s' % (batch, steps, dt, arm.where)},
        'e2e': {'value': value, 'unit': 'problem-iters/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a GPU (the product path has no CPU fallback)')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    group = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
        group = dist.group.WORLD

    from decomp_b200 import dictionary_learning, lasso, nmf, ops
    legs = set(args.legs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_regions(region, repeats):
        """region(r) enqueues region r on the current stream.  Returns ([ms per region, max over ranks], launches of
        one region, clocks sampled over all regions)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out, launches = [], 0
        barrier()
        with ClockSampler(local) as cs:
            for r in range(repeats):
                barrier()
                before = ops.LAUNCHES
                e0.record()
                region(r)
                e1.record()
                torch.cuda.synchronize()
                launches = ops.LAUNCHES - before
                out.append(max_over_ranks(e0.elapsed_time(e1)))
        barrier()
        return out, launches, cs.summary()

    def wall_calls(call, repeats):
        """call() is a blocking host-level API call; returns the median wall time in seconds, max over ranks."""
        times = []
        for _ in range(repeats):
            barrier()
            t0 = time.perf_counter()
            call()
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        return max_over_ranks(median(times)), times

    hbm_peak, hbm_src = load_peaks()
    tf_peak, tf_src = tf32_peak()
    traffic = load_traffic()
    dmma_peak = ops.probe_dmma_tflops()
    K, W, R = max(1, args.steps), max(0, args.warmup), max(1, args.repeats)
    out = {}
    clocks = None            # the primary leg's record
    leg_clocks = {}          # every other timed leg, by name
    cpu_arm_mod = None
    if rank == 0 and world == 1 and 'cpu' in legs:
        from oracle import cpu_arm as cpu_arm_mod
        cpu_cores = cpu_arm_mod.set_blas_threads()
        cpu = cpu_arm_mod.load()

    # ---------------------------------------------------------------- batched FISTA (primary)
    if 'fista' in legs:
        B, k, f, alpha = args.batch, FISTA['k'], FISTA['f'], FISTA['alpha']
        y, A = fista_data_device(torch, B, k, f, rank, device)
        total = W + K * (R + 1)
        solver = lasso.LassoSolver(y, A, alpha, None, 0.0, total, 'fista', False)
        solver.iterate(0, W)
        solver.iterate(W, W + K)                               # untimed: the launch pattern of a region has run once
        ms_list, launches, clocks = timed_regions(lambda r: solver.iterate(W + K * (r + 1), W + K * (r + 2)), R)
        ms = median(ms_list)
        xs = solver.finish().result
        finite = bool(torch.isfinite(xs).all().item())
        nnz = float((xs != 0).double().mean().item())
        del solver, xs
        sec = ms * 1e-3
        per_launch = K / float(max(launches, 1))            # iterations one launch of the dominant kernel runs
        flops_launch = 2.0 * B * k * k * per_launch         # 2 B k^2 per iteration
        bytes_launch = 5.0 * B * k * 8                      # read yAh, w, x_prev; write x_new, w_next: once per launch
        t_launch = sec / max(launches, 1)
        resident = per_launch > 1.0
        out['fista'] = {
            'value': B * world * K / sec, 'iters_per_s': K / sec, 'ms_per_step': ms / K, 'launches': launches,
            'finite': finite, 'nonzero_fraction': nnz,
            'timing': {'regions_ms': ms_list, 'statistic': 'median of %d timed regions of %d iterations each' % (R, K),
                       'min_ms': min(ms_list), 'max_ms': max(ms_list)},
            'roofline': {'bound': 'tensor', 'achieved': flops_launch / t_launch / 1e12, 'peak': dmma_peak,
                         'unit': 'TFLOP/s', 'frac': flops_launch / t_launch / 1e12 / dmma_peak,
                         'traffic': traffic.get('lasso_resident_dram_bytes_per_launch' if resident
                                                else 'fista_proxq_dram_bytes_per_launch'),
                         'traffic_source': traffic.get('lasso_resident_source' if resident else 'fista_proxq_source'),
                         'kernel': ('lasso_resident_kernel (iterate resident in shared memory, Q streamed by TMA, '
                                    'DMMA GEMM + ISTA/FISTA update, up to 32 iterations per launch)') if resident else
                                   'gemm_f64_proxq_kernel (one FISTA iteration per launch)',
                         'iterations_per_launch': per_launch,
                         'peak_source': 'FP64 tensor (DMMA.8x8x4) issue rate measured live by '
                                        'decomp_probe_dmma_tflops(); MEASURED_PEAKS.json has no FP64 entry',
                         'algorithmic_flops_per_launch': flops_launch,
                         'hbm': {'algorithmic_bytes_per_launch': bytes_launch,
                                 'achieved_gbs': bytes_launch / t_launch / 1e9, 'peak_gbs': hbm_peak,
                                 'frac': bytes_launch / t_launch / 1e9 / hbm_peak, 'peak_source': hbm_src}},
        }
        # ---- the TF32-split (tcgen05) variant of the same iterations: reported beside, never as the headline
        if 'tf32' in legs:
            R32 = min(R, 5)
            solver = lasso.LassoSolver(y, A, alpha, None, 0.0, W + K * (R32 + 1), 'fista', False, precision='tf32x3')
            solver.iterate(0, W + K)
            ms32_list, launches32, c32 = timed_regions(
                lambda r: solver.iterate(W + K * (r + 1), W + K * (r + 2)), R32)
            leg_clocks['fista_tf32x3'] = c32
            ms32 = median(ms32_list)
            x32 = solver.finish().result
            ref = lasso.LassoSolver(y, A, alpha, None, 0.0, W + K * (R32 + 1), 'fista', False)
            ref.iterate(0, W + K * (R32 + 1))
            x64 = ref.finish().result
            dev_err = float(((x32 - x64).abs().max() / x64.abs().max()).item())
            del solver, ref, x32, x64
            t32 = ms32 * 1e-3 / K
            bytes32 = 12.0 * B * k * 4            # GEMM 12 B/element + FP64 pass 36 B/element = 48 B/element
            out['fista']['tf32x3'] = {
                'value': B * world * K / (ms32 * 1e-3), 'unit': 'problem-iters/s', 'iters_per_s': K / (ms32 * 1e-3),
                'ms_per_step': ms32 / K, 'launches': launches32, 'dtype': 'tf32x3 GEMM (FP32 accumulate) + f64 update',
                'max_rel_diff_x_vs_fp64': dev_err,
                'roofline': {'bound': 'hbm', 'achieved': bytes32 / t32 / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                             'frac': bytes32 / t32 / 1e9 / hbm_peak, 'traffic': traffic.get('tf32x3_dram_bytes_per_iteration'),
                             'kernel': 'tf32x3_gemm_kernel (tcgen05.mma kind::tf32, TMEM) + proxq_apply_kernel',
                             'algorithmic_bytes_per_iteration': bytes32, 'peak_source': hbm_src},
            }
        # ---- end to end through the public API: ordinary numpy arrays first, page-locked ones beside
        # ---- the other reading of "A 256x1024" (SURVEY.md 8, C2 caveat): 1024 atoms over 256 channels (compressed
        # sensing, over-complete).  The iterate is 1024 doubles wide: one gemm_f64_proxq_kernel launch per iteration
        if 'configs' in legs:
            k2, f2 = FISTA['f'], FISTA['k']
            y2, A2 = fista_data_device(torch, B, k2, f2, rank, device)
            s2 = lasso.LassoSolver(y2, A2, alpha, None, 0.0, 100000, 'fista', False)
            s2.iterate(0, 3)
            ms2, l2_, c2_ = timed_regions(lambda r: s2.iterate(3 + 5 * r, 8 + 5 * r), 3)
            leg_clocks['fista_1024x256'] = c2_
            t2 = median(ms2) * 1e-3 / 5
            fl2, by2 = 2.0 * B * k2 * k2, 5.0 * B * k2 * 8
            out['fista_overcomplete'] = {
                'value': B * world / t2, 'unit': 'problem-iters/s', 'iters_per_s': 1.0 / t2, 'ms_per_step': t2 * 1e3,
                'launches_per_iteration': l2_ / 5.0,
                'finite': bool(torch.isfinite(s2.X).all().item()),
                'roofline': {'bound': 'tensor', 'achieved': fl2 / t2 / 1e12, 'peak': dmma_peak, 'unit': 'TFLOP/s',
                             'frac': fl2 / t2 / 1e12 / dmma_peak, 'traffic': None,
                             'kernel': 'gemm_f64_proxq_kernel (one FISTA iteration per launch, K = N = 1024)',
                             'algorithmic_flops_per_iteration': fl2,
                             'hbm': {'algorithmic_bytes_per_iteration': by2, 'achieved_gbs': by2 / t2 / 1e9,
                                     'peak_gbs': hbm_peak}},
                'config': {'workload': 'batched FISTA Lasso, %d problems per GPU, A (%d,%d): the compressed-sensing '
                                       'reading of BASELINE.json configs[1] (SURVEY.md 8 C2 caveat), float64, tol=0'
                                       % (B, k2, f2)}}
            del s2, y2, A2
            torch.cuda.empty_cache()
        if 'e2e' in legs:
            y_np, A_np = y.cpu().numpy(), A.cpu().numpy()              # pageable host arrays
            del y
            torch.cuda.empty_cache()
            lasso.solve(y_np[:8192], A_np, alpha, tol=0.0, method='fista', maxiter=3)    # allocator / module warm-up
            res = {}

            def call(yy, AA):
                res['x'] = None     # drop the previous result: its page-locked block returns to torch's host cache
                res['it'], res['x'] = lasso.solve(yy, AA, alpha, tol=0.0, method='fista', maxiter=K)

            call(y_np, A_np)
            e2e, e2e_times = wall_calls(lambda: call(y_np, A_np), 5)
            assert res['it'] == K - 1 and res['x'].shape == (B, k)
            h2d, d2h = (y_np.nbytes + A_np.nbytes) / K, res['x'].nbytes / K
            note = ('one lasso.solve(maxiter=steps) call with %s numpy arrays in and a numpy array out: H2D of y and '
                    'A, set-up GEMMs, steps iterations, D2H of x, all inside the timed region (tol = 0: the call runs '
                    'the batch in row chunks so that copies overlap the iterations); median of 5 calls, max over '
                    'ranks; bytes are per call / steps')
            out['fista']['e2e'] = {'value': B * world * K / e2e, 'unit': 'problem-iters/s',
                                   'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'ms_per_call': e2e * 1e3,
                                   'host_memory': 'pageable', 'calls_ms': [t * 1e3 for t in e2e_times],
                                   'h2d_gbs_per_gpu_if_copy_only': (y_np.nbytes + A_np.nbytes) / e2e / 1e9,
                                   'note': note % 'ordinary (pageable)'}
            yh = torch.empty((B, f), dtype=torch.float64, pin_memory=True)
            Ah = torch.empty((k, f), dtype=torch.float64, pin_memory=True)
            yh.copy_(torch.from_numpy(y_np))
            Ah.copy_(torch.from_numpy(A_np))
            del y_np, A_np
            yp, Ap = yh.numpy(), Ah.numpy()
            call(yp, Ap)
            e2ep, e2ep_times = wall_calls(lambda: call(yp, Ap), 5)
            out['fista']['e2e_pinned'] = {'value': B * world * K / e2ep, 'unit': 'problem-iters/s',
                                          'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                                          'ms_per_call': e2ep * 1e3, 'host_memory': 'page-locked',
                                          'calls_ms': [t * 1e3 for t in e2ep_times], 'note': note % 'page-locked'}
            # what the host-to-device link gives this rank while every rank copies at once: the floor of the call above
            ydev = torch.empty((B, f), dtype=torch.float64, device=device)
            ydev.copy_(yh, non_blocking=True)
            probe_ms, _, _ = timed_regions(lambda r: ydev.copy_(yh, non_blocking=True), 3)
            t_probe = median(probe_ms) * 1e-3
            for key in ('e2e', 'e2e_pinned'):
                out['fista'][key]['h2d_copy_only'] = {
                    'ms': t_probe * 1e3, 'gbs_per_gpu': yh.numel() * 8 / t_probe / 1e9,
                    'gbs_all_gpus': world * yh.numel() * 8 / t_probe / 1e9,
                    'note': 'bare cudaMemcpyAsync of y from page-locked memory, all ranks at once, max over ranks: '
                            'the part of ms_per_call no kernel can remove'}
            del ydev
            res.clear()
            del yh, Ah, yp, Ap
        else:
            del y
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- NMF-MU (secondary): weak and strong scaling
    def nmf_leg(n_rows, label, masked=False, shape=NMF, tf32=False):
        f, k = shape['f'], shape['k']

        def sweep_traffic(key):
            # DRAM bytes of one sweep from the committed ncu pass at `sweep_rows` rows, scaled to this leg's rows
            v = traffic.get(key)
            return v * n_rows / float(traffic.get('sweep_rows', n_rows)) if v is not None else None

        Kn, Wn, Rn = max(2, args.nmf_steps), 2, 3
        y, D0, mask = nmf_data_device(torch, n_rows, f, k, rank, device, masked=masked)
        X = torch.ones((n_rows, k), dtype=torch.float64, device=device)
        solver = nmf.MuSolver(y, D0, X, 0.0, mask=mask, group=group)
        for it in range(1, Wn + 1):
            solver.sweep(it)
        solver.comm_events = [] if world > 1 else None
        ms_list, launches, c2 = timed_regions(
            lambda r: [solver.sweep(it) for it in range(Wn + 1 + r * Kn, Wn + 1 + (r + 1) * Kn)], Rn)
        ms = median(ms_list)
        comm_ms = None
        if solver.comm_events:
            torch.cuda.synchronize()
            per_sweep = [sum(a.elapsed_time(b) for a, b in evs) for evs in solver.comm_events]
            comm_ms = max_over_ranks(median(per_sweep))
        D = solver.Dbuf[(Wn + Rn * Kn) % 2]
        finite = bool(torch.isfinite(D).all().item() and torch.isfinite(X).all().item())
        sec = ms * 1e-3
        n_total = n_rows * world if label != 'strong' else args.strong_rows
        if masked:
            flops_sweep = 12.0 * n_rows * k * f                                   # no re-association under a mask
            bytes_sweep = 2.0 * n_rows * f * 8 + 2.0 * n_rows * k * 8             # Y*M and M once (row-block fused)
        else:
            flops_sweep = 4.0 * n_rows * k * f + 4.0 * n_rows * k * k + 4.0 * k * k * f
            bytes_sweep = 1.0 * n_rows * f * 8 + 2.0 * n_rows * k * 8             # Y once + C r/w
        t_sweep = sec / Kn
        res = {
            'metric': 'nmf_mu_row_iterations_per_second', 'value': n_total * Kn / sec, 'unit': 'row-iters/s',
            'iters_per_s': Kn / sec, 'ms_per_step': ms / Kn, 'steps': Kn, 'warmup': Wn, 'launches': launches,
            'rows_per_gpu': n_rows, 'rows_total': n_total, 'finite': finite,
            'timing': {'regions_ms': ms_list, 'statistic': 'median of %d regions of %d sweeps' % (Rn, Kn)},
            'allreduce_ms_per_sweep': comm_ms,
            'allreduce_bytes_per_sweep': (2 * k * f if masked else k * f + k * k) * 8 if world > 1 else 0,
            'roofline': {'bound': 'tensor', 'achieved': flops_sweep / t_sweep / 1e12, 'peak': dmma_peak,
                         'unit': 'TFLOP/s', 'frac': flops_sweep / t_sweep / 1e12 / dmma_peak,
                         'traffic': sweep_traffic('nmf_sweep_dram_bytes' if not masked else 'nmf_masked_sweep_dram_bytes'),
                         'traffic_source': traffic.get('sweep_source'),
                         'kernel': 'whole sweep (per GPU): gemm_f64_kernel<NT, MU_NUM> (Y D^T) + gemm_f64_kernel<TN> '
                                   '(X^T Y) dominate with 2nkf flop each' if not masked else
                                   'whole masked sweep (per GPU): six 2nkf GEMMs, [n,f] intermediate fused where the '
                                   'B2B kernel applies',
                         'algorithmic_flops_per_sweep': flops_sweep,
                         'hbm': {'algorithmic_bytes_per_sweep': bytes_sweep,
                                 'achieved_gbs': bytes_sweep / t_sweep / 1e9, 'peak_gbs': hbm_peak}},
        }
        if tf32:
            # the same sweeps with the three big contractions on the tcgen05 tensor cores (TF32 split, FP32 accumulate)
            D64 = D.clone()
            del solver
            torch.cuda.empty_cache()
            X.fill_(1.0)
            solver = nmf.MuSolver(y, D0, X, 0.0, mask=mask, group=group, precision='tf32x3')
            for it in range(1, Wn + 1):
                solver.sweep(it)
            ms32_list, launches32, c3 = timed_regions(
                lambda r: [solver.sweep(it) for it in range(Wn + 1 + r * Kn, Wn + 1 + (r + 1) * Kn)], Rn)
            t32 = median(ms32_list) * 1e-3 / Kn
            D32 = solver.Dbuf[(Wn + Rn * Kn) % 2]
            err = float(((D32 - D64).abs().max() / D64.abs().max()).item())
            by32 = 2.0 * n_rows * f * 8 + 4.0 * n_rows * k * 8     # y and y^T as TF32 pairs (8 B per element each), x
            fl32 = 3.0 * (4.0 * n_rows * k * f + 4.0 * n_rows * k * k)
            res['tf32x3'] = {
                'value': n_total / t32, 'unit': 'row-iters/s', 'iters_per_s': 1.0 / t32, 'ms_per_step': t32 * 1e3,
                'launches': launches32, 'dtype': 'tf32x3 GEMMs (tcgen05, FP32 accumulate in TMEM, FP64 slab sums) + f64 updates',
                'max_rel_diff_D_vs_fp64': err, 'sweeps_compared': Wn + Rn * Kn,
                'speedup_vs_fp64': t_sweep / t32,
                'roofline': {'bound': 'tensor', 'achieved': fl32 / t32 / 1e12, 'peak': tf_peak, 'unit': 'TFLOP/s',
                             'frac': fl32 / t32 / 1e12 / tf_peak, 'traffic': sweep_traffic('nmf_tf32x3_sweep_dram_bytes'),
                             'peak_source': tf_src,
                             'algorithmic_flops_per_sweep': fl32,
                             'note': 'three TF32 products per FP64-equivalent product (hi*hi + hi*lo + lo*hi): 3 x (4nkf + '
                                     '4nk^2) flop; the tensor pipe under the power cap bounds the sweep, HBM does not',
                             'kernel': 'tf32x3_gemm_pair_kernel<XUPD> (y D^T + ratio) and <PARTIAL> (x^T y), '
                                       'tcgen05.mma.cta_group::2.kind::tf32',
                             'hbm': {'algorithmic_bytes_per_sweep': by32, 'achieved_gbs': by32 / t32 / 1e9,
                                     'peak_gbs': hbm_peak, 'frac': by32 / t32 / 1e9 / hbm_peak,
                                     'note': 'y is read once row-major (y D^T) and once transposed (x^T y), both as '
                                             'TF32 pairs'}}}
            if masked:
                # masked model: six TF32-split GEMMs and the [n, f] intermediate as a TF32 pair through HBM -- bound by
                # HBM, not by the tensor pipe (BASELINE.json configs[4]: "HBM-bound elementwise path")
                fl32m = 3.0 * 12.0 * n_rows * k * f
                by32m = 56.0 * n_rows * f + 10.0 * n_rows * k * 8
                res['tf32x3']['roofline'] = {
                    'bound': 'hbm', 'achieved': by32m / t32 / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': by32m / t32 / 1e9 / hbm_peak, 'traffic': sweep_traffic('nmf_masked_tf32x3_sweep_dram_bytes'),
                    'algorithmic_bytes_per_sweep': by32m,
                    'note': 'per row and feature: y*mask as a TF32 pair row-major and transposed (8 + 8 B), the FP32 '
                            'mask twice (4 + 4 B), f = (x D)*mask written and read as a TF32 pair twice (2 x 16 B)',
                    'kernel': 'tf32x3_gemm_pair_kernel<FMASK> ((x D)*mask -> TF32 pair), <STORE> (f D^T), <XUPD>, '
                              '<PARTIAL> (x^T (y*m), x^T f)',
                    'tensor': {'algorithmic_flops_per_sweep': fl32m, 'achieved_tflops': fl32m / t32 / 1e12,
                               'peak_tflops': tf_peak, 'frac': fl32m / t32 / 1e12 / tf_peak, 'peak_source': tf_src}}
            res['tf32x3']['clocks'] = c3
            del D64, D32
        del solver, y, X, D0, D, mask
        torch.cuda.empty_cache()
        return res, c2

    def guarded(name, fn):
        """Runs a secondary leg; an exception (raised on every rank alike: the legs are deterministic) is recorded
        under the leg's key instead of costing the primary line."""
        try:
            fn()
        except Exception as exc:                                    # noqa: BLE001
            import traceback
            out.setdefault('errors', {})[name] = '%s: %s' % (type(exc).__name__, str(exc)[:300])
            sys.stderr.write('leg %s failed:\n%s\n' % (name, traceback.format_exc()))
            torch.cuda.empty_cache()

    def leg_nmf():
        res, c2 = nmf_leg(args.rows, 'weak', tf32='tf32' in legs)
        res['config'] = {'workload': 'NMF-MU l2, %d rows per GPU x %d features, k=%d, float64, tol=0 '
                                     '(BASELINE.json configs[2] shape, weak scaling)' % (args.rows, NMF['f'], NMF['k'])}
        res['scaling'] = 'weak'
        out['nmf'] = res
        leg_clocks['nmf'] = c2

    if 'nmf' in legs:
        guarded('nmf', leg_nmf)
    def leg_nmf_strong():
        n_tot = args.strong_rows
        lo, hi = rank * n_tot // world, (rank + 1) * n_tot // world
        if world == 1 and 'nmf' in out and args.rows == n_tot:
            res = dict(out['nmf'])
            res['same_run_as_secondary'] = True
        else:
            res, c2 = nmf_leg(hi - lo, 'strong', tf32='tf32' in legs)
            leg_clocks['nmf_strong'] = c2
        res['config'] = {'workload': 'NMF-MU l2, %d rows IN TOTAL x %d features, k=%d, float64, tol=0, sample axis '
                                     'sharded over %d GPU(s) with all-reduce of X^T Y [k,f] and X^T X [k,k] per sweep '
                                     '(BASELINE.json configs[2], strong scaling)' % (n_tot, NMF['f'], NMF['k'], world)}
        res['scaling'] = 'strong'
        res['sweeps_per_s'] = res['iters_per_s']
        out['nmf_strong'] = res

    if 'nmf_strong' in legs:
        guarded('nmf_strong', leg_nmf_strong)
    def leg_nmf_e2e():
        n_e, f, k = args.e2e_nmf_rows, NMF['f'], NMF['k']
        y_d, D0_d, _ = nmf_data_device(torch, n_e, f, k, rank, device)
        y_np, D_np = y_d.cpu().numpy(), D0_d.cpu().numpy()
        del y_d, D0_d
        sweeps = max(2, args.nmf_steps)
        nmf.solve(y_np[:4096], D_np, tol=0.0, maxiter=3)
        res = {}

        def call_nmf():
            res.clear()
            res['out'] = nmf.solve(y_np, D_np, tol=0.0, maxiter=sweeps + 1)

        call_nmf()
        t_e, times = wall_calls(call_nmf, 3)
        it_e, D_e, x_e = res['out']
        assert it_e == sweeps + 1 and x_e.shape == (n_e, k)
        out['nmf_e2e'] = {'value': n_e * world * sweeps / t_e, 'unit': 'row-iters/s', 'ms_per_call': t_e * 1e3,
                          'rows_per_gpu': n_e, 'sweeps': sweeps, 'host_memory': 'pageable',
                          'h2d_bytes_per_step': (y_np.nbytes + D_np.nbytes + n_e * k * 8) / sweeps,
                          'd2h_bytes_per_step': (x_e.nbytes + D_e.nbytes) / sweeps,
                          'calls_ms': [t * 1e3 for t in times],
                          'note': 'one nmf.solve(tol=0, maxiter=sweeps+1) call with ordinary numpy arrays in and out '
                                  '(H2D of y, D and the default x = ones; D2H of D and x), median of 3 calls'}
        res.clear()
        del y_np, D_np, D_e, x_e
        torch.cuda.empty_cache()

    if 'nmf_e2e' in legs:
        guarded('nmf_e2e', leg_nmf_e2e)

    # ---------------------------------------------------------------- BASELINE configs[3] and [4]
    def leg_configs():
        extra = out.setdefault('extra_configs', {})        # filled in place: a failure keeps what was measured
        res, c2 = nmf_leg(args.c5_rows, 'weak', masked=True, shape=C5, tf32='tf32' in legs)
        leg_clocks['c5_masked_nmf'] = c2
        res['config'] = {'workload': 'masked NMF-MU l2, %d rows per GPU x %d, k=%d, 10 %% missing, float64 '
                                     '(BASELINE.json configs[4], per-GPU shard)' % (args.c5_rows, C5['f'], C5['k'])}
        extra['c5_masked_nmf_sweep'] = res
        # masked FISTA iteration (the Lasso-C step of configs[4])
        n5, f5, k5 = args.c5_rows, C5['f'], C5['k']
        y5, A5, m5 = nmf_data_device(torch, n5, f5, k5, rank, device, masked=True)
        s5 = lasso.LassoSolver(y5, A5, 0.1, None, 0.0, 100000, 'fista', False, mask=m5)
        s5.iterate(0, 3)
        ms5, l5, c5 = timed_regions(lambda r: s5.iterate(3 + 5 * r, 8 + 5 * r), 3)
        leg_clocks['c5_masked_fista'] = c5
        t5 = median(ms5) * 1e-3 / 5
        fl5 = 4.0 * n5 * k5 * f5
        by5 = n5 * f5 * 8.0 + 5.0 * n5 * k5 * 8
        extra['c5_masked_fista_iter'] = {
            'value': n5 * world / t5, 'unit': 'problem-iters/s', 'ms_per_step': t5 * 1e3, 'launches_per_iteration': l5 / 5.0,
            'rows_per_gpu': n5,
            'roofline': {'bound': 'tensor', 'achieved': fl5 / t5 / 1e12, 'peak': dmma_peak, 'unit': 'TFLOP/s',
                         'frac': fl5 / t5 / 1e12 / dmma_peak, 'traffic': traffic.get('masked_fista_dram_bytes_per_iteration'),
                         'algorithmic_flops_per_iteration': fl5,
                         'hbm': {'algorithmic_bytes_per_iteration': by5, 'achieved_gbs': by5 / t5 / 1e9,
                                 'peak_gbs': hbm_peak}},
            'config': {'workload': 'masked FISTA iteration ((w A)*M) A^H, %d problems per GPU, A (%d,%d), float64 '
                                   '(the Lasso-C step of BASELINE.json configs[4])' % (n5, k5, f5)}}
        if 'tf32' in legs:
            x64 = s5.X.clone()
            it_done = 3 + 5 * 3
            del s5
            torch.cuda.empty_cache()
            s5 = lasso.LassoSolver(y5, A5, 0.1, None, 0.0, 100000, 'fista', False, mask=m5, precision='tf32x3')
            s5.iterate(0, 3)
            ms5t, l5t, c5t = timed_regions(lambda r: s5.iterate(3 + 5 * r, 8 + 5 * r), 3)
            t5t = median(ms5t) * 1e-3 / 5
            err5 = float(((s5.X - x64).abs().max() / x64.abs().max()).item())
            by5t = n5 * f5 * (4.0 + 8.0 + 8.0) + 8.0 * n5 * k5 * 8
            extra['c5_masked_fista_iter']['tf32x3'] = {
                'value': n5 * world / t5t, 'unit': 'problem-iters/s', 'ms_per_step': t5t * 1e3,
                'launches_per_iteration': l5t / 5.0, 'speedup_vs_fp64': t5 / t5t,
                'max_rel_diff_x_vs_fp64': err5, 'iterations_compared': it_done,
                'dtype': 'tf32x3 GEMMs (tcgen05, FP32 accumulate in TMEM) + f64 threshold / momentum pass',
                'roofline': {'bound': 'hbm', 'achieved': by5t / t5t / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                             'frac': by5t / t5t / 1e9 / hbm_peak, 'traffic': None,
                             'algorithmic_bytes_per_iteration': by5t,
                             'note': 'per problem and feature: the FP32 mask (4 B) and t = (w A)*mask written and read as '
                                     'a TF32 pair (8 + 8 B); per problem and atom: w pair, c, x_prev in, x, w pair out',
                             'tensor': {'achieved_tflops': 3.0 * fl5 / t5t / 1e12, 'peak_tflops': tf_peak,
                                        'frac': 3.0 * fl5 / t5t / 1e12 / tf_peak}},
                'clocks': c5t}
            del x64
        del s5, y5, A5, m5
        torch.cuda.empty_cache()
        # dictionary-learning minibatch step (configs[3]); the ranks share each minibatch when world > 1
        n4, f4, k4, mb = C4['n'], C4['f'], C4['k'], C4['minibatch']
        g = torch.Generator(device=device)
        g.manual_seed(7)

        def crandn(*s):
            return torch.complex(torch.randn(s, dtype=torch.float64, device=device, generator=g),
                                 torch.randn(s, dtype=torch.float64, device=device, generator=g))

        Dt = crandn(k4, f4)
        y4 = (crandn(n4, k4) * torch.rand((n4, k4), dtype=torch.float64, device=device, generator=g)) @ Dt
        y4 += 0.1 * crandn(n4, f4)
        D04 = Dt + 0.2 * crandn(k4, f4)
        m4 = (torch.rand((n4, f4), dtype=torch.float64, device=device, generator=g) > 0.1).double()
        if world > 1:                        # row-sharded inputs: every rank keeps its block of the same data
            lo4, hi4 = rank * n4 // world, (rank + 1) * n4 // world
            y4, m4 = y4[lo4:hi4].contiguous(), m4[lo4:hi4].contiguous()
        kw4 = dict(tol=0.0, minibatch=mb, maxiter=2, lasso_method='fista', lasso_iter=10, mask=m4, random_seed=0,
                   group=group)
        nw = min(2 * mb // world, y4.shape[0])
        dictionary_learning.solve(y4[:nw], D04, 0.1, **dict(kw4, mask=m4[:nw], minibatch=mb // world))   # warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        before = ops.LAUNCHES
        e0.record()
        it4, D4, x4 = dictionary_learning.solve(y4, D04, 0.1, **kw4)
        e1.record()
        torch.cuda.synchronize()
        steps4 = n4 // mb
        ms4 = max_over_ranks(e0.elapsed_time(e1)) / steps4
        fl4 = (10 * 16 * k4 * f4 + 8 * k4 * f4) * float(mb) + 2.0 * k4 * k4 * f4 * mb + 8.0 * k4 * k4 * f4
        extra['c4_dl_masked_step'] = {
            'value': 1e3 / ms4, 'unit': 'minibatch-steps/s', 'ms_per_step': ms4, 'minibatch': mb,
            'launches_per_step': (ops.LAUNCHES - before) / float(steps4),
            'finite': bool(torch.isfinite(torch.view_as_real(D4)).all().item()),
            'roofline': {'bound': 'tensor', 'achieved': fl4 / ms4 / 1e9, 'peak': dmma_peak, 'unit': 'TFLOP/s',
                         'frac': fl4 / ms4 / 1e9 / dmma_peak / world, 'traffic': None,
                         'algorithmic_flops_per_step': fl4,
                         'note': 'flops = 10 masked FISTA iterations + x^H(y*m) + Hermitian half of the masked '
                                 'statistics (2 k^2 f per row) + atom update; frac is per GPU'},
            'config': {'workload': 'dictionary learning block_cd, complex128 y %dx%d (one epoch of %d minibatch steps '
                                   'of %d rows; BASELINE.json configs[3] has n=200000), k=%d, 10 %% mask, fista x10, '
                                   'rows sharded over %d GPU(s), statistics reduce-scattered along f' % (n4, f4, steps4, mb, k4, world)}}
        del y4, D04, m4, Dt, D4, x4
        torch.cuda.empty_cache()
        # configs[0]: NMF-MU 1000 x 200, k = 20, 100 sweeps through nmf.solve with host arrays (one cooperative launch)
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        import golden_cases as gc
        y1, D1, _ = gc._nmf_data(1000, 200, 20, 0, 'l2', reference_order=False)
        nmf.solve(y1, D1.copy(), tol=0.0, maxiter=101)
        t1, times1 = wall_calls(lambda: nmf.solve(y1, D1.copy(), tol=0.0, maxiter=101), 5)
        extra['c1_nmf_small'] = {
            'value': 100.0 / t1, 'unit': 'sweeps/s', 'ms_per_call': t1 * 1e3, 'calls_ms': [t * 1e3 for t in times1],
            'config': {'workload': 'NMF-MU l2, y 1000x200 float64, k=20, 100 sweeps, one nmf.solve call with numpy '
                                   'arrays in and out (BASELINE.json configs[0]); all sweeps in one cooperative launch '
                                   '(decomp_nmf_mu_small_f64)'}}
        if cpu_arm_mod is not None:
            t0 = time.perf_counter()
            cpu.nmf(y1, D1.copy(), tol=0.0, maxiter=101)
            dt = time.perf_counter() - t0
            extra['c1_nmf_small']['cpu_baseline'] = {'value': 100.0 / dt, 'unit': 'sweeps/s', 'cores': cpu_cores,
                                                     'kind': cpu.kind, 'sample': 'the same call (%.3f s), %s' % (dt, cpu.where)}

    def leg_configs_cpu():
        """SURVEY.md 8(d) CPU plan for configs[3] and configs[4]: the reference at a stated reduced n, in rows per
        second (its [n, f] / [mb, f, k] temporaries do not fit or finish at full size)."""
        extra = out.setdefault('extra_configs', {})
        rng = np.random.RandomState(5)
        # ---- configs[4]: masked NMF sweep and masked FISTA iteration at 131072 / 32768 rows x 1024, k = 128
        f5, k5 = C5['f'], C5['k']
        n5 = 131072
        Dt = np.maximum(rng.randn(k5, f5), 0.0)
        y5 = np.maximum(rng.randn(n5, k5), 0.0).dot(Dt) + 0.1 * rng.randn(n5, f5)
        D05 = np.maximum(Dt + 0.3 * rng.randn(k5, f5), 0.1)
        m5 = (rng.rand(n5, f5) > 0.1).astype(np.float64)
        t0 = time.perf_counter()
        cpu.nmf(y5, D05.copy(), tol=0.0, maxiter=3, mask=m5)
        dt = time.perf_counter() - t0
        if 'c5_masked_nmf_sweep' in extra:
            extra['c5_masked_nmf_sweep']['cpu_baseline'] = {
                'value': n5 * 2 / dt, 'unit': 'row-iters/s', 'cores': cpu_cores, 'kind': cpu.kind,
                'sample': 'n = %d rows x %d, k=%d, 10 %% mask, 2 sweeps of nmf.solve (%.1f s), %s'
                          % (n5, f5, k5, dt, cpu.where)}
        b5 = 32768
        t0 = time.perf_counter()
        cpu.lasso(y5[:b5], D05, 0.1, tol=0.0, method='fista', maxiter=10, mask=m5[:b5])
        dt = time.perf_counter() - t0
        if 'c5_masked_fista_iter' in extra:
            extra['c5_masked_fista_iter']['cpu_baseline'] = {
                'value': b5 * 10 / dt, 'unit': 'problem-iters/s', 'cores': cpu_cores, 'kind': cpu.kind,
                'sample': '%d problems, A (%d,%d), 10 %% mask, 10 FISTA iterations in one lasso.solve call (%.1f s, '
                          'set-up included), %s' % (b5, k5, f5, dt, cpu.where)}
        del y5, m5
        # ---- configs[3]: ONE minibatch step of 128 rows (the reference materialises [mb, f, k] complex, 2.1 GB here, and
        # makes several single-threaded passes over the 8.6 GB statistics per step: ~40 s on 16 cores)
        f4, k4, mb4 = C4['f'], C4['k'], 128

        def crandn(*shape):
            return rng.randn(*shape) + 1j * rng.randn(*shape)

        Dt4 = crandn(k4, f4)
        y4 = (crandn(mb4, k4) * rng.rand(mb4, k4)).dot(Dt4) + 0.1 * crandn(mb4, f4)
        m4 = (rng.rand(mb4, f4) > 0.1).astype(np.float64)
        t0 = time.perf_counter()
        cpu.dictionary_learning(y4 * m4, Dt4 + 0.2 * crandn(k4, f4), 0.1, tol=0.0, minibatch=mb4, maxiter=2,
                                lasso_method='fista', lasso_iter=10, mask=m4, random_seed=0)
        dt = time.perf_counter() - t0
        if 'c4_dl_masked_step' in extra:
            gpu_rows = extra['c4_dl_masked_step']['minibatch'] * extra['c4_dl_masked_step']['value']
            extra['c4_dl_masked_step']['rows_per_s'] = gpu_rows
            extra['c4_dl_masked_step']['cpu_baseline'] = {
                'value': mb4 / dt, 'unit': 'rows/s', 'cores': cpu_cores, 'kind': cpu.kind,
                'sample': 'one minibatch step of %d rows (complex128, f=%d, k=%d, 10 %% mask, fista x10; '
                          '%.1f s); a step also pays fixed passes over the [k, f, k] statistics, so rows/s grows '
                          'with the minibatch; %s' % (mb4, f4, k4, dt, cpu.where)}

    if 'configs' in legs:
        guarded('configs', leg_configs)
        if cpu_arm_mod is not None:
            guarded('configs_cpu', leg_configs_cpu)

    # ---------------------------------------------------------------- multi-GPU parity self-check (outside timing)
    def leg_parity():
        out['parity'] = parity_check(np, torch, dist, world, rank, group, device)

    if 'parity' in legs:
        guarded('parity', leg_parity)

    # ---------------------------------------------------------------- CPU baselines (rank 0, N = 1 only)
    cpu_f = cpu_n = None
    if cpu_arm_mod is not None:
        v, dt = cpu_fista(np, cpu, 100, 1, CPU_FISTA_BATCH)
        cpu_f = {'value': v, 'unit': 'problem-iters/s', 'cores': cpu_cores, 'kind': cpu.kind,
                 'sample': '%d problems x 100 FISTA iterations (%.1f s), %s' % (CPU_FISTA_BATCH, dt, cpu.where)}
        if 'nmf' in legs or 'nmf_strong' in legs:
            v, dt = cpu_nmf(np, cpu, 2, CPU_NMF_ROWS)
            cpu_n = {'value': v, 'unit': 'row-iters/s', 'cores': cpu_cores, 'kind': cpu.kind,
                     'sample': 'n = %d rows x %d features, k=%d, 2 sweeps of nmf.solve (%.1f s), %s'
                               % (CPU_NMF_ROWS, NMF['f'], NMF['k'], dt, cpu.where)}

    if rank == 0:
        if 'fista' in out:
            prim = out['fista']
            cfg = fista_config(args.batch, world)
            cfg['timing'] = prim['timing']['statistic']
            line = {
                'metric': 'batched_fista_problem_iterations_per_second', 'value': prim['value'],
                'unit': 'problem-iters/s', 'n_gpus': world, 'steps': K, 'warmup': W,
                'ms_per_step': prim['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
                'iters_per_s': prim['iters_per_s'], 'timing': prim['timing'],
                'roofline': prim['roofline'], 'gpu_launches': prim['launches'],
                'results_finite': prim['finite'],
            }
            for key in ('e2e', 'e2e_pinned', 'tf32x3'):
                if key in prim:
                    line[key] = prim[key]
            if 'fista_overcomplete' in out:
                line['fista_overcomplete'] = out['fista_overcomplete']
            if cpu_f is not None:
                line['cpu_baseline'] = cpu_f
        else:
            first = out.get('nmf') or out.get('nmf_strong') or {}
            line = dict(first)
            line.update({'n_gpus': world, 'higher_is_better': True, 'vs_baseline': None, 'dtype': 'f64',
                         'data': 'synthetic', 'gpu_launches': first.get('launches', 0)})
        if 'nmf' in out and 'fista' in out:
            line['secondary'] = out['nmf']
        if 'nmf_strong' in out:
            line['secondary_strong'] = out['nmf_strong']
        for key in ('secondary', 'secondary_strong'):
            if key in line and cpu_n is not None:
                line[key]['cpu_baseline'] = cpu_n
            if key in line and 'nmf_e2e' in out:
                line[key]['e2e'] = out['nmf_e2e']
        if 'extra_configs' in out:
            line['extra_configs'] = out['extra_configs']
        if 'parity' in out:
            line['parity_multi_gpu'] = out['parity']
        if 'errors' in out:
            line['leg_errors'] = out['errors']
        if clocks is None and leg_clocks:                      # no primary leg in this run
            clocks = merge_clocks(None, list(leg_clocks.values())[0])
        line['clocks'] = clocks
        line['clocks_other_legs'] = leg_clocks
        print(json.dumps(line))
    if world > 1:
        from decomp_b200 import comm
        comm.destroy_all()
        dist.destroy_process_group()


def parity_check(np, torch, dist, world, rank, group, device):
    """Small solves with the sample axis sharded over the ranks (all ranks at once when world > 1), compared with the
    numpy oracle on the same seeded inputs; relative max-norm errors and iteration counts. Test infrastructure."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import golden_cases as gc
    from decomp_b200 import dictionary_learning, lasso, nmf
    from oracle import decomp_oracle as orc

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    def shard(a):
        n = a.shape[0]
        return a[rank * n // world:(rank + 1) * n // world]

    res = {'world': world, 'tolerance': 1e-10}
    worst = 0.0
    y, D0, mask = gc._nmf_data(1501, 130, 24, 5)
    for name, m in (('nmf', None), ('nmf_mask', mask)):
        it, D, x = nmf.solve(shard(y), D0.copy(), tol=1e-4, maxiter=400, mask=None if m is None else shard(m), group=group)
        it0, D_ref, x_ref = orc.nmf_mu(y, D0.copy(), tol=1e-4, maxiter=400, mask=m)
        res[name] = {'it': it, 'it_oracle': it0, 'err_D': rel(D, D_ref), 'err_x': rel(x, shard(x_ref))}
        worst = max(worst, res[name]['err_D'], res[name]['err_x'], float(it != it0))
    A, yl, maskl, _ = gc._lasso_data((901,), 24, 40, 3)
    for name, m in (('lasso', None), ('lasso_mask', maskl)):
        it, x = lasso.solve_fastpath(shard(yl), A, 0.05, None, 1e-6, 1000, 'fista', None,
                                     mask=None if m is None else shard(m), group=group)
        it0, x_ref = orc.lasso(yl, A, 0.05, tol=1e-6, method='fista', maxiter=1000, mask=m)
        res[name] = {'it': it, 'it_oracle': it0, 'err_x': rel(x, shard(x_ref))}
        worst = max(worst, res[name]['err_x'], float(it != it0))
    for cplx in (False, True):
        yd, Dd, md = gc._dl_data(230, 33, 12, 9, cplx)
        for masked in (False, True):
            yy = yd * md if masked else yd
            kw = dict(tol=0.0, minibatch=63, maxiter=3, lasso_method='fista', lasso_iter=10, lasso_tol=1.0e-5,
                      mask=md if masked else None, random_seed=4)
            kw_local = dict(kw, mask=shard(md) if masked else None)
            it, D, x = dictionary_learning.solve(shard(yy), Dd.copy(), 0.05, group=group, **kw_local)
            it0, D_ref, x_ref = orc.dictionary_learning(yy, Dd.copy(), 0.05, **kw)
            name = 'dl_%s_%s' % ('c128' if cplx else 'f64', 'mask' if masked else 'nomask')
            res[name] = {'it': it, 'it_oracle': it0, 'err_D': rel(D, D_ref), 'err_x': rel(x, shard(x_ref))}
            worst = max(worst, res[name]['err_D'], res[name]['err_x'], float(it != it0))
    if world > 1:
        t = torch.tensor([worst], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = float(t.item())
    res['worst_error_over_ranks'] = worst
    res['pass'] = bool(worst < 1e-10)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--repeats', type=int, default=10, help='timed regions of --steps iterations (median reported)')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--legs', default='all', help='comma list of ' + ','.join(ALL_LEGS) + ' (default all)')
    ap.add_argument('--skip', default='', help='comma list of legs to leave out')
    ap.add_argument('--batch', type=int, default=FISTA['batch'], help='FISTA problems per GPU')
    ap.add_argument('--rows', type=int, default=NMF['n'], help='NMF rows per GPU (weak-scaling leg)')
    ap.add_argument('--strong-rows', type=int, default=NMF['n'], help='NMF rows in total (strong-scaling leg)')
    ap.add_argument('--e2e-nmf-rows', type=int, default=E2E_NMF_ROWS)
    ap.add_argument('--c5-rows', type=int, default=C5['n'], help='rows per GPU of the configs[4] legs')
    ap.add_argument('--nmf-steps', type=int, default=4)
    args = ap.parse_args()
    legs = list(ALL_LEGS) if args.legs == 'all' else [s for s in args.legs.split(',') if s]
    args.legs = [s for s in legs if s not in set(args.skip.split(','))]
    for s in args.legs:
        if s not in ALL_LEGS:
            raise SystemExit('unknown leg ' + s)
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
