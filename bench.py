#!/usr/bin/env python
"""Benchmark of the deComP hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload both|fista|nmf]

Primary workload (BASELINE.json configs[1]): batched FISTA Lasso, 100 000 problems per GPU, A (256, 1024),
alpha = 0.1, float64, tol = 0 (fixed iteration count).  One *step* is one FISTA iteration over the whole batch;
the iterate-resident kernel runs up to 32 of them per launch (``gpu_launches`` counts the launches, the roofline is
per launch).  ``value`` is problem-iterations per second with everything resident in HBM; ``e2e`` is the same metric
through ``decomp_b200.lasso.solve`` with pinned HOST arrays in and a host array out (H2D of y and A, D2H of x inside
the timed region; the call overlaps them with the iterations chunk by chunk).

Secondary workload (configs[2], reported under "secondary"): NMF multiplicative update, 1 000 000 rows per GPU
x 4096 features, k = 256, float64; one step is one full sweep (x update + statistics + all-reduce + D update).

Multi-GPU (torchrun, one rank per GPU): the sample axis is sharded, every rank holds the same number of rows
(weak scaling).  FISTA needs no data-path collective; NMF all-reduces the [k, f] and [k, k] statistics per
sweep over NCCL.  Timing: CUDA events on the launching stream between barriers, max over ranks.

``--impl reference`` times the CPU restatement of the reference's numpy algorithm (oracle/decomp_oracle.py,
kind "port": the reference itself is a Python package that does not travel to the GPU box) on the host cores.
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
CPU_FISTA_BATCH = 24576         # bounded CPU sample (problems): ~10 s per 100 iterations on 16 cores
L2_BYTES = 126 * 2 ** 20


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
    out = dict(a)
    if b.get('sm_mhz') is not None and (a.get('sm_mhz') is None or b['sm_mhz'] < a['sm_mhz']):
        out['sm_mhz'] = b['sm_mhz']
    out['reasons'] = sorted(set(a.get('reasons', [])) | set(b.get('reasons', [])))
    out['samples'] = a.get('samples', 0) + b.get('samples', 0)
    return out


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


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
    xt = rng.randn(B, k) * np.rint(rng.uniform(size=(B, k)))
    y = xt.dot(A) + 0.1 * rng.randn(B, f)
    return y, A


def nmf_data_device(torch, n, f, k, seed, device):
    """Ct, Dt = max(N(0,1), 0); Y = Ct Dt + 0.1 N(0,1); D0 = max(Dt + 0.3 N, 0.1) replicated."""
    g = torch.Generator(device=device)
    g.manual_seed(0)
    Dt = torch.randn((k, f), dtype=torch.float64, device=device, generator=g).clamp_(min=0.0)
    D0 = (Dt + 0.3 * torch.randn((k, f), dtype=torch.float64, device=device, generator=g)).clamp_(min=0.1)
    g.manual_seed(2000 + seed)
    y = torch.empty((n, f), dtype=torch.float64, device=device)
    chunk = 32768
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        ct = torch.randn((r1 - r0, k), dtype=torch.float64, device=device, generator=g).clamp_(min=0.0)
        y[r0:r1] = ct @ Dt
        y[r0:r1] += 0.1 * torch.randn((r1 - r0, f), dtype=torch.float64, device=device, generator=g)
    return y, D0


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_fista(np, steps, warmup, batch=CPU_FISTA_BATCH):
    """The oracle port of decomp.lasso (fista, tol=0) on the host cores; returns problem-iterations/s."""
    from oracle import decomp_oracle as orc
    k, f, alpha = FISTA['k'], FISTA['f'], FISTA['alpha']
    y, A = fista_data_host(np, batch, k, f, 0)
    if warmup > 0:
        orc.lasso(y[:1024], A, alpha, tol=0.0, method='fista', maxiter=max(1, min(warmup, 5)))
    t0 = time.perf_counter()
    orc.lasso(y, A, alpha, tol=0.0, method='fista', maxiter=steps)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') == 'blas']
        return max(n) if n else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import numpy as np
    steps = max(1, args.steps)
    batch = CPU_FISTA_BATCH
    value, dt = cpu_fista(np, steps, args.warmup, batch)
    line = {
        'impl': 'reference',
        'metric': 'batched_fista_problem_iterations_per_second', 'value': value, 'unit': 'problem-iters/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup, 'ms_per_step': dt / steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'batched FISTA Lasso, A (256,1024), alpha=0.1, float64, tol=0 '
                               '(BASELINE.json configs[1]); CPU sample of %d problems' % batch},
        'cpu_baseline': {'value': value, 'unit': 'problem-iters/s', 'cores': blas_threads(), 'kind': 'port',
                         'sample': '%d problems x %d FISTA iterations, numpy/OpenBLAS via oracle/decomp_oracle.py '
                                   '(set-up GEMMs included)' % (batch, steps)},
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

    from decomp_b200 import lasso, nmf, ops

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn):
        """fn enqueues work on the current stream; returns (device ms max over ranks, launches, clocks)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        before = ops.LAUNCHES
        with ClockSampler(local) as cs:
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = ops.LAUNCHES - before
        barrier()
        return max_over_ranks(ms), launches, cs.summary()

    hbm_peak, hbm_src = load_peaks()
    dmma_peak = ops.probe_dmma_tflops()
    K, W = max(1, args.steps), max(0, args.warmup)
    out = {}
    clocks = None

    # ---------------------------------------------------------------- batched FISTA (primary)
    if args.workload in ('both', 'fista'):
        B, k, f, alpha = args.batch, FISTA['k'], FISTA['f'], FISTA['alpha']
        y, A = fista_data_device(torch, B, k, f, rank, device)
        solver = lasso.LassoSolver(y, A, alpha, None, 0.0, W + K, 'fista', False)
        solver.iterate(0, W)
        ms, launches, clocks = timed(lambda: solver.iterate(W, W + K))
        state = solver.finish()
        xs = state.result
        finite = bool(torch.isfinite(xs).all().item())
        nnz = float((xs != 0).double().mean().item())
        del solver, state, xs
        sec = ms * 1e-3
        per_launch = K / float(max(launches, 1))            # iterations one launch of the resident kernel runs
        flops_launch = 2.0 * B * k * k * per_launch         # 2 B k^2 per iteration
        bytes_launch = 5.0 * B * k * 8                      # read yAh, w, x_prev; write x_new, w_next: once per launch
        t_launch = sec / max(launches, 1)
        out['fista'] = {
            'value': B * world * K / sec, 'iters_per_s': K / sec, 'ms_per_step': ms / K, 'launches': launches,
            'finite': finite, 'nonzero_fraction': nnz,
            'roofline': {'bound': 'tensor', 'achieved': flops_launch / t_launch / 1e12, 'peak': dmma_peak,
                         'unit': 'TFLOP/s', 'frac': flops_launch / t_launch / 1e12 / dmma_peak,
                         'traffic': args.traffic_fista,
                         'kernel': 'lasso_resident_kernel (iterate resident in shared memory, Q streamed by TMA, '
                                   'DMMA GEMM + ISTA/FISTA update, up to 32 iterations per launch)',
                         'iterations_per_launch': per_launch,
                         'peak_source': 'FP64 tensor (DMMA.8x8x4) issue rate measured live by '
                                        'decomp_probe_dmma_tflops(); MEASURED_PEAKS.json has no FP64 entry',
                         'algorithmic_flops_per_launch': flops_launch,
                         'hbm': {'algorithmic_bytes_per_launch': bytes_launch,
                                 'achieved_gbs': bytes_launch / t_launch / 1e9, 'peak_gbs': hbm_peak,
                                 'frac': bytes_launch / t_launch / 1e9 / hbm_peak, 'peak_source': hbm_src}},
        }
        # ---- the TF32-split (tcgen05) variant of the same iterations: reported beside, never as the headline
        if not args.no_tf32:
            solver = lasso.LassoSolver(y, A, alpha, None, 0.0, W + K, 'fista', False, precision='tf32x3')
            solver.iterate(0, W)
            ms32, launches32, c32 = timed(lambda: solver.iterate(W, W + K))
            clocks = merge_clocks(clocks, c32)
            x32 = solver.finish().result
            ref = lasso.LassoSolver(y, A, alpha, None, 0.0, W + K, 'fista', False)
            ref.iterate(0, W + K)
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
                             'frac': bytes32 / t32 / 1e9 / hbm_peak, 'traffic': None,
                             'kernel': 'tf32x3_gemm_kernel (tcgen05.mma kind::tf32, TMEM) + proxq_apply_kernel',
                             'algorithmic_bytes_per_iteration': bytes32, 'peak_source': hbm_src},
            }
        # ---- end to end through the public API with pinned host buffers
        yh = torch.empty((B, f), dtype=torch.float64, pin_memory=True)
        Ah = torch.empty((k, f), dtype=torch.float64, pin_memory=True)
        yh.copy_(y)
        Ah.copy_(A)
        del y
        torch.cuda.synchronize()
        y_np, A_np = yh.numpy(), Ah.numpy()
        lasso.solve(y_np[:4096], A_np, alpha, tol=0.0, method='fista', maxiter=3)       # allocator / module warm-up
        e2e_ms = []
        x_np = None
        for _ in range(3):
            x_np = None        # drop the previous result: its page-locked block returns to torch's host cache
            barrier()
            t0 = time.perf_counter()
            it, x_np = lasso.solve(y_np, A_np, alpha, tol=0.0, method='fista', maxiter=K)
            torch.cuda.synchronize()
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
        e2e = max_over_ranks(min(e2e_ms)) * 1e-3
        assert it == K - 1 and x_np.shape == (B, k)
        out['fista']['e2e'] = {'value': B * world * K / e2e, 'unit': 'problem-iters/s',
                               'h2d_bytes_per_step': (y_np.nbytes + A_np.nbytes) / K,
                               'd2h_bytes_per_step': x_np.nbytes / K, 'ms_per_call': e2e * 1e3,
                               'note': 'one lasso.solve(maxiter=steps) call with host arrays: H2D of y and A, '
                                       'set-up GEMMs, steps iterations, D2H of x, all inside the timed region '
                                       '(tol = 0: the call runs the batch in row chunks so that copies overlap the '
                                       'iterations); bytes are per call / steps'}
        del yh, Ah, y_np, A_np, x_np
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- NMF-MU (secondary)
    if args.workload in ('both', 'nmf'):
        n, f, k = args.rows, NMF['f'], NMF['k']
        Kn = max(2, args.nmf_steps)
        Wn = 3
        y, D0 = nmf_data_device(torch, n, f, k, rank, device)
        X = torch.ones((n, k), dtype=torch.float64, device=device)
        solver = nmf.MuSolver(y, D0, X, 0.0, group=group)
        for it in range(1, Wn + 1):
            solver.sweep(it)
        ms, launches, c2 = timed(lambda: [solver.sweep(it) for it in range(Wn + 1, Wn + Kn + 1)])
        clocks = merge_clocks(clocks, c2)
        D = solver.Dbuf[(Wn + Kn) % 2]
        finite = bool(torch.isfinite(D).all().item() and torch.isfinite(X).all().item())
        sec = ms * 1e-3
        flops_sweep = 4.0 * n * k * f + 4.0 * n * k * k + 4.0 * k * k * f
        bytes_sweep = 2.0 * n * f * 8 + 6.0 * n * k * 8                        # Y twice; X, NEG r/w
        t_sweep = sec / Kn
        out['nmf'] = {
            'metric': 'nmf_mu_row_iterations_per_second', 'value': n * world * Kn / sec, 'unit': 'row-iters/s',
            'iters_per_s': Kn / sec, 'ms_per_step': ms / Kn, 'steps': Kn, 'warmup': Wn, 'launches': launches,
            'finite': finite,
            'config': {'workload': 'NMF-MU l2, %d rows per GPU x %d features, k=%d, float64, tol=0 '
                                   '(BASELINE.json configs[2], weak scaling)' % (n, f, k)},
            'roofline': {'bound': 'tensor', 'achieved': flops_sweep / t_sweep / 1e12, 'peak': dmma_peak,
                         'unit': 'TFLOP/s', 'frac': flops_sweep / t_sweep / 1e12 / dmma_peak, 'traffic': None,
                         'kernel': 'whole sweep: gemm_f64_kernel<NT, MU_NUM> (Y D^T) + gemm_f64_kernel<TN> (X^T Y) '
                                   'dominate with 2nkf flop each',
                         'algorithmic_flops_per_sweep': flops_sweep,
                         'hbm': {'algorithmic_bytes_per_sweep': bytes_sweep,
                                 'achieved_gbs': bytes_sweep / t_sweep / 1e9, 'peak_gbs': hbm_peak}},
        }
        del solver, y, X, D0, D
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt = cpu_fista(np, 100, 1)
        cpu = {'value': v, 'unit': 'problem-iters/s', 'cores': blas_threads(), 'kind': 'port',
               'sample': '%d problems x 100 FISTA iterations (%.1f s), numpy/OpenBLAS via oracle/decomp_oracle.py'
                         % (CPU_FISTA_BATCH, dt)}

    if rank == 0:
        prim = out.get('fista') or out.get('nmf')
        if 'fista' in out:
            line = {
                'metric': 'batched_fista_problem_iterations_per_second', 'value': prim['value'],
                'unit': 'problem-iters/s', 'n_gpus': world, 'steps': K, 'warmup': W,
                'ms_per_step': prim['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': 'batched FISTA Lasso: %d problems per GPU, A (256,1024), alpha=0.1, float64, '
                                       'tol=0 (BASELINE.json configs[1])' % args.batch,
                           'l2': 'working set per iteration 5*B*k*8 = %.0f MB > %d MB L2 (inputs larger than L2, '
                                 'no flush)' % (5.0 * args.batch * 256 * 8 / 1e6, L2_BYTES // 2 ** 20),
                           'parallelism': 'sample axis sharded over %d GPU(s), no data-path collective' % world},
                'iters_per_s': prim['iters_per_s'],
                'roofline': prim['roofline'], 'e2e': prim['e2e'], 'gpu_launches': prim['launches'],
                'results_finite': prim['finite'],
            }
            if 'nmf' in out:
                line['secondary'] = out['nmf']
            if 'tf32x3' in prim:
                line['tf32x3'] = prim['tf32x3']
        else:
            line = dict(out['nmf'])
            line.update({'n_gpus': world, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                         'dtype': 'f64', 'data': 'synthetic', 'gpu_launches': prim['launches']})
        if cpu is not None:
            line['cpu_baseline'] = cpu
        line['clocks'] = clocks
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='both', choices=['both', 'fista', 'nmf'])
    ap.add_argument('--batch', type=int, default=FISTA['batch'], help='FISTA problems per GPU')
    ap.add_argument('--rows', type=int, default=NMF['n'], help='NMF rows per GPU')
    ap.add_argument('--nmf-steps', type=int, default=5)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-tf32', action='store_true')
    args = ap.parse_args()
    # dram bytes per launch of the dominant kernel, copied from the committed ncu --set full capture
    args.traffic_fista = None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as fh:
            args.traffic_fista = json.load(fh).get('fista_prox_dram_bytes_per_launch')
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
