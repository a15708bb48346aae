"""BASELINE.json configs[0]: NMF-MU, Y 1000x200 float64, k=20, 100 sweeps -- GPU (host arrays in/out) vs the numpy oracle.
Also separates the fixed cost of a call (copies, allocations) from the cost per sweep."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import golden_cases as gc
from decomp_b200 import nmf
from oracle import decomp_oracle as orc
y, D0, mask = gc._nmf_data(1000, 200, 20, 0, 'l2', reference_order=False)
res = {}
for name, m in (('nomask', None), ('mask', mask)):
    nmf.solve(y, D0.copy(), tol=0.0, maxiter=11, mask=m)
    torch.cuda.synchronize()
    times = {}
    for sweeps in (1, 100, 1000):
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter(); it, D, x = nmf.solve(y, D0.copy(), tol=0.0, maxiter=sweeps + 1, mask=m); best = min(best, time.perf_counter() - t0)
        times[sweeps] = best
    it, D, x = nmf.solve(y, D0.copy(), tol=0.0, maxiter=101, mask=m)
    t0 = time.perf_counter(); it0, D_ref, x_ref = orc.nmf_mu(y, D0.copy(), tol=0.0, maxiter=101, mask=m); cpu = time.perf_counter() - t0
    res[name] = dict(gpu_s=times[100], gpu_sweeps_per_s=100 / times[100], call_1_sweep_s=times[1], call_1000_sweeps_s=times[1000],
                     us_per_sweep=(times[1000] - times[100]) / 900 * 1e6, cpu_s=cpu, cpu_sweeps_per_s=100 / cpu,
                     rel_err_D=float(np.max(np.abs(D - D_ref)) / np.max(np.abs(D_ref))))
print(json.dumps(res, indent=1))
