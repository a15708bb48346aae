"""Timeline of the pipelined host path of lasso.solve (C2): per chunk upload / iterate / download, on one clock."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from decomp_b200 import lasso
from decomp_b200._device import to_device2d
dev = torch.device('cuda', 0)
B, k, f, K = 100000, 256, 1024, 200
y, A = bench.fista_data_device(torch, B, k, f, 0, dev)
yh = torch.empty((B, f), dtype=torch.float64, pin_memory=True); yh.copy_(y)
Ah = torch.empty((k, f), dtype=torch.float64, pin_memory=True); Ah.copy_(A)
y_np, A_np = yh.numpy(), Ah.numpy()
del y
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); it, x = lasso.solve(y_np, A_np, 0.1, tol=0.0, method='fista', maxiter=K); t1 = T()
    print('pipelined solve %.1f ms' % ((t1 - t0) * 1e3)); del x
lasso.PIPELINE_MIN_BYTES = 1 << 60
for rep in range(3):
    t0 = T(); it, x = lasso.solve(y_np, A_np, 0.1, tol=0.0, method='fista', maxiter=K); t1 = T()
    print('one-piece solve %.1f ms' % ((t1 - t0) * 1e3)); del x
lasso.PIPELINE_MIN_BYTES = 64 << 20
# manual timeline
chunks = lasso._row_chunks(B, f, k, dev)
for rep in range(3):
    print(chunks)
    cur = torch.cuda.current_stream(dev); up, down = lasso._copy_streams(dev)
    A2 = to_device2d(A_np, dev, copy=False)
    host = torch.empty((B, k), dtype=torch.float64, pin_memory=True)
    torch.cuda.synchronize()
    E = lambda: torch.cuda.Event(enable_timing=True)
    start = E(); start.record(cur); up.wait_stream(cur); down.wait_stream(cur)
    th0 = time.perf_counter()
    staged = []
    with torch.cuda.stream(up):
        for r0, r1 in chunks:
            yc = to_device2d(y_np[r0:r1], dev, copy=False); ev = E(); ev.record(up); staged.append((yc, ev))
    th1 = time.perf_counter()
    marks = []
    for (r0, r1), (yc, ev) in zip(chunks, staged):
        cur.wait_event(ev); b = E(); b.record(cur)
        h0 = time.perf_counter()
        st = lasso.lasso_device(yc, A2, 0.1, None, 0.0, K, 'fista', False, None)
        h1 = time.perf_counter()
        done = E(); done.record(cur); down.wait_event(done)
        with torch.cuda.stream(down):
            host[r0:r1].copy_(st.result, non_blocking=True); dd = E(); dd.record(down)
        marks.append((ev, b, done, dd, (h0 - th0) * 1e3, (h1 - th0) * 1e3, st))
    down.synchronize(); torch.cuda.synchronize()
    print('host: uploads enqueued by %.1f ms' % ((th1 - th0) * 1e3))
    for (r0, r1), (ev, b, done, dd, h0, h1, st) in zip(chunks, marks):
        print('rows %6d-%6d  upload done %6.1f | compute %6.1f -> %6.1f | download done %6.1f | host enqueue %6.1f -> %6.1f' % (
            r0, r1, start.elapsed_time(ev), start.elapsed_time(b), start.elapsed_time(done), start.elapsed_time(dd), h0, h1))
