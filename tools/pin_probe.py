"""Upload of a 0.8 GB pageable numpy array: torch's staged copy vs cudaHostRegister in place + asynchronous copy."""
import time, numpy as np, torch
dev = torch.device('cuda', 0)
B, f = 100000, 1024
y = np.random.RandomState(0).randn(B, f)
d = torch.empty((B, f), dtype=torch.float64, device=dev)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); d.copy_(torch.from_numpy(y), non_blocking=True); t1 = T()
    print('pageable H2D 819 MB: %.1f ms = %.1f GB/s' % ((t1 - t0) * 1e3, 0.8192 / (t1 - t0)))
rt = torch.cuda.cudart()
for rep in range(3):
    t0 = T(); rc = rt.cudaHostRegister(y.ctypes.data, y.nbytes, 0); t1 = T()
    d.copy_(torch.from_numpy(y), non_blocking=True); t2 = T()
    rt.cudaHostUnregister(y.ctypes.data); t3 = T()
    print('register %.1f ms (rc %s) + copy %.1f ms + unregister %.1f ms' % ((t1 - t0) * 1e3, rc, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
# chunked register (pipelining-friendly): 8 chunks
rows = B // 8
t0 = T()
for c in range(8):
    a = y[c * rows:(c + 1) * rows]
    rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
    d[c * rows:(c + 1) * rows].copy_(torch.from_numpy(a), non_blocking=True)
t1 = T()
for c in range(8):
    rt.cudaHostUnregister(y[c * rows:(c + 1) * rows].ctypes.data)
t2 = T()
print('8 chunks register+copy %.1f ms, unregister %.1f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
