"""Does a kernel stream start after the FIRST of several queued pinned H2D copies on another stream?"""
import time, torch
dev = torch.device('cuda', 0)
n = 4
hs = [torch.empty(40_000_000, dtype=torch.float64, pin_memory=True) for _ in range(n)]   # 320 MB each
ds = [torch.empty(40_000_000, dtype=torch.float64, device=dev) for _ in range(n)]
w = torch.zeros(1 << 20, dtype=torch.float64, device=dev)
E = lambda: torch.cuda.Event(enable_timing=True)
def run(label, cur):
    up = torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(cur):
        start = E(); start.record(cur)
        up.wait_stream(cur)
        evs = []
        with torch.cuda.stream(up):
            for h, d in zip(hs, ds):
                d.copy_(h, non_blocking=True); e = E(); e.record(up); evs.append(e)
        marks = []
        for e in evs:
            cur.wait_event(e); w.add_(1.0); m = E(); m.record(cur); marks.append(m)
    torch.cuda.synchronize()
    print(label, 'copies done at', ['%.1f' % start.elapsed_time(e) for e in evs],
          '| kernels after wait at', ['%.1f' % start.elapsed_time(m) for m in marks])
run('default stream ', torch.cuda.current_stream(dev))
run('default stream ', torch.cuda.current_stream(dev))
run('side stream    ', torch.cuda.Stream(dev))
run('side stream    ', torch.cuda.Stream(dev))
