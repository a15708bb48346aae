"""Worker functions for the multi-process tests (spawned by torch.multiprocessing)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def _init(rank, world, port, backend):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    if backend == 'nccl':
        torch.cuda.set_device(rank)
        dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    else:
        dist.init_process_group(backend, rank=rank, world_size=world)


def shard(a, rank, world):
    """Contiguous row block of rank `rank` (the last rank takes the remainder)."""
    n = a.shape[0]
    per = n // world
    lo = rank * per
    hi = n if rank == world - 1 else lo + per
    return a[lo:hi]


# --------------------------------------------------------------------------------------- CPU / gloo
def cpu_host_logic(rank, world, port, out):
    """(1) padded-statistic all-reduce helper, (2) MIN-all-reduce convergence latch, (3) the exchange scheme of the
    sharded NMF sweep (local x update, all-reduced x^T y and x^T x, replicated D update) equals the single-process
    oracle."""
    _init(rank, world, port, 'gloo')
    from decomp_b200 import nmf
    import golden_cases as gc
    from oracle import decomp_oracle as orc
    res = {}

    base = torch.arange(24, dtype=torch.float64).reshape(4, 6) * (rank + 1)
    view = base[:, :5]                                     # row-padded (non-contiguous) statistic
    nmf._allreduce2d(view, dist.group.WORLD)
    expect = torch.arange(24, dtype=torch.float64).reshape(4, 6)[:, :5] * sum(range(1, world + 1))
    res['allreduce2d'] = bool(torch.equal(view, expect)) and float(base[0, 5]) == 5.0 * (rank + 1)

    latch = torch.tensor([7 if rank == 0 else 0], dtype=torch.int32)
    dist.all_reduce(latch, op=dist.ReduceOp.MIN)
    both = torch.tensor([7], dtype=torch.int32)
    dist.all_reduce(both, op=dist.ReduceOp.MIN)
    res['latch'] = int(latch.item()) == 0 and int(both.item()) == 7

    y, D0, _ = gc._nmf_data(203, 31, 5, 2)
    yl = shard(y, rank, world)
    D = orc.unit_rows(D0)
    x = np.ones((yl.shape[0], D.shape[0]))
    for _ in range(15):
        G = D.dot(D.T)
        x = x * np.maximum(yl.dot(D.T), 0.0) / np.maximum(x.dot(G), orc.EPS)
        T = torch.from_numpy(x.T.dot(yl))
        S = torch.from_numpy(x.T.dot(x))
        dist.all_reduce(T)
        dist.all_reduce(S)
        D = orc.unit_rows(D * np.maximum(T.numpy(), 0.0) / np.maximum(S.numpy().dot(D), orc.EPS))
    _, D_ref, x_ref = orc.nmf_mu(y, D0.copy(), tol=0.0, maxiter=16)
    res['nmf_D'] = float(np.max(np.abs(D - D_ref)) / np.max(np.abs(D_ref)))
    res['nmf_x'] = float(np.max(np.abs(x - shard(x_ref, rank, world))) / np.max(np.abs(x_ref)))
    out[rank] = res
    dist.destroy_process_group()


def gpu_sharded_dictionary_learning(rank, world, port, out):
    """Dictionary learning on row-sharded inputs (each rank holds a block of rows; masked: statistics reduce-scattered
    along f) vs the single-process oracle on the concatenated rows.  minibatch 20 of 230 rows over 2 ranks: some
    minibatches leave a rank few rows; f = 33 is not a multiple of the rank count (ragged channel slabs)."""
    _init(rank, world, port, 'nccl')
    from decomp_b200 import dictionary_learning
    import golden_cases as gc
    from oracle import decomp_oracle as orc
    res = {}

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    for cplx in (False, True):
        y, D0, mask = gc._dl_data(230, 33, 12, 9, cplx)
        for masked in (False, True):
            for mb, tol in ((63, 0.0), (20, 1e-3), (150, 0.0)):      # 150: more than the 115 rows a rank holds
                yy = y * mask if masked else y
                kw = dict(tol=tol, minibatch=mb, maxiter=3, lasso_method='fista', lasso_iter=10, lasso_tol=1.0e-5,
                          random_seed=4)
                it, D, x = dictionary_learning.solve(shard(yy, rank, world), D0.copy(), 0.05, group=dist.group.WORLD,
                                                     mask=shard(mask, rank, world) if masked else None, **kw)
                it0, D_ref, x_ref = orc.dictionary_learning(yy, D0.copy(), 0.05, mask=mask if masked else None, **kw)
                res['dl_%s_%s_mb%d' % ('c' if cplx else 'f', 'mask' if masked else 'nomask', mb)] = (
                    it, it0, rel(D, D_ref), rel(x, shard(x_ref, rank, world)))
    from decomp_b200 import comm
    comm.destroy_all()
    out[rank] = res
    dist.destroy_process_group()


def cpu_dl_row_sharding(rank, world, port, out):
    """Host logic of the row-sharded dictionary learning over gloo: the row layout exchanged between the ranks and the
    per-epoch selections (every minibatch is partitioned over the ranks, order preserved)."""
    _init(rank, world, port, 'gloo')
    from decomp_b200 import dictionary_learning as dl
    n_loc = 7 if rank == 0 else 5
    n, row0, w, r = dl._row_layout(n_loc, dist.group.WORLD, torch.device('cpu'))
    rng = np.random.RandomState(3)
    order = np.arange(n)
    index = np.arange(n)
    rng.shuffle(index)
    order = order[index]
    mb, steps = 5, n // 5
    sel, bounds = dl._epoch_selection(order, steps, mb, row0, n_loc)
    mine = [(sel[bounds[i]:bounds[i + 1]] + row0).tolist() for i in range(steps)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = (n, w, r) == (12, world, rank) and row0 == (0 if rank == 0 else 7)
    for i in range(steps):
        want = order[i * mb:(i + 1) * mb].tolist()
        got = sorted(sum((g[i] for g in gathered), []))
        ok = ok and got == sorted(want)
        ok = ok and mine[i] == [v for v in want if row0 <= v < row0 + n_loc]          # minibatch order kept
    out[rank] = {'ok': bool(ok)}
    dist.destroy_process_group()


# --------------------------------------------------------------------------------------- GPU / nccl
def gpu_comm_wrappers(rank, world, port, out):
    """decomp_comm_* (the C ABI's NCCL wrappers) against the expected sums / slabs."""
    _init(rank, world, port, 'nccl')
    from decomp_b200 import comm
    dev = torch.device('cuda', rank)
    res = {}
    g = dist.group.WORLD
    t = torch.arange(12, dtype=torch.float64, device=dev).reshape(3, 4) * (rank + 1)
    comm.all_reduce_sum(t, g)
    res['sum'] = bool(torch.equal(t.cpu(), torch.arange(12, dtype=torch.float64).reshape(3, 4) * sum(range(1, world + 1))))
    padded = (torch.ones((4, 6), dtype=torch.float64, device=dev) * (rank + 1))[:, :5]      # row-padded view
    comm.all_reduce_sum(padded, g)
    res['sum_padded'] = bool((padded.cpu() == float(sum(range(1, world + 1)))).all())
    latch = torch.tensor([5 if rank == 0 else 0], dtype=torch.int32, device=dev)
    comm.all_reduce_min_i32(latch, g)
    res['min'] = int(latch.item()) == 0
    inp = torch.stack([torch.full((2, 3), float(10 * d + rank), dtype=torch.float64, device=dev) for d in range(world)])
    slab = torch.empty((2, 3), dtype=torch.float64, device=dev)
    comm.reduce_scatter_sum(slab, inp, g)
    res['reduce_scatter'] = bool((slab.cpu() == float(sum(10 * rank + r for r in range(world)))).all())
    gathered = torch.empty((world, 2, 3), dtype=torch.float64, device=dev)
    comm.all_gather(gathered, slab, g)
    want = torch.stack([torch.full((2, 3), float(sum(10 * d + r for r in range(world))), dtype=torch.float64)
                        for d in range(world)])
    res['all_gather'] = bool(torch.equal(gathered.cpu(), want))
    comm.destroy_all()
    out[rank] = res
    dist.destroy_process_group()


def gpu_sharded_solves(rank, world, port, out):
    """Sharded NMF (unmasked, masked) and sharded Lasso (tol > 0: global convergence decision) against the oracle."""
    _init(rank, world, port, 'nccl')
    from decomp_b200 import lasso, nmf
    import golden_cases as gc
    from oracle import decomp_oracle as orc
    res = {}

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    y, D0, mask = gc._nmf_data(1501, 130, 24, 5)
    for name, m in (('nmf', None), ('nmf_mask', mask)):
        it, D, x = nmf.solve(shard(y, rank, world), D0.copy(), tol=1e-4, maxiter=400,
                             mask=None if m is None else shard(m, rank, world), group=dist.group.WORLD)
        it0, D_ref, x_ref = orc.nmf_mu(y, D0.copy(), tol=1e-4, maxiter=400, mask=m)
        res[name] = (it, it0, rel(D, D_ref), rel(x, shard(x_ref, rank, world)))

    A, yl, maskl, _ = gc._lasso_data((901,), 24, 40, 3)
    for name, m in (('lasso', None), ('lasso_mask', maskl)):
        it, x = lasso.solve_fastpath(shard(yl, rank, world), A, 0.05, None, 1e-6, 1000, 'fista', None,
                                     mask=None if m is None else shard(m, rank, world), group=dist.group.WORLD)
        it0, x_ref = orc.lasso(yl, A, 0.05, tol=1e-6, method='fista', maxiter=1000, mask=m)
        res[name] = (it, it0, rel(x, shard(x_ref, rank, world)))
    # the TF32-split paths sharded the same way (statistics all-reduced in FP64): against the FP64 oracle at the
    # TF32 tolerance; k a multiple of 32
    y2, D2, mask2 = gc._nmf_data(2001, 260, 32, 7)
    for name, m in (('nmf_tf32', None), ('nmf_mask_tf32', mask2)):
        it, D, x = nmf.solve(shard(y2, rank, world), D2.copy(), tol=0.0, maxiter=16,
                             mask=None if m is None else shard(m, rank, world), group=dist.group.WORLD,
                             precision='tf32x3')
        it0, D_ref, x_ref = orc.nmf_mu(y2, D2.copy(), tol=0.0, maxiter=16, mask=m)
        res[name] = (it, it0, rel(D, D_ref), rel(x, shard(x_ref, rank, world)))
    A2, yl2, maskl2, _ = gc._lasso_data((901,), 32, 40, 3)
    it, x = lasso.solve_fastpath(shard(yl2, rank, world), A2, 0.05, None, 1e-6, 1000, 'fista', None,
                                 mask=shard(maskl2, rank, world), group=dist.group.WORLD, precision='tf32x3')
    it0, x_ref = orc.lasso(yl2, A2, 0.05, tol=1e-6, method='fista', maxiter=1000, mask=maskl2)
    res['lasso_mask_tf32'] = (it, it0, rel(x, shard(x_ref, rank, world)))
    from decomp_b200 import comm
    comm.destroy_all()
    out[rank] = res
    dist.destroy_process_group()
