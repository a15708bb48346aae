"""Multi-process paths: world_size 2 over gloo on the CPU (host logic and exchange scheme), and world_size 2 over
NCCL on two GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_dist.py -m gpu`)."""
import random

import pytest
import torch
import torch.multiprocessing as mp

import dist_workers


def _spawn(fn, world):
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + random.randint(0, 2000)
    mp.spawn(fn, args=(world, port, out), nprocs=world, join=True)
    return dict(out)


def test_world_size_2_gloo_host_logic():
    out = _spawn(dist_workers.cpu_host_logic, 2)
    assert set(out) == {0, 1}
    for rank, res in out.items():
        assert res['allreduce2d'] and res['latch'], (rank, res)
        assert res['nmf_D'] < 1e-12 and res['nmf_x'] < 1e-12, (rank, res)


def test_world_size_2_gloo_dictionary_learning_row_sharding():
    out = _spawn(dist_workers.cpu_dl_row_sharding, 2)
    assert set(out) == {0, 1} and all(res['ok'] for res in out.values()), out


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_world_size_2_comm_wrappers():
    out = _spawn(dist_workers.gpu_comm_wrappers, 2)
    assert set(out) == {0, 1}
    for rank, res in out.items():
        assert all(res.values()), (rank, res)


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_world_size_2_nccl_sharded_solves():
    out = _spawn(dist_workers.gpu_sharded_solves, 2)
    assert set(out) == {0, 1}
    for rank, res in out.items():
        for name in ('nmf', 'nmf_mask'):
            it, it0, eD, ex = res[name]
            assert it == it0 and eD < 1e-10 and ex < 1e-10, (rank, name, res[name])
        for name in ('lasso', 'lasso_mask'):
            it, it0, ex = res[name]
            assert it == it0 and ex < 1e-10, (rank, name, res[name])
        # TF32-split paths: tolerance of tests/test_tf32x3_gpu.py
        for name in ('nmf_tf32', 'nmf_mask_tf32'):
            it, it0, eD, ex = res[name]
            assert it == it0 and eD < 1e-4 and ex < 1e-4, (rank, name, res[name])
        it, it0, ex = res['lasso_mask_tf32']
        assert abs(it - it0) <= 10 and ex < 1e-3, (rank, res['lasso_mask_tf32'])


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_world_size_2_nccl_dictionary_learning():
    out = _spawn(dist_workers.gpu_sharded_dictionary_learning, 2)
    assert set(out) == {0, 1}
    for rank, res in out.items():
        for name, (it, it0, eD, ex) in res.items():
            assert it == it0 and eD < 1e-10 and ex < 1e-10, (rank, name, it, it0, eD, ex)
