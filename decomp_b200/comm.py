"""Collectives of the sharded solves through the C ABI (``decomp_comm_*``: thin NCCL wrappers, include/decomp_b200.h).

``torch.distributed`` stays the plumbing: it tells the ranks about each other and carries the 128-byte NCCL id from
rank 0 of the group to the others, once per group.  The data-path collectives -- the all-reduce of the [k, f] / [k, k]
sufficient statistics, the MIN-all-reduce of the Lasso convergence latch, the reduce-scatter / all-gather of the
masked dictionary update -- then run on OUR communicator, enqueued on the same CUDA stream as the kernels around them
(no hop to a communication stream and back).  CPU tensors (the gloo tests of the host logic) go through
``torch.distributed`` directly.
"""
import ctypes

import torch

from . import _lib

_comms = {}


class Communicator(object):
    def __init__(self, handle, world, rank):
        self.handle, self.world, self.rank = handle, world, rank


def communicator(group):
    """The native communicator of ``group`` (created on first use; a collective call: every rank must get here)."""
    key = id(group)
    if key in _comms:
        return _comms[key]
    dist = torch.distributed
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lib = _lib.lib()
    ident = (ctypes.c_ubyte * 128)()
    if rank == 0:
        _lib.check(lib.decomp_comm_unique_id(ident), 'decomp_comm_unique_id')
    carrier = torch.tensor(list(ident), dtype=torch.uint8, device=torch.device('cuda', torch.cuda.current_device()))
    dist.broadcast(carrier, src=dist.get_global_rank(group, 0), group=group)
    ident = (ctypes.c_ubyte * 128)(*carrier.cpu().tolist())
    handle = ctypes.c_void_p()
    _lib.check(lib.decomp_comm_init(ident, world, rank, ctypes.byref(handle)), 'decomp_comm_init')
    _comms[key] = Communicator(handle, world, rank)
    return _comms[key]


def destroy_all():
    """Destroys every native communicator (call before ``torch.distributed.destroy_process_group``)."""
    for c in _comms.values():
        _lib.lib().decomp_comm_destroy(c.handle)
    _comms.clear()


def _contiguous(t):
    return t if t.is_contiguous() else t.contiguous()


def all_reduce_sum(t, group):
    """In-place sum over the ranks of a float64 / complex128 tensor (row-padded 2-D views included)."""
    if not t.is_cuda:
        flat = _contiguous(t)
        torch.distributed.all_reduce(flat, group=group)
        if flat is not t:
            t.copy_(flat)
        return t
    assert t.dtype in (torch.float64, torch.complex128)
    flat = _contiguous(t)
    count = flat.numel() * (2 if t.is_complex() else 1)          # complex128: interleaved doubles
    c = communicator(group)
    _lib.check(_lib.lib().decomp_comm_allreduce_sum_f64(c.handle, ctypes.c_void_p(flat.data_ptr()), count,
                                                        _lib.stream_ptr()), 'decomp_comm_allreduce_sum_f64')
    if flat is not t:
        t.copy_(flat)
    return t


def all_reduce_min_i32(t, group):
    """In-place MIN over the ranks of an int32 tensor (the convergence latch)."""
    if not t.is_cuda:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN, group=group)
        return t
    assert t.dtype == torch.int32 and t.is_contiguous()
    c = communicator(group)
    _lib.check(_lib.lib().decomp_comm_allreduce_min_i32(c.handle, ctypes.c_void_p(t.data_ptr()), t.numel(),
                                                        _lib.stream_ptr()), 'decomp_comm_allreduce_min_i32')
    return t


def reduce_scatter_sum(out, inp, group):
    """out = this rank's slab of the sum over the ranks of inp ([world, ...] float64, contiguous)."""
    assert out.is_cuda and out.is_contiguous() and inp.is_contiguous() and out.dtype == inp.dtype == torch.float64
    c = communicator(group)
    assert inp.numel() == out.numel() * c.world
    _lib.check(_lib.lib().decomp_comm_reduce_scatter_sum_f64(c.handle, ctypes.c_void_p(inp.data_ptr()),
                                                             ctypes.c_void_p(out.data_ptr()), out.numel(),
                                                             _lib.stream_ptr()), 'decomp_comm_reduce_scatter_sum_f64')
    return out


def all_gather(out, inp, group):
    """out ([world, ...]) = the inp of every rank, in rank order (float64, contiguous)."""
    assert out.is_cuda and out.is_contiguous() and inp.is_contiguous() and out.dtype == inp.dtype == torch.float64
    c = communicator(group)
    assert out.numel() == inp.numel() * c.world
    _lib.check(_lib.lib().decomp_comm_allgather_f64(c.handle, ctypes.c_void_p(inp.data_ptr()),
                                                    ctypes.c_void_p(out.data_ptr()), inp.numel(), _lib.stream_ptr()),
               'decomp_comm_allgather_f64')
    return out
