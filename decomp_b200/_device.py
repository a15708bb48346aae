"""Device-buffer plumbing: torch owns the memory, nothing here computes.

Arrays handed to the kernels are float64 (complex128 viewed as interleaved doubles) with an
even row pitch, which is what the TMA tensor maps of the GEMM need (16-byte row alignment).
float32 / complex64 inputs are widened on entry and narrowed on exit: the hot path computes
in FP64 throughout.
"""
import ctypes
import os
import threading
import warnings

import numpy as np
import torch

from ._lib import DecompError


def require_cuda():
    if not torch.cuda.is_available():
        raise DecompError('decomp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
    return torch.device('cuda', torch.cuda.current_device())


def is_torch(a):
    return isinstance(a, torch.Tensor)


def np_dtype(a):
    """numpy dtype of a numpy array or torch tensor."""
    if is_torch(a):
        return np.dtype(str(a.dtype).replace('torch.', ''))
    return a.dtype


def empty2d(rows, cols, cplx=False, device=None):
    """Uninitialised [rows, cols] buffer with an even pitch (in doubles)."""
    device = device or require_cuda()
    if cplx:
        return torch.empty((rows, cols), dtype=torch.complex128, device=device)
    pitch = cols + (cols & 1)
    base = torch.empty((max(rows, 1), max(pitch, 2)), dtype=torch.float64, device=device)
    return base[:rows, :cols]


def zeros2d(rows, cols, cplx=False, device=None):
    t = empty2d(rows, cols, cplx, device)
    t.zero_()
    return t


def full2d(rows, cols, value, cplx=False, device=None):
    t = empty2d(rows, cols, cplx, device)
    t.fill_(value)
    return t


_STAGED_DTYPES = (torch.float32, torch.float64, torch.complex64, torch.complex128)
STAGE_MIN_BYTES = 64 << 20    # pageable arrays from this size on go up through the staged path below
STAGE_PIECE_BYTES = (int(os.environ['DECOMP_STAGE_PIECE_KB']) << 10 if 'DECOMP_STAGE_PIECE_KB' in os.environ
                     else int(os.environ.get('DECOMP_STAGE_PIECE_MB', '4')) << 20)


def _stage_threads():
    """memcpy threads of the staged upload: the cores this process may use, shared between the ranks of one box."""
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 4
    ranks = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1')))
    return max(2, min(6, cores // ranks - 1))


STAGE_THREADS = int(os.environ.get('DECOMP_STAGE_THREADS', '0')) or _stage_threads()
STAGE_SLOTS = STAGE_THREADS + 2
_stage = {}
_stage_lock = threading.Lock()    # one staged upload at a time per process: the ring of pinned buffers is shared


def _staged_upload(src, out):
    """Pageable host array -> device through a ring of page-locked staging buffers filled by a few threads.

    torch's own copy from pageable memory stages through one buffer on the calling thread (11 GB/s measured on the
    B200 boxes); several memcpy threads feeding asynchronous copies get closer to what the link carries.  ``src`` is a
    C-contiguous numpy array, ``out`` a contiguous device tensor with the same number of elements.  float64 /
    complex128 sources go through ``decomp_staged_upload`` (native threads, no GIL); float32 / complex64 sources are
    staged by Python threads and widened on the device, piece by piece.
    Enqueues on the current stream and returns when the last piece has been handed to the copy engine."""
    with _stage_lock:
        if 'ring' not in _stage:
            _stage['ring'] = torch.empty(STAGE_PIECE_BYTES * STAGE_SLOTS, dtype=torch.uint8, pin_memory=True)
        ring = _stage['ring']
        if src.dtype in (np.float64, np.complex128):
            from . import _lib
            rc = _lib.lib().decomp_staged_upload(
                ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(src.ctypes.data), src.nbytes,
                ctypes.c_void_p(ring.data_ptr()), STAGE_PIECE_BYTES, STAGE_SLOTS, STAGE_THREADS,
                ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream))
            _lib.check(rc, 'decomp_staged_upload')
            return
        _staged_upload_widening(src, out, ring)


def _staged_upload_widening(src, out, ring):
    """float32 / complex64 pieces are staged by Python threads and converted by the device-side copy."""
    from concurrent.futures import ThreadPoolExecutor
    if 'pool' not in _stage:
        _stage['pool'] = ThreadPoolExecutor(max_workers=STAGE_THREADS)
    pool = _stage['pool']
    bufs = [ring[i * STAGE_PIECE_BYTES:(i + 1) * STAGE_PIECE_BYTES] for i in range(STAGE_SLOTS)]
    src_dtype = getattr(torch, src.dtype.name)
    flat = src.reshape(-1)
    out_flat = out.view(-1)
    total = flat.shape[0]
    step = max(1, STAGE_PIECE_BYTES // src.itemsize)
    pieces = [(e0, min(total, e0 + step)) for e0 in range(0, total, step)]
    stream = torch.cuda.current_stream(out.device)
    events, futures = [None] * STAGE_SLOTS, {}

    def fill(q):
        e0, e1 = pieces[q]
        np.copyto(bufs[q % STAGE_SLOTS].numpy()[:(e1 - e0) * src.itemsize].view(src.dtype), flat[e0:e1])

    submitted = 0
    try:
        for p, (e0, e1) in enumerate(pieces):
            while submitted < len(pieces) and submitted < p + STAGE_SLOTS:
                slot = submitted % STAGE_SLOTS
                if events[slot] is not None and submitted >= STAGE_SLOTS:
                    events[slot].synchronize()        # the copy out of this slot (piece submitted - SLOTS) is done
                futures[submitted] = pool.submit(fill, submitted)
                submitted += 1
            futures.pop(p).result()
            slot = p % STAGE_SLOTS
            piece = bufs[slot][:(e1 - e0) * src.itemsize].view(src_dtype)
            out_flat[e0:e1].copy_(piece, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            events[slot] = ev
    finally:
        for fut in futures.values():                  # after an error: nobody may still be writing into the ring
            try:
                fut.result()
            except Exception:
                pass
        for ev in events:
            if ev is not None:
                ev.synchronize()                      # the staging buffers are reused by the next call


def _from_numpy(a):
    """torch view of a numpy array; arrays torch cannot alias (read-only, negative or otherwise exotic strides) are
    copied into a C-contiguous one first, as the reference's numpy code would accept them."""
    if any(s < 0 for s in a.strides):
        a = np.ascontiguousarray(a)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', UserWarning)       # read-only arrays: the inputs are only ever read
        try:
            return torch.from_numpy(a)
        except (ValueError, TypeError):
            return torch.from_numpy(np.ascontiguousarray(a))


def to_device2d(a, device=None, copy=True):
    """numpy array / torch tensor [rows, cols] -> FP64 (or complex128) device buffer, even pitch.

    With copy=False a suitably laid out CUDA tensor that already lives on ``device`` is used in place (read-only
    inputs)."""
    device = device or require_cuda()
    src = a if is_torch(a) else _from_numpy(a)
    cplx = src.is_complex()
    want = torch.complex128 if cplx else torch.float64
    if (not copy and src.is_cuda and src.device == device and src.dtype == want and src.dim() == 2
            and src.stride(1) == 1 and (cplx or src.stride(0) % 2 == 0) and src.data_ptr() % 16 == 0):
        return src
    out = empty2d(src.shape[0], src.shape[1], cplx, device)
    if (not is_torch(a) and src.dtype in _STAGED_DTYPES and src.dim() == 2 and src.is_contiguous() and out.is_contiguous()
            and src.numel() * src.element_size() >= STAGE_MIN_BYTES and not src.is_pinned()):
        _staged_upload(src.numpy(), out)
        return out
    out.copy_(src, non_blocking=True)
    return out


def to_device1d(a, device=None):
    device = device or require_cuda()
    src = a if is_torch(a) else _from_numpy(np.ascontiguousarray(a))
    return src.to(device=device, dtype=torch.float64, non_blocking=True).contiguous()


def to_host(t, like, dtype):
    """Device result -> same array kind and dtype as the user's input `like`.

    numpy results are copied into page-locked memory from torch's caching host allocator and returned as an array
    that owns that block (it goes back to the cache when the array is dropped): a D2H copy into fresh pageable
    memory runs at ~2 GB/s (page faults), the pinned copy at PCIe speed (~55 GB/s measured)."""
    tdt = getattr(torch, np.dtype(dtype).name)
    if is_torch(like):
        return t.to(dtype=tdt).contiguous()
    src = t.to(dtype=tdt).contiguous()
    host = torch.empty(src.shape, dtype=tdt, pin_memory=True)
    host.copy_(src, non_blocking=True)
    torch.cuda.current_stream(src.device).synchronize()
    return host.numpy()


def array_kind(*arrays):
    """'numpy' or 'torch' for the non-None arguments; TypeError when they are mixed.

    Mirrors the reference's ``get_array_module`` contract (decomp/utils/cp_compat.py:9-15): all
    arrays of one call live in the same array library."""
    kinds = set()
    for a in arrays:
        if a is None:
            continue
        if is_torch(a):
            kinds.add('torch')
        elif isinstance(a, np.ndarray):
            kinds.add('numpy')
        else:
            raise TypeError('expected numpy arrays or torch tensors, given ' + type(a).__name__)
    if len(kinds) > 1:
        raise TypeError('All the data types should be the same.')
    return kinds.pop() if kinds else 'numpy'


def flatten_rows(a):
    """[..., c] -> [prod(...), c] view (numpy or torch)."""
    lead = 1
    for d in a.shape[:-1]:
        lead *= int(d)
    return a.reshape(lead, a.shape[-1])


def is_complex(a):
    return np_dtype(a).kind == 'c'


def real_dtype(dtype):
    """complex -> matching real dtype (decomp/utils/dtype.py:5-14)."""
    return np.zeros(1, dtype).real.dtype
