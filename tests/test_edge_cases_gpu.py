"""Edge cases the reference's own tests touch (tests/test_lasso.py vector / tensor shapes) plus degenerate sizes and
memory layouts: empty batch, single problem, k = 1, f = 1, strided / Fortran-ordered numpy inputs."""
import numpy as np
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu
RTOL = 1.0e-10


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if b.size else 0.0


def test_lasso_empty_batch_and_single_problem():
    from decomp_b200 import lasso
    from oracle import decomp_oracle as orc
    A, y, mask, _ = gc._lasso_data((7,), 5, 10, 0)
    it, x = lasso.solve(y[:0], A, 0.1, tol=0.0, method='fista', maxiter=5)
    assert x.shape == (0, 5) and it == 4
    for m in (None, mask[:1]):
        it, x = lasso.solve(y[:1], A, 0.1, tol=0.0, method='fista', maxiter=30, mask=m)
        it0, x0 = orc.lasso(y[:1], A, 0.1, tol=0.0, method='fista', maxiter=30, mask=m)
        assert it == it0 and rel(x, x0) <= RTOL


@pytest.mark.parametrize('k,f', [(1, 10), (5, 1), (1, 1), (2, 3)])
def test_lasso_degenerate_dictionary_shapes(k, f):
    from decomp_b200 import lasso
    from oracle import decomp_oracle as orc
    rng = np.random.RandomState(k * 10 + f)
    A = rng.randn(k, f) + 0.5
    y = rng.randn(33, f)
    for method in ('ista', 'fista'):
        it, x = lasso.solve(y, A, 0.05, tol=0.0, method=method, maxiter=25)
        it0, x0 = orc.lasso(y, A, 0.05, tol=0.0, method=method, maxiter=25)
        assert it == it0 and rel(x, x0) <= RTOL


def test_strided_and_fortran_ordered_inputs():
    from decomp_b200 import lasso, nmf
    from oracle import decomp_oracle as orc
    A, y, mask, _ = gc._lasso_data((40,), 6, 12, 3)
    y_strided = np.repeat(y, 2, axis=0)[::2]                 # non-contiguous view with the same values
    y_fortran, A_fortran = np.asfortranarray(y), np.asfortranarray(A)
    it0, x0 = orc.lasso(y, A, 0.1, tol=0.0, method='fista', maxiter=20)
    for yy, AA in ((y_strided, A), (y_fortran, A_fortran)):
        it, x = lasso.solve(yy, AA, 0.1, tol=0.0, method='fista', maxiter=20)
        assert it == it0 and rel(x, x0) <= RTOL
    yn, D0, _ = gc._nmf_data(61, 17, 4, 2)
    it0, D_ref, x_ref = orc.nmf_mu(yn, D0.copy(), tol=0.0, maxiter=9)
    it, D, x = nmf.solve(np.asfortranarray(yn), np.asfortranarray(D0), tol=0.0, maxiter=9)
    assert it == it0 and rel(D, D_ref) <= RTOL and rel(x, x_ref) <= RTOL


def test_nmf_single_latent_and_tiny_sizes():
    from decomp_b200 import nmf
    from oracle import decomp_oracle as orc
    for n, f, k in ((1, 1, 1), (2, 5, 1), (9, 3, 2)):
        y, D0, mask = gc._nmf_data(n, f, k, 7)
        for m in (None, mask):
            it0, D_ref, x_ref = orc.nmf_mu(y, D0.copy(), tol=0.0, maxiter=6, mask=m)
            it, D, x = nmf.solve(y, D0.copy(), tol=0.0, maxiter=6, mask=m)
            assert it == it0
            ok = np.isfinite(D_ref).all() and np.isfinite(x_ref).all()
            if ok:
                assert rel(D, D_ref) <= 1e-10 and rel(x, x_ref) <= 1e-10
            else:                                            # degenerate masks can produce 0/0 in the reference too
                assert np.array_equal(np.isfinite(D), np.isfinite(D_ref))


@pytest.mark.parametrize('variant', ['plain', 'x0', 'mask2d', 'mask1d', 'complex', 'pos', 'f32', 'many_chunks', 'pinned'])
def test_lasso_pipelined_host_path_is_bitwise_the_one_piece_solve(variant, monkeypatch):
    """tol = 0 with host arrays: the batch is uploaded / solved / downloaded chunk by chunk (lasso._solve_pipelined);
    every row must come out exactly as from the one-piece device solve."""
    import torch
    from decomp_b200 import lasso
    monkeypatch.setattr(lasso, 'PIPELINE_MIN_BYTES', 0)
    if variant == 'many_chunks':            # bounded device memory: small chunks, two resident at a time
        monkeypatch.setattr(lasso, 'PIPELINE_MAX_CHUNK_BYTES', 5 << 20)
        monkeypatch.setattr(lasso, 'PIPELINE_DEPTH', 2)
    rng = np.random.RandomState(11)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B, k, f = 128 * sms * 5 + 77, 24, 32
    cplx = variant == 'complex'
    A = rng.randn(k, f) + (1j * rng.randn(k, f) if cplx else 0.0)
    y = rng.randn(B, f) + (1j * rng.randn(B, f) if cplx else 0.0)
    kw = {}
    if variant == 'x0':
        kw['x'] = rng.randn(B, k)
    if variant == 'mask2d':
        kw['mask'] = np.rint(rng.uniform(0.3, 1.0, size=(B, f)))
    if variant == 'mask1d':
        kw['mask'] = np.rint(rng.uniform(0.3, 1.0, size=f))
    if variant == 'f32':
        A, y = A.astype(np.float32), y.astype(np.float32)
    if variant in ('pinned', 'many_chunks'):    # page-locked input: short first / last chunk, long middle ones
        yp = torch.empty(y.shape, dtype=torch.float64, pin_memory=True)
        yp.copy_(torch.from_numpy(y))
        y = yp.numpy()
        assert lasso._is_pinned(y) and not lasso._is_pinned(A)
    method = 'fista_pos' if variant == 'pos' else 'fista'
    chunks = lasso._row_chunks(B, f, k * (2 if cplx else 1), torch.device('cuda', 0), pinned=lasso._is_pinned(y))
    assert chunks is not None and (variant != 'many_chunks' or len(chunks) >= 5)
    it, x = lasso.solve(y, A, 0.05, tol=0.0, method=method, maxiter=12, **kw)
    dev = {n: torch.from_numpy(v).cuda() for n, v in kw.items()}
    it_d, x_d = lasso.solve(torch.from_numpy(y).cuda(), torch.from_numpy(A).cuda(), 0.05, tol=0.0, method=method,
                            maxiter=12, **dev)
    assert it == it_d == 11
    assert isinstance(x, np.ndarray) and x.shape == (B, k) and x.dtype == y.dtype
    assert np.array_equal(x, x_d.cpu().numpy())
    assert np.any(x != 0)


def test_lasso_row_chunks_cover_the_batch():
    import torch
    from decomp_b200 import lasso
    dev = torch.device('cuda', 0)
    assert lasso._row_chunks(1000, 64, 16, dev) is None              # below PIPELINE_MIN_BYTES
    for B, f, kc in [(100000, 1024, 256), (19000, 1024, 256), (400000, 512, 1024), (75777, 256, 32)]:
        ch = lasso._row_chunks(B, f, kc, dev)
        if ch is None:
            continue
        assert ch[0][0] == 0 and ch[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(ch, ch[1:])) and all(r1 > r0 for r0, r1 in ch)


@pytest.mark.parametrize('cplx', [False, True])
def test_staged_upload_of_pageable_arrays(cplx, monkeypatch):
    """Big pageable arrays go up through a ring of page-locked buffers filled by several threads
    (_device._staged_upload); here with tiny pieces so that the ring wraps many times."""
    import torch
    from decomp_b200 import _device
    monkeypatch.setattr(_device, 'STAGE_MIN_BYTES', 0)
    monkeypatch.setattr(_device, 'STAGE_PIECE_BYTES', 4096)
    monkeypatch.setattr(_device, '_stage', {})
    rng = np.random.RandomState(2)
    # (4, 5000): one row is several pieces long (short-and-wide arrays used to raise in the worker thread)
    for rows, cols in [(1, 2), (37, 6), (1000, 30), (513, 128), (4, 5000), (1, 2050)]:
        a = rng.randn(rows, cols) + (1j * rng.randn(rows, cols) if cplx else 0.0)
        d = _device.to_device2d(a, torch.device('cuda', 0))
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy(), a)
        a32 = a.astype(np.complex64 if cplx else np.float32)            # widened on the device, piece by piece
        d = _device.to_device2d(a32, torch.device('cuda', 0))
        torch.cuda.synchronize()
        assert d.dtype == (torch.complex128 if cplx else torch.float64)
        assert np.array_equal(d.cpu().numpy(), a32.astype(a.dtype))
    monkeypatch.setattr(_device, '_stage', {})
