"""The TF32-split (tcgen05) variant of the unmasked ISTA/FISTA iteration.

Tolerances (stated separately from the 1e-10 of the FP64 path, BASELINE.json north_star): the split GEMM carries
~22 significant bits per product and accumulates in FP32, so one GEMM is good to ~1e-6 of |A||B|; after a full
solve x agrees with the FP64 path to 1e-4 (max-norm, relative) and the objective to 1e-5 relative."""
import numpy as np
import pytest
import torch

import golden_cases as gc

pytestmark = pytest.mark.gpu

GEMM_RTOL = 2.0e-6
X_RTOL = 1.0e-4
OBJ_RTOL = 1.0e-5


@pytest.mark.parametrize('M,N,K', [(128, 32, 32), (300, 64, 40), (1000, 256, 256), (4099, 96, 130)])
def test_gemm_nt_tf32x3(M, N, K):
    from decomp_b200 import ops
    from decomp_b200._device import to_device2d
    rng = np.random.RandomState(M + N + K)
    A, B = rng.randn(M, K), rng.randn(N, K)
    Ah, Al = ops.split_tf32(to_device2d(A))
    Bh, Bl = ops.split_tf32(to_device2d(B))
    # the pieces are TF32-valued and reproduce the input to ~2^-22
    rec = Ah.double().cpu().numpy() + Al.double().cpu().numpy()
    assert np.max(np.abs(rec - A)) <= 2.0 ** -21 * np.max(np.abs(A))
    assert int((Ah.view(torch.int32) & 0x1fff).abs().max().item()) == 0
    P = ops.empty_f32(M, N, 'cuda')
    ops.gemm_nt_tf32x3(Ah, Al, Bh, Bl, P)
    torch.cuda.synchronize()
    ref = A.dot(B.T)
    scale = np.abs(A).dot(np.abs(B.T)).max()
    err = np.max(np.abs(P.double().cpu().numpy() - ref)) / scale
    assert err <= GEMM_RTOL, 'rel err %g' % err


@pytest.mark.parametrize('method', ['fista', 'ista', 'fista_pos'])
def test_lasso_tf32x3_vs_fp64(method):
    from decomp_b200 import lasso
    from oracle import decomp_oracle as orc
    A, y, _, _ = gc._lasso_data((3000,), 64, 100, 5, positive=method.endswith('_pos'))
    it64, x64 = lasso.solve(y, A, 0.05, tol=0.0, method=method, maxiter=60)
    it32, x32 = lasso.solve(y, A, 0.05, tol=0.0, method=method, maxiter=60, precision='tf32x3')
    assert it64 == it32 == 59
    err = np.max(np.abs(x32 - x64)) / np.max(np.abs(x64))
    assert err <= X_RTOL, 'x rel err %g' % err
    o64, o32 = orc.lasso_objective(y, A, x64, 0.05), orc.lasso_objective(y, A, x32, 0.05)
    assert abs(o32 - o64) <= OBJ_RTOL * abs(o64)


def test_lasso_tf32x3_complex_and_convergence():
    from decomp_b200 import lasso
    A, y, _, _ = gc._lasso_data((500,), 16, 40, 6, complex_=True)      # 2k = 32 real columns
    it64, x64 = lasso.solve(y, A, 0.05, tol=1e-5, method='fista', maxiter=500)
    it32, x32 = lasso.solve(y, A, 0.05, tol=1e-5, method='fista', maxiter=500, precision='tf32x3')
    assert 0 < it32 < 499 and abs(it32 - it64) <= 10
    assert np.max(np.abs(x32 - x64)) / np.max(np.abs(x64)) <= 1.0e-3


def test_lasso_tf32x3_unsupported_shapes():
    from decomp_b200 import lasso
    A, y, mask, _ = gc._lasso_data((50,), 5, 10, 0)
    with pytest.raises(NotImplementedError):
        lasso.solve(y, A, 0.05, precision='tf32x3')                    # k = 5 is not a multiple of 32
    A, y, mask, _ = gc._lasso_data((50,), 5, 10, 0)
    with pytest.raises(NotImplementedError):
        lasso.solve(y, A, 0.05, mask=mask, precision='tf32x3')          # masked iteration: even real width
    with pytest.raises(ValueError):
        lasso.solve(y, A, 0.05, precision='fp16')


# ------------------------------------------------------------------------------------------------ NMF on tcgen05
NMF_RTOL = 1.0e-4        # D and x after a full solve, relative max-norm, against the FP64 path
NMF_OBJ_RTOL = 1.0e-5    # objective 1/2 |y - x D|^2


@pytest.mark.parametrize('M,N,K', [(500, 300, 100), (260, 1024, 96), (129, 520, 33)])
def test_gemm_nt_tf32x3_tiles_over_n(M, N, K):
    """N beyond one 256-wide MMA: several B boxes, the last one partly outside the matrix (zero-filled by TMA)."""
    from decomp_b200 import ops
    from decomp_b200._device import to_device2d
    rng = np.random.RandomState(M + N + K)
    A, B = rng.randn(M, K), rng.randn(N, K)
    Ah, Al = ops.split_tf32(to_device2d(A))
    Bh, Bl = ops.split_tf32(to_device2d(B))
    P = ops.empty_f32(M, N, 'cuda')
    P.fill_(float('nan'))
    ops.gemm_nt_tf32x3(Ah, Al, Bh, Bl, P)
    torch.cuda.synchronize()
    err = np.max(np.abs(P.double().cpu().numpy() - A.dot(B.T))) / np.abs(A).dot(np.abs(B.T)).max()
    assert err <= GEMM_RTOL, 'rel err %g' % err


@pytest.mark.parametrize('M,N,K,per', [(64, 300, 10000, 4096), (256, 256, 5000, 1024), (40, 70, 100, 4096)])
def test_gemm_nt_tf32x3_splitk(M, N, K, per):
    """Contraction cut into FP32-accumulated slabs that are summed in FP64 (x^T y, x^T x of the NMF sweep)."""
    from decomp_b200 import ops
    from decomp_b200._device import to_device2d, empty2d
    rng = np.random.RandomState(M + N + K)
    out = empty2d(M, N, False, torch.device('cuda', 0))
    ws = ops.gemm_nt_tf32x3_splitk_workspace(M, N, K, 'cuda', per)
    # mixed signs: rounding errors of the FP32 accumulation average out
    A, B = rng.randn(M, K), rng.randn(N, K)
    Ah, Al = ops.split_tf32(to_device2d(A))
    Bh, Bl = ops.split_tf32(to_device2d(B))
    ops.gemm_nt_tf32x3_splitk(Ah, Al, Bh, Bl, out, ws, per)
    torch.cuda.synchronize()
    err = np.max(np.abs(out.cpu().numpy() - A.dot(B.T))) / np.abs(A).dot(np.abs(B.T)).max()
    assert err <= GEMM_RTOL, 'rel err %g' % err
    # same-sign terms (what NMF feeds it): the tensor core truncates its FP32 accumulator, about half an ulp per
    # accumulator update, 3 updates per 8 contraction elements -> a systematic relative deficit of
    # ~ slab / 8 * 3 * 2^-25 (4.6e-5 for slabs of 4096 rows), which is why the slabs are bounded and summed in FP64
    A, B = np.abs(A), np.abs(B)
    Ah, Al = ops.split_tf32(to_device2d(A))
    Bh, Bl = ops.split_tf32(to_device2d(B))
    ops.gemm_nt_tf32x3_splitk(Ah, Al, Bh, Bl, out, ws, per)
    torch.cuda.synchronize()
    ref = A.dot(B.T)
    err = np.max(np.abs(out.cpu().numpy() - ref) / ref)
    bound = max(2.0e-6, 1.5 * min(per, K) / 8 * 3 * 2.0 ** -25)
    print('split-K tf32x3 M=%d N=%d K=%d slab=%d: same-sign max rel err %.3g (bound %.3g)' % (M, N, K, per, err, bound))
    assert err <= bound, 'rel err %g > %g' % (err, bound)


def test_split_transpose_tf32():
    from decomp_b200 import ops
    from decomp_b200._device import to_device2d
    rng = np.random.RandomState(3)
    for rows, cols in [(1, 1), (33, 65), (300, 40), (64, 1000)]:
        A = rng.randn(rows, cols)
        hT, lT = ops.split_transpose_tf32(to_device2d(A))
        h, l = ops.split_tf32(to_device2d(A))
        torch.cuda.synchronize()
        assert hT.shape == (cols, rows)
        assert torch.equal(hT, h.t()) and torch.equal(lT, l.t())


@pytest.mark.parametrize('n,f,k', [(300, 96, 32), (1000, 200, 64), (257, 4100, 256)])
def test_nmf_xupdate_tf32x3_kernel(n, f, k):
    from decomp_b200 import ops
    from decomp_b200._device import to_device2d
    rng = np.random.RandomState(n + f + k)
    y, D, x = np.abs(rng.randn(n, f)), np.abs(rng.randn(k, f)), np.abs(rng.randn(n, k)) + 0.1
    neg = x.dot(D.dot(D.T))
    Yh, Yl = ops.split_tf32(to_device2d(y))
    Dh, Dl = ops.split_tf32(to_device2d(D))
    X = to_device2d(x)
    NEG = ops.empty_f32(n, k, 'cuda')
    NEG.copy_(torch.from_numpy(neg.astype(np.float32)))
    Xh, Xl = ops.empty_f32(n, k, 'cuda'), ops.empty_f32(n, k, 'cuda')
    XTh, XTl = ops.empty_f32(k, n, 'cuda'), ops.empty_f32(k, n, 'cuda')
    ops.nmf_xupdate_tf32x3(Yh, Yl, Dh, Dl, X, NEG, Xh, Xl, XTh, XTl)
    torch.cuda.synchronize()
    ref = x * np.maximum(y.dot(D.T), 0.0) / np.maximum(neg, 1e-15)
    got = X.cpu().numpy()
    err = np.max(np.abs(got - ref) / ref)
    print('nmf x update tf32x3 n=%d f=%d k=%d: max rel err %.3g' % (n, f, k, err))
    assert err <= max(5.0e-6, 1.5 * f / 8 * 3 * 2.0 ** -25)      # same-sign sums over f terms, see the split-K test
    rec = Xh.double().cpu().numpy() + Xl.double().cpu().numpy()
    assert np.max(np.abs(rec - got)) <= 2.0 ** -21 * np.max(np.abs(got))
    assert torch.equal(XTh, Xh.t()) and torch.equal(XTl, Xl.t())


@pytest.mark.parametrize('n,f,k,sweeps', [(3001, 517, 64, 20), (9000, 260, 256, 10)])
def test_nmf_tf32x3_vs_fp64(n, f, k, sweeps):
    """Whole solves: the TF32-split path against the FP64 path (which matches the reference to 1e-10)."""
    from decomp_b200 import nmf
    from oracle import decomp_oracle as orc
    y, D0, _ = gc._nmf_data(n, f, k, 11)
    it64, D64, x64 = nmf.solve(y, D0.copy(), tol=0.0, maxiter=sweeps + 1)
    it32, D32, x32 = nmf.solve(y, D0.copy(), tol=0.0, maxiter=sweeps + 1, precision='tf32x3')
    assert it64 == it32 == sweeps + 1
    eD = np.max(np.abs(D32 - D64)) / np.max(np.abs(D64))
    ex = np.max(np.abs(x32 - x64)) / np.max(np.abs(x64))
    o64, o32 = orc.nmf_objective(y, x64, D64), orc.nmf_objective(y, x32, D32)
    print('nmf tf32x3 n=%d f=%d k=%d sweeps=%d: err_D %.3g err_x %.3g objective %.3g'
          % (n, f, k, sweeps, eD, ex, abs(o32 - o64) / abs(o64)))
    assert eD <= NMF_RTOL and ex <= NMF_RTOL
    assert abs(o32 - o64) <= NMF_OBJ_RTOL * abs(o64)


def test_nmf_tf32x3_convergence_and_unsupported():
    from decomp_b200 import nmf
    y, D0, mask = gc._nmf_data(1501, 130, 32, 5)
    it64, D64, _ = nmf.solve(y, D0.copy(), tol=1e-4, maxiter=400)
    it32, D32, _ = nmf.solve(y, D0.copy(), tol=1e-4, maxiter=400, precision='tf32x3')
    assert 1 < it32 < 400 and abs(it32 - it64) <= 3
    assert np.max(np.abs(D32 - D64)) / np.max(np.abs(D64)) <= 1.0e-3
    with pytest.raises(NotImplementedError):
        nmf.solve(np.abs(y), D0.copy(), mask=mask, likelihood='kl', precision='tf32x3')
    y2, D2, _ = gc._nmf_data(100, 20, 3, 0)
    with pytest.raises(NotImplementedError):
        nmf.solve(y2, D2.copy(), precision='tf32x3')                     # k = 3
    with pytest.raises(ValueError):
        nmf.solve(y, D0.copy(), precision='bf16')


@pytest.mark.parametrize('M,N,K,block', [(64, 300, 1000, 256), (256, 520, 9000, 4096), (32, 32, 100, 128),
                                         (64, 64, 50, 128), (256, 96, 3001, 4096)])
def test_tf32x3_splitk_k_blocked_layout(M, N, K, block):
    """x^T / y^T stored K-blocked ([K / block][rows][block], zero-filled tail): split-K reads block z for piece z;
    the blocked transpose and the blocked x-update output agree with the plain layouts."""
    from decomp_b200 import ops
    from decomp_b200._device import to_device2d, empty2d
    rng = np.random.RandomState(M + N + K)
    At, Bt = rng.randn(K, M), rng.randn(K, N)                        # the operands arrive as [K, rows] (x, y)
    Ah, Al = ops.split_transpose_tf32(to_device2d(At), block=block)
    Bh, Bl = ops.split_transpose_tf32(to_device2d(Bt), block=block)
    nblk = -(-K // block)
    assert Ah.shape == (nblk, M, block) and Bh.shape == (nblk, N, block)
    plain_h, _ = ops.split_transpose_tf32(to_device2d(At))
    torch.cuda.synchronize()
    full = Ah.permute(1, 0, 2).reshape(M, nblk * block)
    assert torch.equal(full[:, :K], plain_h) and float(full[:, K:].abs().sum().item()) == 0.0
    out = empty2d(M, N, False, torch.device('cuda', 0))
    ws = ops.gemm_nt_tf32x3_splitk_workspace(M, N, K, 'cuda', block)
    ops.gemm_nt_tf32x3_splitk(Ah, Al, Bh, Bl, out, ws, block, K=K)
    torch.cuda.synchronize()
    ref = At.T.dot(Bt)
    err = np.max(np.abs(out.cpu().numpy() - ref)) / np.abs(At.T).dot(np.abs(Bt)).max()
    assert err <= 2 * GEMM_RTOL, 'rel err %g' % err      # slabs of up to 4096 FP32-accumulated terms (2.2e-6 measured)


def test_nmf_xupdate_tf32x3_blocked_transpose():
    from decomp_b200 import ops
    from decomp_b200._device import to_device2d
    n, f, k, block = 1000, 96, 64, 256
    rng = np.random.RandomState(7)
    y, D, x = np.abs(rng.randn(n, f)), np.abs(rng.randn(k, f)), np.abs(rng.randn(n, k)) + 0.1
    Yh, Yl = ops.split_tf32(to_device2d(y))
    Dh, Dl = ops.split_tf32(to_device2d(D))
    NEG = ops.empty_f32(n, k, 'cuda')
    NEG.copy_(torch.from_numpy(x.dot(D.dot(D.T)).astype(np.float32)))
    res = []
    for blocked in (False, True):
        X = to_device2d(x)
        Xh, Xl = ops.empty_f32(n, k, 'cuda'), ops.empty_f32(n, k, 'cuda')
        if blocked:
            XTh = ops.empty_f32_blocked(n, k, block, 'cuda', zero_tail=True)
            XTl = ops.empty_f32_blocked(n, k, block, 'cuda', zero_tail=True)
        else:
            XTh, XTl = ops.empty_f32(k, n, 'cuda'), ops.empty_f32(k, n, 'cuda')
        ops.nmf_xupdate_tf32x3(Yh, Yl, Dh, Dl, X, NEG, Xh, Xl, XTh, XTl)
        torch.cuda.synchronize()
        res.append((X, XTh, XTl))
    (X0, Th0, Tl0), (X1, Th1, Tl1) = res
    assert torch.equal(X0, X1)
    nblk = -(-n // block)
    assert torch.equal(Th1.permute(1, 0, 2).reshape(k, nblk * block)[:, :n], Th0)
    assert torch.equal(Tl1.permute(1, 0, 2).reshape(k, nblk * block)[:, :n], Tl0)
    assert float(Th1.permute(1, 0, 2).reshape(k, nblk * block)[:, n:].abs().sum().item()) == 0.0


@pytest.mark.parametrize('M,N,K,block', [(300, 200, 32, 128), (1000, 1024, 128, 4096), (5000, 530, 64, 4096),
                                         (128, 96, 256, 0), (257, 36, 40, 0)])
def test_gemm_nt_mask_tf32x3_kernel(M, N, K, block):
    """F = (A B^T) * mask as a TF32 pair, row-major and transposed (plain / K-blocked), ragged edges included."""
    from decomp_b200 import ops
    from decomp_b200._device import to_device2d
    rng = np.random.RandomState(M + N + K)
    A, B = rng.randn(M, K), rng.randn(N, K)
    mask = (rng.rand(M, N) > 0.3).astype(np.float64) * (1.0 + 0.25 * rng.rand(M, N))
    Ah, Al = ops.split_tf32(to_device2d(A))
    Bh, Bl = ops.split_tf32(to_device2d(B))
    m32 = ops.to_f32(to_device2d(mask))
    assert np.array_equal(m32.cpu().numpy(), mask.astype(np.float32))
    Fh, Fl = ops.empty_f32(M, N, 'cuda'), ops.empty_f32(M, N, 'cuda')
    if block:
        FTh = ops.empty_f32_blocked(M, N, block, 'cuda', zero_tail=True)
        FTl = ops.empty_f32_blocked(M, N, block, 'cuda', zero_tail=True)
    else:
        FTh, FTl = ops.empty_f32(N, M, 'cuda'), ops.empty_f32(N, M, 'cuda')
    ops.gemm_nt_mask_tf32x3(Ah, Al, Bh, Bl, m32, F=(Fh, Fl), FT=(FTh, FTl))
    torch.cuda.synchronize()
    ref = A.dot(B.T) * mask
    scale = np.abs(A).dot(np.abs(B.T)).max() * mask.max()
    got = Fh.double().cpu().numpy() + Fl.double().cpu().numpy()
    err = np.max(np.abs(got - ref)) / scale
    print('masked product tf32x3 %dx%dx%d: rel err %.3g' % (M, N, K, err))
    assert err <= GEMM_RTOL
    assert int((Fh.view(torch.int32) & 0x1fff).abs().max().item()) == 0      # TF32-valued pieces
    assert int((Fl.view(torch.int32) & 0x1fff).abs().max().item()) == 0
    if block:
        nblk = (M + block - 1) // block
        th = FTh.permute(0, 2, 1).reshape(nblk * block, N)
        tl = FTl.permute(0, 2, 1).reshape(nblk * block, N)
        assert torch.equal(th[:M], Fh) and torch.equal(tl[:M], Fl)
        assert float(th[M:].abs().max().item()) == 0.0 if nblk * block > M else True
    else:
        assert torch.equal(FTh, Fh.t()) and torch.equal(FTl, Fl.t())
    # one output only, no mask
    F2h, F2l = ops.empty_f32(M, N, 'cuda'), ops.empty_f32(M, N, 'cuda')
    ops.gemm_nt_mask_tf32x3(Ah, Al, Bh, Bl, None, F=(F2h, F2l))
    torch.cuda.synchronize()
    got2 = F2h.double().cpu().numpy() + F2l.double().cpu().numpy()
    assert np.max(np.abs(got2 - A.dot(B.T))) / scale <= GEMM_RTOL


@pytest.mark.parametrize('n,f,k,sweeps', [(3001, 517, 64, 20), (9000, 260, 128, 10), (700, 1030, 32, 15)])
def test_nmf_masked_tf32x3_vs_fp64(n, f, k, sweeps):
    """Masked 'l2' solves: TF32-split path against the FP64 path (which matches the reference to 1e-10)."""
    from decomp_b200 import nmf
    from oracle import decomp_oracle as orc
    y, D0, mask = gc._nmf_data(n, f, k, 13)
    it64, D64, x64 = nmf.solve(y, D0.copy(), tol=0.0, maxiter=sweeps + 1, mask=mask)
    it32, D32, x32 = nmf.solve(y, D0.copy(), tol=0.0, maxiter=sweeps + 1, mask=mask, precision='tf32x3')
    assert it64 == it32 == sweeps + 1
    eD = np.max(np.abs(D32 - D64)) / np.max(np.abs(D64))
    ex = np.max(np.abs(x32 - x64)) / np.max(np.abs(x64))
    o64, o32 = orc.nmf_objective(y, x64, D64, mask), orc.nmf_objective(y, x32, D32, mask)
    print('masked nmf tf32x3 n=%d f=%d k=%d sweeps=%d: err_D %.3g err_x %.3g objective %.3g'
          % (n, f, k, sweeps, eD, ex, abs(o32 - o64) / abs(o64)))
    assert eD <= NMF_RTOL and ex <= NMF_RTOL
    assert abs(o32 - o64) <= NMF_OBJ_RTOL * abs(o64)


def test_nmf_masked_tf32x3_convergence():
    from decomp_b200 import nmf
    y, D0, mask = gc._nmf_data(1501, 130, 32, 5)
    it64, D64, _ = nmf.solve(y, D0.copy(), tol=1e-4, maxiter=400, mask=mask)
    it32, D32, _ = nmf.solve(y, D0.copy(), tol=1e-4, maxiter=400, mask=mask, precision='tf32x3')
    assert 1 < it32 < 400 and abs(it32 - it64) <= 3
    assert np.max(np.abs(D32 - D64)) / np.max(np.abs(D64)) <= 1.0e-3


@pytest.mark.parametrize('method,complex_,k,f', [('fista', False, 64, 100), ('ista', False, 30, 257),
                                                 ('fista_pos', False, 128, 300), ('fista', True, 16, 40),
                                                 ('ista', True, 33, 130)])
def test_lasso_masked_tf32x3_vs_fp64(method, complex_, k, f):
    """Masked iteration ((w A) * M) A^H with both GEMMs on tcgen05 against the FP64 path."""
    from decomp_b200 import lasso
    from oracle import decomp_oracle as orc
    A, y, mask, _ = gc._lasso_data((2500,), k, f, 9, complex_=complex_, positive=method.endswith('_pos'))
    it64, x64 = lasso.solve(y, A, 0.05, tol=0.0, method=method, maxiter=60, mask=mask)
    it32, x32 = lasso.solve(y, A, 0.05, tol=0.0, method=method, maxiter=60, mask=mask, precision='tf32x3')
    assert it64 == it32 == 59
    err = np.max(np.abs(x32 - x64)) / np.max(np.abs(x64))
    o64, o32 = orc.lasso_objective(y, A, x64, 0.05, mask), orc.lasso_objective(y, A, x32, 0.05, mask)
    print('masked lasso tf32x3 %s complex=%s k=%d f=%d: x rel err %.3g objective %.3g'
          % (method, complex_, k, f, err, abs(o32 - o64) / abs(o64)))
    assert err <= X_RTOL, 'x rel err %g' % err
    assert abs(o32 - o64) <= OBJ_RTOL * abs(o64)


def test_lasso_masked_tf32x3_convergence():
    from decomp_b200 import lasso
    A, y, mask, _ = gc._lasso_data((700,), 32, 60, 4)
    it64, x64 = lasso.solve(y, A, 0.05, tol=1e-5, method='fista', maxiter=500, mask=mask)
    it32, x32 = lasso.solve(y, A, 0.05, tol=1e-5, method='fista', maxiter=500, mask=mask, precision='tf32x3')
    assert 0 < it32 < 499 and abs(it32 - it64) <= 10
    assert np.max(np.abs(x32 - x64)) / np.max(np.abs(x64)) <= 1.0e-3
