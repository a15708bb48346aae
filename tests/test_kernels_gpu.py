"""Unit parity of every C-ABI kernel against numpy on the same inputs (GPU only)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RT = 1.0e-12


def dev(a):
    from decomp_b200._device import to_device2d
    return to_device2d(a)


def rv(t):
    from decomp_b200._lib import rview
    return rview(t)


def host(t):
    return t.cpu().numpy()


def close(a, b, rtol=RT):
    scale = max(np.max(np.abs(b)), 1e-300)
    err = np.max(np.abs(a - b)) / scale
    assert err <= rtol, 'rel err %g' % err


SHAPES = [(1, 1, 1), (5, 3, 7), (8, 8, 16), (130, 70, 45), (257, 129, 300), (1000, 20, 200), (64, 256, 256)]


@pytest.mark.parametrize('M,N,K', SHAPES)
def test_gemm_nt_store(M, N, K):
    from decomp_b200 import ops
    rng = np.random.RandomState(M + N + K)
    A, B = rng.randn(M, K), rng.randn(N, K)
    dA, dB = dev(A), dev(B)
    out = dev(np.zeros((M, N)))
    ops.gemm_nt(dA, dB, ops.epilogue(ops.EPI_STORE, out))
    torch.cuda.synchronize()
    close(host(out), A.dot(B.T))


@pytest.mark.parametrize('M,N,K', [(130, 70, 45), (257, 33, 300)])
def test_gemm_nt_elementwise_epilogues(M, N, K):
    from decomp_b200 import ops
    rng = np.random.RandomState(1)
    A, B = np.abs(rng.randn(M, K)), np.abs(rng.randn(N, K))
    X, O = np.abs(rng.randn(M, N)), rng.randn(M, N)
    mask = np.rint(rng.uniform(0.3, 1, size=(M, N)))
    acc = A.dot(B.T)
    dA, dB, dX, dO, dM = dev(A), dev(B), dev(X), dev(O), dev(mask)
    out = dev(np.zeros((M, N)))
    ops.gemm_nt(dA, dB, ops.epilogue(ops.EPI_MU_NUM, out, x=dX, other=dO))
    close(host(out), X * np.maximum(acc, 0) / np.maximum(O, 1e-15))
    ops.gemm_nt(dA, dB, ops.epilogue(ops.EPI_MU_DEN, out, x=dX, other=dO))
    close(host(out), X * np.maximum(O, 0) / np.maximum(acc, 1e-15))
    ops.gemm_nt(dA, dB, ops.epilogue(ops.EPI_STORE_MASK, out, mask=dM))
    close(host(out), acc * mask)
    ops.gemm_nt(dA, dB, ops.epilogue(ops.EPI_KL_RATIO, out, other=dO, mask=dM))
    close(host(out), (O * mask) / (acc + 1e-15))
    ops.gemm_nt(dA, dB, ops.epilogue(ops.EPI_KL_RATIO, out, other=dO))
    close(host(out), O / (acc + 1e-15))
    # in-place MU update (out aliases x), as the NMF driver uses it
    ops.gemm_nt(dA, dB, ops.epilogue(ops.EPI_MU_NUM, dX, x=dX, other=dO))
    close(host(dX), X * np.maximum(acc, 0) / np.maximum(O, 1e-15))


def test_gemm_nt_complex_mask_epilogue():
    from decomp_b200 import ops
    rng = np.random.RandomState(2)
    Bn, k, f = 77, 6, 19
    w = rng.randn(Bn, k) + 1j * rng.randn(Bn, k)
    A = rng.randn(k, f) + 1j * rng.randn(k, f)
    mask = np.rint(rng.uniform(0.3, 1, size=(Bn, f)))
    dw, dA, dM = dev(w), dev(A), dev(mask)
    rhs = ops.make_rhs(rv(dA), True, False)
    out = dev(np.zeros((Bn, f), dtype=complex))
    ops.gemm_nt(rv(dw), rhs, ops.epilogue(ops.EPI_STORE_MASK, rv(out), cwidth=2, mask=dM))
    close(host(out), w.dot(A) * mask)
    rhs_h = ops.make_rhs(rv(dA), True, True)
    y = rng.randn(Bn, f) + 1j * rng.randn(Bn, f)
    out2 = dev(np.zeros((Bn, k), dtype=complex))
    ops.gemm_nt(rv(dev(y)), rhs_h, ops.epilogue(ops.EPI_STORE, rv(out2)))
    close(host(out2), y.dot(np.conj(A.T)))


def _prox_ref(w, G, yAt, xprev, step, alpha, tol, mom, kind, rowfac=None):
    z = w + step * (yAt - w.dot(G))
    thr = step * (alpha * rowfac[:, None]) if rowfac is not None else step * alpha
    if kind == 'complex':
        r = np.abs(z)
        xn = np.maximum(r - thr, 0) * (z / (r + 1e-15))
    elif kind == 'positive':
        xn = np.maximum(z - thr, 0)
    else:
        xn = np.maximum(np.abs(z) - thr, 0) * np.sign(z)
    wn = xn + mom * (xn - xprev)
    viol = not (np.max(np.abs(xn - xprev) - tol) < 0)
    return xn, wn, viol


@pytest.mark.parametrize('kind', ['real', 'positive', 'complex'])
@pytest.mark.parametrize('rowvec', [False, True])
def test_gemm_nt_prox(kind, rowvec):
    from decomp_b200 import ops
    rng = np.random.RandomState(3)
    Bn, k = 203, 23
    cplx = kind == 'complex'

    def randn(*s):
        return rng.randn(*s) + 1j * rng.randn(*s) if cplx else rng.randn(*s)

    A = randn(k, 31)
    G = A.dot(np.conj(A.T))
    w, yAt, xprev = randn(Bn, k), randn(Bn, k) * 5, randn(Bn, k)
    alpha = np.abs(rng.randn(k)) * 0.5
    tol = np.full(k, 1e-3)
    rowfac = np.abs(rng.randn(Bn)) + 0.5 if rowvec else None
    step = 1.0 / np.max(np.sum(np.abs(G), axis=0))
    xn_ref, wn_ref, viol = _prox_ref(w, G, yAt, xprev, step, alpha, tol, 0.37, kind, rowfac)

    dG = dev(G)
    rhs = ops.make_rhs(rv(dG), cplx, False)
    dw, dy, dxp = dev(w), dev(yAt), dev(xprev)
    xn, wn = dev(np.zeros_like(w)), dev(np.zeros_like(w))
    dalpha, dtol = ops.vector(k, 'cuda'), ops.vector(k, 'cuda')     # readable up to an even element count
    dalpha.copy_(torch.from_numpy(alpha))
    dtol.copy_(torch.from_numpy(tol))
    drow = torch.from_numpy(rowfac).cuda() if rowvec else None
    dstep = torch.zeros(1, dtype=torch.float64, device='cuda')
    ops.gershgorin_step(rv(dG), cplx, dstep)
    close(host(dstep), np.array([step]))
    latch = torch.zeros(1, dtype=torch.int32, device='cuda')
    scratch = torch.zeros(2, dtype=torch.int32, device='cuda')
    shrink = {'real': ops.SHRINK_REAL, 'positive': ops.SHRINK_POSITIVE, 'complex': ops.SHRINK_COMPLEX}[kind]
    epi = ops.epilogue(ops.EPI_PROX, rv(xn), cwidth=2 if cplx else 1, out2=rv(wn), x=rv(dw), other=rv(dy),
                       prev=rv(dxp), colvec=dalpha, colvec2=dtol, rowvec=drow, step=dstep, momentum=0.37,
                       shrink=shrink, check=True, latch=latch, scratch=scratch, latch_value=7)
    ops.gemm_nt(rv(dw), rhs, epi)
    torch.cuda.synchronize()
    close(host(xn), xn_ref)
    close(host(wn), wn_ref)
    assert viol and int(latch.item()) == 0
    assert scratch.tolist() == [0, 0]
    # converged case: prev == new iterate -> latch fires, and a latched launch is a no-op
    dxp2 = dev(xn_ref)
    epi = ops.epilogue(ops.EPI_PROX, rv(xn), cwidth=2 if cplx else 1, out2=rv(wn), x=rv(dw), other=rv(dy),
                       prev=rv(dxp2), colvec=dalpha, colvec2=dtol, rowvec=drow, step=dstep, momentum=0.0,
                       shrink=shrink, check=True, latch=latch, scratch=scratch, latch_value=7)
    ops.gemm_nt(rv(dw), rhs, epi)
    torch.cuda.synchronize()
    assert int(latch.item()) == 7
    xn.zero_()
    ops.gemm_nt(rv(dw), rhs, epi, skip=latch)
    torch.cuda.synchronize()
    assert float(xn.abs().max().item()) == 0.0


@pytest.mark.parametrize('kind', ['real', 'positive', 'complex'])
@pytest.mark.parametrize('Bn,k', [(203, 23), (1000, 96), (129, 5)])
def test_gemm_nt_proxq(kind, Bn, k):
    """The fused unmasked ISTA/FISTA launch: z = c + w Q with Q = I - step G, c = step yAh (lasso.py:245-246)."""
    from decomp_b200 import ops
    rng = np.random.RandomState(Bn + k)
    cplx = kind == 'complex'
    cw = 2 if cplx else 1

    def randn(*s):
        return rng.randn(*s) + 1j * rng.randn(*s) if cplx else rng.randn(*s)

    A = randn(k, k + 9)
    G = A.dot(np.conj(A.T))
    w, yAt, xprev = randn(Bn, k), randn(Bn, k) * 5, randn(Bn, k)
    alpha = np.abs(rng.randn(k)) * 0.5
    tol = np.full(k, 1e-3)
    step = 1.0 / np.max(np.sum(np.abs(G), axis=0))
    xn_ref, wn_ref, viol = _prox_ref(w, G, yAt, xprev, step, alpha, tol, 0.37, kind)

    dG = dev(G)
    dstep = torch.zeros(1, dtype=torch.float64, device='cuda')
    dalpha, dtol, dthr = ops.vector(k, 'cuda'), ops.vector(k, 'cuda'), ops.vector(k, 'cuda')
    dalpha.copy_(torch.from_numpy(alpha))
    dtol.copy_(torch.from_numpy(tol))
    ops.gershgorin_step(rv(dG), cplx, dstep, alpha_scaled=dalpha, thr_out=dthr)
    Q = dev(np.zeros_like(G))
    ops.lasso_q(rv(dG), cplx, dstep, rv(Q))
    close(host(Q), np.eye(k) - step * G)
    rhs = ops.make_rhs(rv(Q), cplx, False)
    dc = dev(yAt)
    ops.scale_scalar(rv(dc), dstep, rv(dc))
    close(host(dc), step * yAt)
    dw, dxp = dev(w), dev(xprev)
    xn, wn = dev(np.zeros_like(w)), dev(np.zeros_like(w))
    latch = torch.zeros(1, dtype=torch.int32, device='cuda')
    scratch = torch.zeros(2, dtype=torch.int32, device='cuda')
    shrink = {'real': ops.SHRINK_REAL, 'positive': ops.SHRINK_POSITIVE, 'complex': ops.SHRINK_COMPLEX}[kind]

    def epi(out, prev, momentum):
        return ops.epilogue(ops.EPI_PROXQ, rv(out), cwidth=cw, out2=rv(wn), other=rv(dc), prev=rv(prev), colvec=dthr,
                            colvec2=dtol, flags=ops.EPI_FLAG_COLVEC_IS_THRESHOLD, momentum=momentum, shrink=shrink,
                            check=True, latch=latch, scratch=scratch, latch_value=7)

    ops.gemm_nt(rv(dw), rhs, epi(xn, dxp, 0.37))
    torch.cuda.synchronize()
    close(host(xn), xn_ref, 1e-11)
    close(host(wn), wn_ref, 1e-11)
    assert viol and int(latch.item()) == 0 and scratch.tolist() == [0, 0]
    # in place on the previous iterate (how the solver runs it), then the converged case fires the latch
    dxp_inplace = dev(xprev)
    ops.gemm_nt(rv(dw), rhs, epi(dxp_inplace, dxp_inplace, 0.37))
    close(host(dxp_inplace), xn_ref, 1e-11)
    dxp2 = dev(host(xn))
    ops.gemm_nt(rv(dw), rhs, epi(xn, dxp2, 0.0))
    torch.cuda.synchronize()
    assert int(latch.item()) == 7
    xn.zero_()
    ops.gemm_nt(rv(dw), rhs, epi(xn, dxp2, 0.0), skip=latch)
    torch.cuda.synchronize()
    assert float(xn.abs().max().item()) == 0.0


@pytest.mark.parametrize('K,M,N', [(1, 1, 1), (37, 5, 9), (5000, 20, 200), (4099, 130, 70), (20000, 256, 64)])
def test_gemm_tn(K, M, N):
    from decomp_b200 import ops
    rng = np.random.RandomState(K)
    A, B = rng.randn(K, M), rng.randn(K, N)
    dA, dB = dev(A), dev(B)
    out = dev(np.zeros((M, N)))
    ops.gemm_tn(dA, dB, out, combine=0)
    torch.cuda.synchronize()
    ref = A.T.dot(B)
    close(host(out), ref)
    first = host(out).copy()
    ops.gemm_tn(dA, dB, out, combine=0)
    assert np.array_equal(host(out), first), 'split-K reduction must be bitwise reproducible'
    ops.gemm_tn(dA, dB, out, combine=1, beta=0.25)
    close(host(out), 0.25 * ref + ref)


@pytest.mark.parametrize('K,M,N', [(300, 3, 5), (2500, 12, 33)])
def test_gemm_tn_complex(K, M, N):
    from decomp_b200 import ops
    rng = np.random.RandomState(K)
    A = rng.randn(K, M) + 1j * rng.randn(K, M)
    B = rng.randn(K, N) + 1j * rng.randn(K, N)
    dA, dB = dev(A), dev(B)
    out = dev(np.zeros((M, N), dtype=complex))
    ops.gemm_tn(rv(dA), rv(dB), rv(out), combine=2)
    ref = np.conj(A.T).dot(B)
    close(host(out), ref)
    ops.gemm_tn(rv(dA), rv(dB), rv(out), combine=3, beta=0.5)
    close(host(out), 0.5 * ref + ref)


def test_make_rhs_real():
    from decomp_b200 import ops
    rng = np.random.RandomState(0)
    S = rng.randn(37, 101)
    dS = dev(S)
    close(host(ops.make_rhs(dS, False, False)), S.T)
    close(host(ops.make_rhs(dS, False, True)), S)


def test_vector_kernels():
    from decomp_b200 import ops
    rng = np.random.RandomState(5)
    A = rng.randn(41, 67)
    Ac = rng.randn(41, 33) + 1j * rng.randn(41, 33)
    dA, dAc = dev(A), dev(Ac)
    close(host(ops.row_norms(dA, False)), np.sqrt((A * A).sum(-1)))
    close(host(ops.row_norms(rv(dAc), True)), np.sqrt((np.abs(Ac) ** 2).sum(-1)))
    close(host(ops.col_sums(dA, 0.5)), A.sum(0) * 0.5)
    close(host(ops.row_sums(dA, 2.0)), A.sum(1) * 2.0)
    big = rng.randn(9000, 13)
    close(host(ops.col_sums(dev(big), 1.0 / 9000)), big.mean(0))
    rs, cs = np.abs(rng.randn(41)) + 0.1, np.abs(rng.randn(67)) + 0.1
    drs, dcs = torch.from_numpy(rs).cuda(), torch.from_numpy(cs).cuda()
    out = dev(np.zeros_like(A))
    close(host(ops.scale(dA, out, rowscale=drs, invert_row=True, colscale=dcs)), A / rs[:, None] * cs)
    cs2 = np.abs(rng.randn(33)) + 0.1
    outc = dev(np.zeros_like(Ac))
    ops.scale(rv(dAc), rv(outc), cwidth=2, colscale=torch.from_numpy(cs2).cuda(), invert_col=True)
    close(host(outc), Ac / cs2)
    mask = np.rint(rng.uniform(0.3, 1, size=A.shape))
    close(host(ops.mask_mul(dA, dev(mask), out)), A * mask)
    maskc = np.rint(rng.uniform(0.3, 1, size=Ac.shape))
    ops.mask_mul(rv(dAc), dev(maskc), rv(outc), cwidth=2)
    close(host(outc), Ac * maskc)
    idx = rng.permutation(41)
    g = dev(np.zeros_like(A))
    ops.gather_rows(dA, torch.from_numpy(idx).cuda(), g)
    close(host(g), A[idx])


@pytest.mark.parametrize('cplx', [False, True])
def test_normalize_rows_and_latch(cplx):
    from decomp_b200 import ops
    rng = np.random.RandomState(6)
    D = rng.randn(9, 45) + (1j * rng.randn(9, 45) if cplx else 0)
    ref_prev = D / np.sqrt((np.abs(D) ** 2).sum(-1, keepdims=True)) + 1e-3 * rng.randn(9, 45)
    dD, dR = dev(D), dev(ref_prev)
    out = dev(np.zeros_like(D))
    latch = torch.zeros(1, dtype=torch.int32, device='cuda')
    scratch = torch.zeros(1, dtype=torch.int32, device='cuda')
    maxdiff = torch.zeros(2, dtype=torch.float64, device='cuda')
    ops.normalize_rows(rv(dD), rv(out), cplx, True, D_ref=rv(dR), tol=1e-6, latch=latch, latch_value=5,
                       maxdiff=maxdiff, scratch=scratch)
    want = D / np.sqrt((np.abs(D) ** 2).sum(-1, keepdims=True))
    close(host(out), want)
    close(host(maxdiff)[1:], np.array([np.max(np.abs(ref_prev - want))]))
    assert int(latch.item()) == 0 and float(maxdiff[0].item()) == 0.0
    ops.normalize_rows(rv(dD), rv(out), cplx, True, D_ref=rv(dR), tol=1.0, latch=latch, latch_value=5,
                       maxdiff=maxdiff, scratch=scratch)
    assert int(latch.item()) == 5
    big = D * 10
    ops.normalize_rows(rv(dev(big)), rv(out), cplx, False)
    close(host(out), big / np.sqrt(np.maximum((np.abs(big) ** 2).sum(-1, keepdims=True), 1.0)))
    small = D * 1e-3
    ops.normalize_rows(rv(dev(small)), rv(out), cplx, False)
    close(host(out), small)


@pytest.mark.parametrize('cplx', [False, True])
@pytest.mark.parametrize('k,f', [(3, 5), (9, 70), (40, 300), (130, 5000), (1500, 3000)])
def test_dl_sweep(cplx, k, f):
    """Slice-resident sweep (D slice in shared memory; (130, 5000): 34 columns per block, 64-wide thread rows) and,
    for (1500, 3000), the L2-streaming kernel that serves slices too large for shared memory."""
    from decomp_b200 import ops
    if cplx and k > 1000:
        pytest.skip('the real case covers the large-slice kernel')
    rng = np.random.RandomState(k * f)

    def randn(*s):
        return rng.randn(*s) + 1j * rng.randn(*s) if cplx else rng.randn(*s)

    X = randn(4 * k + 3, k)
    S = np.conj(X.T).dot(X)
    T = randn(k, f) * 3
    D = randn(k, f)
    D = D / np.sqrt((np.abs(D) ** 2).sum(-1, keepdims=True))
    Dn = D.copy()
    for a in range(k):
        u = (T[a] - np.dot(S[a], Dn)) / (S[a, a] + 1e-15) + Dn[a]
        Dn[a] = u / np.sqrt(np.maximum(np.sum(np.abs(u) ** 2), 1.0))
    dD = dev(D)
    ops.dl_sweep(rv(dev(S)), rv(dev(T)), rv(dD), cplx)
    torch.cuda.synchronize()
    # The sweep is a k-step recurrence on synthetic (random T, ill-conditioned) statistics, which amplifies the
    # rounding of every dot product: the bar is self-calibrating -- the same numpy recurrence with the dot products
    # accumulated in extended precision (k > 200: in the opposite order) moves the answer by `spread`; the kernel has to
    # stay within 1e-10 or 20x that spread.  Real dictionary-learning states sit at ~1e-15 (tools/parity_errors.py).
    if k <= 200:
        Da = D.astype(np.clongdouble if cplx else np.longdouble)
        Sl, Tl = S.astype(Da.dtype), T.astype(Da.dtype)
        for a in range(k):
            u = (Tl[a] - np.dot(Sl[a], Da)) / (Sl[a, a] + 1e-15) + Da[a]
            Da[a] = u / np.sqrt(np.maximum(np.sum(np.abs(u) ** 2), 1.0))
    else:               # extended precision has no BLAS: use float64 with the atoms summed in the opposite order
        Da = D.copy()
        for a in range(k):
            u = (T[a] - np.dot(S[a][::-1], Da[::-1])) / (S[a, a] + 1e-15) + Da[a]
            Da[a] = u / np.sqrt(np.maximum(np.sum(np.abs(u) ** 2), 1.0))
    spread = float(np.max(np.abs(Da.astype(Dn.dtype) - Dn)) / np.max(np.abs(Dn)))
    err = float(np.max(np.abs(host(dD) - Dn)) / np.max(np.abs(Dn)))
    print('dl_sweep k=%d f=%d complex=%s: kernel vs numpy %.3g, numpy (float64) vs numpy (extended) %.3g'
          % (k, f, cplx, err, spread))
    assert err <= max(1e-10, 20.0 * spread), (err, spread)


@pytest.mark.parametrize('cplx', [False, True])
@pytest.mark.parametrize('k,f,mb', [(3, 5, 20), (7, 33, 50)])
def test_dl_masked_stats_and_update(cplx, k, f, mb):
    from decomp_b200 import ops
    from decomp_b200._device import zeros2d, empty2d
    rng = np.random.RandomState(k * f)

    def randn(*s):
        return rng.randn(*s) + 1j * rng.randn(*s) if cplx else rng.randn(*s)

    X = randn(mb, k)
    mask = np.rint(rng.uniform(0.3, 1, size=(mb, f)))
    S0 = randn(k, f, k)
    beta = 0.3
    xh = np.conj(X.T)
    S_ref = beta * S0 + np.tensordot(xh, np.expand_dims(X, -2) * np.expand_dims(mask, -1), axes=1)
    cw = 2 if cplx else 1
    dX, dM = dev(X), dev(mask)
    dS = torch.from_numpy(np.ascontiguousarray(S0)).cuda()          # [k, f, k] (complex) contiguous
    dS_r = torch.view_as_real(dS).reshape(k, f, k * 2) if cplx else dS
    W = empty2d(mb, k, cplx)
    for a in range(k):
        ops.dl_atom_weighted(rv(dX), cplx, a, rv(W))
        ops.gemm_tn(dM, rv(W), dS_r[a], combine=1, beta=beta)
    torch.cuda.synchronize()
    close(host(dS), S_ref)
    # Jacobi update
    T = randn(k, f)
    D = randn(k, f)
    Dn = D.copy()
    for a in range(k):
        SaD = np.einsum('jk,kj->j', S_ref[a], D)
        Saa = np.sum(S_ref[a, :, a] + 1e-15)
        u = (T[a] - SaD) / Saa + Dn[a]
        Dn[a] = u / np.sqrt(np.maximum(np.sum(np.abs(u) ** 2), 1.0))
    dOut = dev(np.zeros_like(D))
    ws = torch.empty(f * k * cw, dtype=torch.float64, device='cuda')
    ops.dl_masked_update(dS_r, rv(dev(T)), rv(dev(D)), rv(dOut), cplx, ws)
    torch.cuda.synchronize()
    close(host(dOut), Dn, 1e-10)


@pytest.mark.parametrize('cplx', [False, True])
@pytest.mark.parametrize('k,f,mb', [(3, 5, 20), (9, 33, 70)])
def test_dl_packed_masked_stats(cplx, k, f, mb):
    """Hermitian half of S via transposed pair products + NT GEMM + scatter, then mirrored
    (dictionary_learning.py:210-213)."""
    from decomp_b200 import ops
    from decomp_b200._device import empty2d
    from decomp_b200.dictionary_learning import _pair_chunks
    rng = np.random.RandomState(k * f + mb)

    def randn(*s):
        return rng.randn(*s) + 1j * rng.randn(*s) if cplx else rng.randn(*s)

    X = randn(mb, k)
    mask = np.rint(rng.uniform(0.3, 1, size=(mb, f)))
    S0 = randn(k, f, k)
    S0 = 0.5 * (S0 + np.conj(np.transpose(S0, (2, 1, 0))))          # Hermitian in (a, b) like the real statistics
    beta = 0.3
    S_ref = beta * S0 + np.tensordot(np.conj(X.T), np.expand_dims(X, -2) * np.expand_dims(mask, -1), axes=1)
    cw = 2 if cplx else 1
    dX, dM = dev(X), dev(mask)
    dS = torch.from_numpy(np.ascontiguousarray(S0)).cuda()
    dS_r = torch.view_as_real(dS).reshape(k, f, k * 2) if cplx else dS
    Xt, Mt = empty2d(k * cw, mb, False, 'cuda'), empty2d(f, mb, False, 'cuda')
    ops.make_rhs(rv(dX), False, False, out=Xt)
    ops.make_rhs(dM, False, False, out=Mt)
    for colA, colB in _pair_chunks(k, 7, 'cuda'):
        wd = colA.numel()
        Wt = empty2d(wd * cw, mb, False, 'cuda')
        P = empty2d(f, wd, cplx, 'cuda')
        ops.dl_pair_products_t(Xt, cplx, colA, colB, Wt)
        ops.gemm_nt(Mt, Wt, ops.epilogue(ops.EPI_STORE, rv(P)))
        ops.dl_scatter_stats(rv(P), cplx, colA, colB, k, beta, dS_r)
    ops.dl_mirror(dS_r, k, f, cplx)
    torch.cuda.synchronize()
    close(host(dS), S_ref)


@pytest.mark.parametrize('cplx', [False, True])
@pytest.mark.parametrize('k,f', [(1, 3), (2, 1), (33, 5), (70, 9), (128, 2)])
def test_dl_mirror_tiles(cplx, k, f):
    """Lower triangle := conjugate of the upper one, the upper one and the diagonal untouched; k across several
    32-wide tiles and not a multiple of the tile."""
    from decomp_b200 import ops
    rng = np.random.RandomState(k * 100 + f)
    S = rng.randn(k, f, k) + (1j * rng.randn(k, f, k) if cplx else 0.0)
    ref = S.copy()
    for a in range(k):
        for b in range(a + 1, k):
            ref[b, :, a] = np.conj(S[a, :, b])
    dS = torch.from_numpy(np.ascontiguousarray(S)).cuda()
    dS_r = torch.view_as_real(dS).reshape(k, f, k * 2) if cplx else dS
    ops.dl_mirror(dS_r, k, f, cplx)
    torch.cuda.synchronize()
    assert np.array_equal(host(dS), ref)


@pytest.mark.parametrize('M', [20000, 131072])
def test_gemm_nt_large_k_epilogues_are_exact(M):
    """Long mainloops (K = 4096) with every fused epilogue, repeated: the hand-over of accumulators to the epilogue
    warps and the operand ring must not depend on timing."""
    from decomp_b200 import ops
    torch.manual_seed(M)
    N, K = 256, 4096
    A = torch.rand((M, K), dtype=torch.float64, device='cuda')
    B = torch.rand((N, K), dtype=torch.float64, device='cuda')
    X = torch.rand((M, N), dtype=torch.float64, device='cuda') + 0.5
    O = torch.rand((M, N), dtype=torch.float64, device='cuda') + 0.5
    acc = A @ B.T

    def check(out, ref, what):
        err = float(((out - ref).abs() / ref.abs()).max().item())
        assert err <= 1e-12, '%s rel err %g' % (what, err)

    for rep in range(3):
        out = torch.empty_like(X)
        ops.gemm_nt(A, B, ops.epilogue(ops.EPI_STORE, out))
        check(out, acc, 'STORE')
        ops.gemm_nt(A, B, ops.epilogue(ops.EPI_STORE_MASK, out, mask=O))
        check(out, acc * O, 'STORE_MASK')
        ops.gemm_nt(A, B, ops.epilogue(ops.EPI_KL_RATIO, out, other=O, mask=X))
        check(out, O * X / (acc + 1e-15), 'KL_RATIO')
        for inplace in (False, True):
            Xc = X.clone()
            dst = Xc if inplace else out
            ops.gemm_nt(A, B, ops.epilogue(ops.EPI_MU_NUM, dst, x=Xc, other=O))
            check(dst, X * acc / O, 'MU_NUM')
            Xc = X.clone()
            dst = Xc if inplace else out
            ops.gemm_nt(A, B, ops.epilogue(ops.EPI_MU_DEN, dst, x=Xc, other=O))
            check(dst, X * O / acc, 'MU_DEN')


def test_gemm_nt_large_k_prox_is_exact():
    """The masked-path proximal epilogue (three operand streams) behind a long mainloop, repeated."""
    from decomp_b200 import ops
    torch.manual_seed(7)
    M, N, K = 131072, 256, 4096
    A = torch.rand((M, K), dtype=torch.float64, device='cuda') - 0.5
    B = torch.rand((N, K), dtype=torch.float64, device='cuda') - 0.5
    w = torch.randn((M, N), dtype=torch.float64, device='cuda')
    ya = torch.randn((M, N), dtype=torch.float64, device='cuda') * 20
    xp = torch.randn((M, N), dtype=torch.float64, device='cuda')
    alpha = ops.vector(N, 'cuda')
    alpha.copy_(torch.rand(N, dtype=torch.float64, device='cuda'))
    tolv = ops.vector(N, 'cuda')
    rowfac = torch.rand(M, dtype=torch.float64, device='cuda') + 0.5
    step = torch.full((1,), 0.01, dtype=torch.float64, device='cuda')
    acc = A @ B.T
    z = w + 0.01 * (ya - acc)
    thr = 0.01 * (alpha[None, :] * rowfac[:, None])
    x_ref = torch.clamp(z.abs() - thr, min=0) * torch.sign(z)
    w_ref = x_ref + 0.3 * (x_ref - xp)
    for rep in range(3):
        xn, wn = torch.empty_like(w), torch.empty_like(w)
        epi = ops.epilogue(ops.EPI_PROX, xn, out2=wn, x=w, other=ya, prev=xp, colvec=alpha, colvec2=tolv, rowvec=rowfac,
                           step=step, momentum=0.3, shrink=ops.SHRINK_REAL)
        ops.gemm_nt(A, B, epi)
        torch.cuda.synchronize()
        scale = float(x_ref.abs().max().item())
        assert float((xn - x_ref).abs().max().item()) <= 1e-11 * scale
        assert float((wn - w_ref).abs().max().item()) <= 1e-11 * scale


@pytest.mark.parametrize('method,cplx', [('fista', False), ('ista', False), ('fista_pos', False), ('fista', True),
                                         ('acc_ista', False), ('acc_ista', True)])
@pytest.mark.parametrize('k', [32, 64, 128, 256, 6, 20, 100, 250])
def test_lasso_resident_kernel_matches_the_per_iteration_kernel(method, cplx, k, monkeypatch):
    """Several iterations per launch with the iterate resident on chip (decomp_lasso_resident_f64) against one
    launch per iteration (DECOMP_EPI_PROXQ): same stopping iteration, x equal to rounding (the resident kernel starts
    its accumulators from c, the other one adds c after the sum), ragged last row block."""
    from decomp_b200 import lasso
    if cplx:
        k //= 2
    rng = np.random.RandomState(k + 7)
    f, B = 48, (148 * 32 * 2 + 37 if k >= 16 else 1003)            # narrow problems are zero-padded to 32 columns
    A = rng.randn(k, f) + (1j * rng.randn(k, f) if cplx else 0.0)
    xt = rng.randn(B, k) * np.rint(rng.uniform(size=(B, k)))
    y = xt.dot(A) + 0.1 * rng.randn(B, f) + (0.1j * rng.randn(B, f) if cplx else 0.0)
    dy, dA = torch.from_numpy(y).cuda(), torch.from_numpy(A).cuda()
    fired = 0
    for tol, maxiter in [(0.0, 1), (0.0, 2), (0.0, 12), (0.0, 45), (1e-3, 200), (5e-2, 300), (1e-12, 35), (1e-12, 31),
                         (1e-1, 11)]:
        monkeypatch.setattr(lasso, 'USE_RESIDENT', True)
        it1, x1 = lasso.solve(dy, dA, 0.1, tol=tol, method=method, maxiter=maxiter)
        monkeypatch.setattr(lasso, 'USE_RESIDENT', False)
        it0, x0 = lasso.solve(dy, dA, 0.1, tol=tol, method=method, maxiter=maxiter)
        assert it1 == it0, (tol, maxiter, it1, it0)
        err = float((x1 - x0).abs().max()) / max(float(x0.abs().max()), 1e-300)   # acc_ista, maxiter 1: x = 0
        assert err <= 1e-12, (tol, maxiter, err)
        fired += int(it1 < maxiter - 1)
    assert fired > 0                          # the latch did fire inside a multi-iteration launch sequence


def test_store_mask_fast_paths_are_exact():
    """The STORE_MASK epilogue multiplies by weights 0 and 1 with integer instructions: same bits as the FP64 product,
    signed zeros and non-finite accumulators included; other weights take the multiply."""
    from decomp_b200 import ops
    torch.manual_seed(3)
    M, N, K = 300, 130, 40
    A = torch.randn((M, K), dtype=torch.float64, device='cuda')
    B = torch.randn((N, K), dtype=torch.float64, device='cuda')
    A[7, 3] = float('inf')
    A[9, 0] = float('nan')
    mask = torch.tensor([0.0, 1.0, 0.5, 2.0], dtype=torch.float64, device='cuda')[
        torch.randint(0, 4, (M, N), device='cuda')].contiguous()
    out = torch.empty((M, N), dtype=torch.float64, device='cuda')
    plain = torch.empty_like(out)
    ops.gemm_nt(A, B, ops.epilogue(ops.EPI_STORE, plain))
    ops.gemm_nt(A, B, ops.epilogue(ops.EPI_STORE_MASK, out, mask=mask))
    torch.cuda.synchronize()
    ref = plain * mask
    nan = torch.isnan(ref)
    assert torch.equal(torch.isnan(out), nan) and bool(nan.any())
    same_bits = out.view(torch.int64) == ref.view(torch.int64)
    assert bool((same_bits | nan).all())
    assert bool(((ref == 0) & torch.signbit(ref)).any())      # negative zeros were exercised


@pytest.mark.parametrize('cplx', [False, True])
@pytest.mark.parametrize('M,k,f', [(300, 32, 70), (1000, 64, 333), (257, 128, 1024), (129, 16, 40)])
def test_gemm_b2b_masked(cplx, M, k, f):
    """((W R^T) * mask) R in one kernel (lasso.py:259-271, grads.py:112-115) against numpy; ragged rows and channels.
    Complex data goes through the real embedding: one operand serves (w A) and (. A^H)."""
    import torch
    from decomp_b200 import ops
    from decomp_b200._device import to_device2d, empty2d
    from decomp_b200._lib import rview
    cw = 2 if cplx else 1
    if not ops.gemm_b2b_masked_supported(k * cw):
        pytest.skip('width not covered by the fused kernel')
    rng = np.random.RandomState(M + k + f)

    def randn(*s):
        return rng.randn(*s) + 1j * rng.randn(*s) if cplx else rng.randn(*s)

    w, A = randn(M, k), randn(k, f)
    mask = np.rint(rng.uniform(0.3, 1.0, size=(M, f)))
    mask[::7] *= 0.5                                                  # not only 0 / 1 weights
    dev = torch.device('cuda', 0)
    Wd, Ad, Md = to_device2d(w, dev), to_device2d(A, dev), to_device2d(mask, dev)
    R = ops.make_rhs(rview(Ad), cplx, False)                          # NT operand of w . A: [f*cw, k*cw]
    out = empty2d(M, k, cplx, dev)
    out.fill_(float('nan'))
    ops.gemm_b2b_masked(rview(Wd), R, ops.epilogue(ops.EPI_STORE, rview(out), cwidth=cw, mask=Md))
    torch.cuda.synchronize()
    ref = (w.dot(A) * mask).dot(np.conj(A.T))
    err = np.max(np.abs(out.cpu().numpy() - ref)) / np.max(np.abs(ref))
    assert err <= 1e-13, err
