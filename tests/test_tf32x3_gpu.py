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
    A, y, mask, _ = gc._lasso_data((50,), 32, 40, 0)
    with pytest.raises(NotImplementedError):
        lasso.solve(y, A, 0.05, mask=mask, precision='tf32x3')          # masked iteration is FP64 only
    with pytest.raises(ValueError):
        lasso.solve(y, A, 0.05, precision='fp16')
