"""Typed Python wrappers over the C ABI (one function per exported symbol).

Every argument that is a matrix is a 2-D float64 CUDA tensor with unit inner stride and an
even row pitch ("real view"; complex data is passed as interleaved doubles, see
``_lib.rview``).  Nothing here computes: each function marshals pointers and calls the
shared library on torch's current stream.
"""
import ctypes

import torch

from . import _lib
from ._lib import (EPI_FLAG_COLVEC_IS_THRESHOLD, EPI_KL_RATIO, EPI_PROXQ, EPI_MU_DEN, EPI_MU_NUM, EPI_PROX, EPI_STORE, EPI_STORE_MASK,  # noqa: F401
                   SHRINK_COMPLEX, SHRINK_POSITIVE, SHRINK_REAL, Epilogue, ld, ptr, rview)
from ._device import empty2d


LAUNCHES = 0   # kernels of libdecomp_b200.so launched through this module (bench.py reports the count)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def probe_dmma_tflops():
    """Measured DMMA.8x8x4 issue rate of the current device (TFLOP/s); roofline denominator of the FP64 GEMMs."""
    v = ctypes.c_double(0.0)
    _lib.check(_lib.lib().decomp_probe_dmma_tflops(ctypes.byref(v)), 'decomp_probe_dmma_tflops')
    return v.value


def epilogue(kind, out, cwidth=1, **kw):
    """Build a decomp_epilogue_t. Keyword tensors: out2, x, other, prev, mask, colvec, colvec2, rowvec, step,
    latch, scratch; scalars: shrink, check, momentum, latch_value, flags."""
    e = Epilogue()
    e.kind = kind
    e.cwidth = cwidth
    e.shrink = kw.get('shrink', SHRINK_REAL)
    e.check = 1 if kw.get('check', False) else 0
    e.out, e.ldo = out.data_ptr(), ld(out)
    for name, ldname in (('out2', 'ldo2'), ('x', 'ldx'), ('other', 'ldother'), ('prev', 'ldprev'),
                         ('mask', 'ldmask')):
        t = kw.get(name)
        if t is not None:
            setattr(e, name, t.data_ptr())
            setattr(e, ldname, ld(t))
    for name in ('colvec', 'colvec2', 'rowvec', 'step', 'latch', 'scratch'):
        t = kw.get(name)
        if t is not None:
            setattr(e, name, t.data_ptr())
    e.momentum = float(kw.get('momentum', 0.0))
    e.latch_value = int(kw.get('latch_value', 0))
    e.flags = int(kw.get('flags', 0))
    return e


def gemm_nt(A, B, epi, skip=None):
    """acc = A . B^T  (A [M, K], B [N, K]) followed by the fused epilogue ``epi``."""
    M, K = A.shape
    N, K2 = B.shape
    assert K == K2, 'inner dimensions differ'
    rc = _lib.lib().decomp_gemm_nt_f64(_p(A), ld(A), _p(B), ld(B), M, N, K, ctypes.byref(epi), _p(skip),
                                       _lib.stream_ptr())
    _lib.check(rc, 'decomp_gemm_nt_f64')
    _count(1)


def gemm_b2b_masked_supported(K1):
    """True if the fused back-to-back masked GEMM covers this (real) code width."""
    return bool(_lib.lib().decomp_gemm_b2b_masked_supported(int(K1)))


def gemm_b2b_masked(W, R, epi, skip=None):
    """epilogue(((W . R^T) * mask) . R): W [M, K1], R [F, K1]; epi.mask [M, F / cwidth]; see decomp_b200.h."""
    M, K1 = W.shape
    F = R.shape[0]
    assert R.shape[1] == K1
    rc = _lib.lib().decomp_gemm_b2b_masked_f64(_p(W), ld(W), _p(R), ld(R), M, K1, F, ctypes.byref(epi), _p(skip),
                                               _lib.stream_ptr())
    _lib.check(rc, 'decomp_gemm_b2b_masked_f64')
    _count(1)


def gemm_tn_workspace(M, N, K, device):
    nbytes = _lib.lib().decomp_gemm_tn_workspace_bytes(M, N, K)
    return torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device=device)


def gemm_tn_workspace_for(shapes, device):
    """One workspace large enough for every (M, N, K) in ``shapes``."""
    nbytes = max(_lib.lib().decomp_gemm_tn_workspace_bytes(M, N, K) for M, N, K in shapes)
    return torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device=device)


def gemm_tn(A, B, out, combine=0, beta=0.0, workspace=None, skip=None):
    """out = [beta*out +] A^T . B  (A [K, M], B [K, N]); combine 2/3: conj(A)^T . B on interleaved complex."""
    K, M = A.shape
    K2, N = B.shape
    assert K == K2, 'row counts differ'
    if workspace is None:
        workspace = gemm_tn_workspace(M, N, K, A.device)
    rc = _lib.lib().decomp_gemm_tn_f64(_p(A), ld(A), _p(B), ld(B), M, N, K, _p(out), ld(out), combine, float(beta),
                                       _p(workspace), workspace.numel() * 8, _p(skip), _lib.stream_ptr())
    _lib.check(rc, 'decomp_gemm_tn_f64')
    _count(2)
    return out


def make_rhs(S, is_complex, conj_transpose, out=None, skip=None):
    """NT right-hand operand for X.S (conj_transpose=False) or X.S^H (True); S is the real view of [p, q]."""
    cw = 2 if is_complex else 1
    p, q = S.shape[0], S.shape[1] // cw
    rows, cols = (p * cw, q * cw) if conj_transpose else (q * cw, p * cw)
    if out is None:
        out = empty2d(rows, cols, False, S.device)
    rc = _lib.lib().decomp_make_rhs_f64(_p(S), ld(S), p, q, int(is_complex), int(conj_transpose), _p(out), ld(out),
                                        _p(skip), _lib.stream_ptr())
    _lib.check(rc, 'decomp_make_rhs_f64')
    _count(1)
    return out


def row_norms(A, is_complex, out=None):
    cw = 2 if is_complex else 1
    rows, cols = A.shape[0], A.shape[1] // cw
    if out is None:
        out = torch.empty(rows, dtype=torch.float64, device=A.device)
    rc = _lib.lib().decomp_row_norms_f64(_p(A), ld(A), rows, cols, int(is_complex), _p(out), _lib.stream_ptr())
    _lib.check(rc, 'decomp_row_norms_f64')
    _count(1)
    return out


def scale(A, out, cwidth=1, rowscale=None, invert_row=False, colscale=None, invert_col=False):
    rows, cols = A.shape
    rc = _lib.lib().decomp_scale_f64(_p(A), ld(A), rows, cols, cwidth, _p(rowscale), int(invert_row), _p(colscale),
                                     int(invert_col), _p(out), ld(out), _lib.stream_ptr())
    _lib.check(rc, 'decomp_scale_f64')
    _count(1)
    return out


def mask_mul(A, mask, out, cwidth=1):
    rows, cols = A.shape
    rc = _lib.lib().decomp_mask_mul_f64(_p(A), ld(A), _p(mask), ld(mask), rows, cols, cwidth, _p(out), ld(out),
                                        _lib.stream_ptr())
    _lib.check(rc, 'decomp_mask_mul_f64')
    _count(1)
    return out


def workspace(nbytes, device):
    """Caller-owned scratch for the entry points that take (workspace, workspace_bytes)."""
    return torch.empty(max((int(nbytes) + 7) // 8, 1), dtype=torch.float64, device=device)


def col_sums(A, scale_=1.0, out=None, ws=None):
    rows, cols = A.shape
    if out is None:
        out = torch.empty(cols, dtype=torch.float64, device=A.device)
    if ws is None:
        ws = workspace(_lib.lib().decomp_col_sums_workspace_bytes(rows, cols), A.device)
    rc = _lib.lib().decomp_col_sums_f64(_p(A), ld(A), rows, cols, float(scale_), _p(out), _p(ws), ws.numel() * 8,
                                        _lib.stream_ptr())
    _lib.check(rc, 'decomp_col_sums_f64')
    _count(2)
    return out


def row_sums(A, scale_=1.0, out=None):
    rows, cols = A.shape
    if out is None:
        out = torch.empty(rows, dtype=torch.float64, device=A.device)
    rc = _lib.lib().decomp_row_sums_f64(_p(A), ld(A), rows, cols, float(scale_), _p(out), _lib.stream_ptr())
    _lib.check(rc, 'decomp_row_sums_f64')
    _count(1)
    return out


def vector(k, device):
    """[k] float64 vector whose storage is readable up to an even element count (16-byte epilogue loads)."""
    return torch.zeros(k + (k & 1), dtype=torch.float64, device=device)[:k]


def gershgorin_step(G, is_complex, step_out, alpha_scaled=None, thr_out=None):
    cw = 2 if is_complex else 1
    k = G.shape[0]
    assert G.shape[1] == k * cw
    rc = _lib.lib().decomp_gershgorin_step_f64(_p(G), ld(G), k, int(is_complex), _p(alpha_scaled), _p(step_out),
                                               _p(thr_out), _lib.stream_ptr())
    _lib.check(rc, 'decomp_gershgorin_step_f64')
    _count(1)
    return step_out


def normalize_rows(D_in, D_out, is_complex, strict, D_ref=None, tol=0.0, latch=None, latch_value=0, maxdiff=None,
                   scratch=None, skip=None):
    cw = 2 if is_complex else 1
    rows, cols = D_in.shape[0], D_in.shape[1] // cw
    rc = _lib.lib().decomp_normalize_rows_f64(
        _p(D_in), ld(D_in), rows, cols, int(is_complex), int(strict), _p(D_out), ld(D_out), _p(D_ref),
        ld(D_ref) if D_ref is not None else 0, float(tol), _p(latch), int(latch_value), _p(maxdiff), _p(scratch),
        _p(skip), _lib.stream_ptr())
    _lib.check(rc, 'decomp_normalize_rows_f64')
    _count(1)
    return D_out


def gather_rows(src, index, out):
    rows, cols = out.shape
    rc = _lib.lib().decomp_gather_rows_f64(_p(src), ld(src), _p(index), rows, cols, _p(out), ld(out),
                                           _lib.stream_ptr())
    _lib.check(rc, 'decomp_gather_rows_f64')
    _count(1)
    return out


def scatter_rows(src, index, out):
    """out[index[r]] = src[r]."""
    rows, cols = src.shape
    rc = _lib.lib().decomp_scatter_rows_f64(_p(src), ld(src), _p(index), rows, cols, _p(out), ld(out), _lib.stream_ptr())
    _lib.check(rc, 'decomp_scatter_rows_f64')
    _count(1)
    return out


def dl_sweep_workspace(k, f, is_complex, device):
    return workspace(_lib.lib().decomp_dl_sweep_workspace_bytes(k, f, int(is_complex)), device)


def dl_sweep(S, T, D, is_complex, ws=None):
    cw = 2 if is_complex else 1
    k, f = D.shape[0], D.shape[1] // cw
    if ws is None:
        ws = dl_sweep_workspace(k, f, is_complex, D.device)
    rc = _lib.lib().decomp_dl_sweep_f64(_p(S), ld(S), _p(T), ld(T), _p(D), ld(D), k, f, int(is_complex), _p(ws),
                                        ws.numel() * 8, _lib.stream_ptr())
    _lib.check(rc, 'decomp_dl_sweep_f64')
    _count(1)
    return D


def dl_atom_weighted(X, is_complex, atom, W):
    cw = 2 if is_complex else 1
    rows, k = X.shape[0], X.shape[1] // cw
    rc = _lib.lib().decomp_dl_atom_weighted_f64(_p(X), ld(X), rows, k, int(is_complex), int(atom), _p(W), ld(W),
                                                _lib.stream_ptr())
    _lib.check(rc, 'decomp_dl_atom_weighted_f64')
    _count(1)
    return W


def dl_pair_products_t(Xt, is_complex, colA, colB, Wt):
    """Wt[c*cw + part, :] = parts of conj(x[:, colA[c]]) * x[:, colB[c]], from the transposed real view Xt of x."""
    rows, width = Xt.shape[1], colA.numel()
    rc = _lib.lib().decomp_dl_pair_products_t_f64(_p(Xt), ld(Xt), rows, int(is_complex), _p(colA), _p(colB), width,
                                                  _p(Wt), ld(Wt), _lib.stream_ptr())
    _lib.check(rc, 'decomp_dl_pair_products_t_f64')
    _count(1)
    return Wt


def dl_scatter_stats(P, is_complex, colA, colB, k, beta, S, slab_channels=0):
    """S[colA[c], j, colB[c]] = beta * S[...] + P[j, c]  (S: contiguous [k, f, k*cw] doubles, or channel slabs
    [f / slab_channels, k, slab_channels, k*cw])."""
    f, width = P.shape[0], colA.numel()
    rc = _lib.lib().decomp_dl_scatter_stats_f64(_p(P), ld(P), f, width, int(is_complex), _p(colA), _p(colB), k,
                                                float(beta), _p(S), int(slab_channels), _lib.stream_ptr())
    _lib.check(rc, 'decomp_dl_scatter_stats_f64')
    _count(1)
    return S


def dl_mirror(S, k, f, is_complex):
    """Fill S[b][j][a] = conj(S[a][j][b]) for b > a (S: contiguous [k, f, k*cw] doubles)."""
    rc = _lib.lib().decomp_dl_mirror_f64(_p(S), k, f, int(is_complex), _lib.stream_ptr())
    _lib.check(rc, 'decomp_dl_mirror_f64')
    _count(1)
    return S


def dl_masked_update(S, T, D, D_out, is_complex, workspace):
    cw = 2 if is_complex else 1
    k, f = D.shape[0], D.shape[1] // cw
    rc = _lib.lib().decomp_dl_masked_update_f64(_p(S), _p(T), ld(T), _p(D), ld(D), k, f, int(is_complex), _p(D_out),
                                                ld(D_out), _p(workspace), _lib.stream_ptr())
    _lib.check(rc, 'decomp_dl_masked_update_f64')
    _count(2)
    return D_out


def dl_masked_update_phase(phase, S_slab, j0, T, D, is_complex, D_slab_out, stats, workspace):
    """One phase (1, 2, 3) of the masked Jacobi update on the channel slab S_slab [k, fs, k*cw]; see decomp_b200.h."""
    cw = 2 if is_complex else 1
    k, f = D.shape[0], D.shape[1] // cw
    fs = S_slab.shape[1]
    rc = _lib.lib().decomp_dl_masked_update_phase_f64(int(phase), _p(S_slab), fs, int(j0), _p(T), ld(T), _p(D), ld(D), k, f,
                                                      int(is_complex), _p(D_slab_out), _p(stats), _p(workspace),
                                                      _lib.stream_ptr())
    _lib.check(rc, 'decomp_dl_masked_update_phase_f64')
    _count(2 if phase == 1 else 1)


def lasso_vectors(s, alpha, tol, mult=1.0, mult_dev=None):
    k = s.numel()
    alpha_out = vector(k, s.device)
    tol_out = vector(k, s.device)
    rc = _lib.lib().decomp_lasso_vectors_f64(_p(s), k, float(alpha), float(tol), float(mult), _p(mult_dev),
                                             _p(alpha_out), _p(tol_out), _lib.stream_ptr())
    _lib.check(rc, 'decomp_lasso_vectors_f64')
    _count(1)
    return alpha_out, tol_out


def axpby(a, X, b, Y, out):
    """out = a * X + b * Y (real views)."""
    rows, cols = X.shape
    rc = _lib.lib().decomp_axpby_f64(float(a), _p(X), ld(X), float(b), _p(Y), ld(Y), rows, cols, _p(out), ld(out),
                                     _lib.stream_ptr())
    _lib.check(rc, 'decomp_axpby_f64')
    _count(1)
    return out


def svrmu_update(D, P, Q, alpha, out):
    """out = max(D * ((1 - alpha) + alpha * P / max(Q, eps)), 0)."""
    rows, cols = D.shape
    rc = _lib.lib().decomp_svrmu_update_f64(_p(D), ld(D), _p(P), ld(P), _p(Q), ld(Q), float(alpha), rows, cols, _p(out),
                                            ld(out), _lib.stream_ptr())
    _lib.check(rc, 'decomp_svrmu_update_f64')
    _count(1)
    return out


def lasso_q(G, is_complex, step, Q):
    """Q = I - step * G (real views of [k, k] matrices; step is a device scalar)."""
    k = G.shape[0]
    rc = _lib.lib().decomp_lasso_q_f64(_p(G), ld(G), k, int(is_complex), _p(step), _p(Q), ld(Q), _lib.stream_ptr())
    _lib.check(rc, 'decomp_lasso_q_f64')
    _count(1)
    return Q


def scale_scalar(A, scalar_dev, out):
    """out = scalar * A with the scalar read from device memory."""
    rows, cols = A.shape
    rc = _lib.lib().decomp_scale_scalar_f64(_p(A), ld(A), rows, cols, _p(scalar_dev), _p(out), ld(out),
                                            _lib.stream_ptr())
    _lib.check(rc, 'decomp_scale_scalar_f64')
    _count(1)
    return out


def nmf_mu_small_supported(n, f, k, masked):
    return bool(_lib.lib().decomp_nmf_mu_small_supported(int(n), int(f), int(k), int(bool(masked))))


def nmf_mu_small(y, mask, X, D_in, D_out, sweeps, tol, it_out):
    """`sweeps` multiplicative updates of a small problem in ONE cooperative launch; see decomp_b200.h."""
    n, f = y.shape
    k = D_in.shape[0]
    ws = workspace(_lib.lib().decomp_nmf_mu_small_workspace_bytes(n, f, k, int(mask is not None)), y.device)
    rc = _lib.lib().decomp_nmf_mu_small_f64(_p(y), ld(y), _p(mask), ld(mask) if mask is not None else 0, _p(X), ld(X),
                                            _p(D_in), ld(D_in), _p(D_out), ld(D_out), n, f, k, int(sweeps), float(tol),
                                            _p(it_out), _p(ws), ws.numel() * 8, _lib.stream_ptr())
    _lib.check(rc, 'decomp_nmf_mu_small_f64')
    _count(1)
    return ws


def mu_update(x, num, den, out, skip=None):
    rows, cols = x.shape
    rc = _lib.lib().decomp_mu_update_f64(_p(x), ld(x), _p(num), ld(num), _p(den), ld(den), rows, cols, _p(out), ld(out),
                                         _p(skip), _lib.stream_ptr())
    _lib.check(rc, 'decomp_mu_update_f64')
    _count(1)
    return out


def max_abs_diff(A, B, is_complex, result, scratch, tol=0.0, latch=None, latch_value=0, skip=None):
    cw = 2 if is_complex else 1
    rows, cols = A.shape[0], A.shape[1] // cw
    rc = _lib.lib().decomp_max_abs_diff_f64(_p(A), ld(A), _p(B), ld(B), rows, cols, int(is_complex), float(tol),
                                            _p(latch), int(latch_value), _p(result), _p(scratch), _p(skip),
                                            _lib.stream_ptr())
    _lib.check(rc, 'decomp_max_abs_diff_f64')
    _count(1)
    return result


# ------------------------------------------------------------------------------------- TF32-split path
def empty_f32(rows, cols, device):
    """[rows, cols] float32 buffer with a row pitch that is a multiple of 4 floats (TMA: 16-byte rows)."""
    pitch = (cols + 3) // 4 * 4
    return torch.empty((max(rows, 1), max(pitch, 4)), dtype=torch.float32, device=device)[:rows, :cols]


def split_tf32(A, hi=None, lo=None):
    """A (float64 real view) ~= hi + lo with TF32-valued float32 pieces."""
    rows, cols = A.shape
    if hi is None:
        hi, lo = empty_f32(rows, cols, A.device), empty_f32(rows, cols, A.device)
    rc = _lib.lib().decomp_split_tf32_f64(_p(A), ld(A), rows, cols, _p(hi), _p(lo), ld(hi), _lib.stream_ptr())
    _lib.check(rc, 'decomp_split_tf32_f64')
    _count(1)
    return hi, lo


def gemm_nt_tf32x3(A_hi, A_lo, B_hi, B_lo, P, skip=None):
    """P = A . B^T in split TF32 on the tcgen05 tensor cores (A [M, K], B [N, K], P [M, N] float32)."""
    M, K = A_hi.shape
    N = B_hi.shape[0]
    rc = _lib.lib().decomp_gemm_nt_tf32x3(_p(A_hi), _p(A_lo), ld(A_hi), _p(B_hi), _p(B_lo), ld(B_hi), M, N, K, _p(P),
                                          ld(P), _p(skip), _lib.stream_ptr())
    _lib.check(rc, 'decomp_gemm_nt_tf32x3')
    _count(1)
    return P


TF32_K_PER_SPLIT = 4096   # FP32 accumulation length of the split-K statistics (rows per slab / K block)


def empty_f32_blocked(rows, cols, block, device, zero_tail=False):
    """K-blocked FP32 storage of a [cols, rows] (transposed) operand: [ceil(rows / block), cols, block]."""
    nblk = (rows + block - 1) // block
    t = torch.empty((nblk, cols, block), dtype=torch.float32, device=device)
    if zero_tail and rows % block:
        t[nblk - 1, :, rows % block:].zero_()
    return t


def split_transpose_tf32(A, hiT=None, loT=None, block=0):
    """TF32 pair of A^T from the float64 [rows, cols] matrix A: float32 [cols, rows], or with ``block`` > 0 the
    K-blocked [ceil(rows / block), cols, block] layout (tail of the last block zero-filled)."""
    rows, cols = A.shape
    if hiT is None:
        if block:
            hiT, loT = empty_f32_blocked(rows, cols, block, A.device), empty_f32_blocked(rows, cols, block, A.device)
        else:
            hiT, loT = empty_f32(cols, rows, A.device), empty_f32(cols, rows, A.device)
    rc = _lib.lib().decomp_split_transpose_tf32_f64(_p(A), ld(A), rows, cols, _p(hiT), _p(loT), 0 if block else ld(hiT),
                                                    block, _lib.stream_ptr())
    _lib.check(rc, 'decomp_split_transpose_tf32_f64')
    _count(1)
    return hiT, loT


def gemm_nt_tf32x3_splitk_workspace(M, N, K, device, k_per_split=TF32_K_PER_SPLIT):
    return workspace(_lib.lib().decomp_gemm_nt_tf32x3_splitk_workspace_bytes(M, N, K, k_per_split), device)


def gemm_nt_tf32x3_splitk(A_hi, A_lo, B_hi, B_lo, out, ws, k_per_split=TF32_K_PER_SPLIT, skip=None, K=None):
    """out (float64 [M, N]) = A . B^T in split TF32, the contraction cut into FP32-accumulated slabs summed in FP64.
    Operands: [rows, K] row-major, or (3-D tensors, ``K`` given) K-blocked [ceil(K / k_per_split), rows, k_per_split]."""
    blocked = A_hi.dim() == 3
    if blocked:
        assert K is not None and A_hi.shape[2] == k_per_split == B_hi.shape[2]
        M, N, lda, ldb = A_hi.shape[1], B_hi.shape[1], 0, 0
    else:
        M, K = A_hi.shape
        N, lda, ldb = B_hi.shape[0], ld(A_hi), ld(B_hi)
    rc = _lib.lib().decomp_gemm_nt_tf32x3_splitk_f64(_p(A_hi), _p(A_lo), lda, _p(B_hi), _p(B_lo), ldb, M, N, K,
                                                     k_per_split, int(blocked), _p(out), ld(out), _p(ws), ws.numel() * 8,
                                                     _p(skip), _lib.stream_ptr())
    _lib.check(rc, 'decomp_gemm_nt_tf32x3_splitk_f64')
    _count(2)
    return out


def nmf_xupdate_tf32x3(Y_hi, Y_lo, D_hi, D_lo, X, NEG, X_hi, X_lo, XT_hi, XT_lo, skip=None):
    """X <- X * max(Y D^T, 0) / max(NEG, eps) with the GEMM on tcgen05 (TF32 split); see decomp_b200.h.
    XT_hi / XT_lo: [k, n] float32, or K-blocked [ceil(n / block), k, block]."""
    n, f = Y_hi.shape
    k = D_hi.shape[0]
    xt_block = XT_hi.shape[2] if XT_hi.dim() == 3 else 0
    rc = _lib.lib().decomp_nmf_xupdate_tf32x3(_p(Y_hi), _p(Y_lo), ld(Y_hi), _p(D_hi), _p(D_lo), ld(D_hi), n, k, f, _p(X),
                                              ld(X), _p(NEG), ld(NEG), _p(X_hi), _p(X_lo), ld(X_hi), _p(XT_hi), _p(XT_lo),
                                              0 if xt_block else ld(XT_hi), xt_block, _p(skip), _lib.stream_ptr())
    _lib.check(rc, 'decomp_nmf_xupdate_tf32x3')
    _count(1)


def to_f32(A, out=None):
    """float32 copy of the float64 matrix A (the mask of the TF32-split masked path)."""
    rows, cols = A.shape
    if out is None:
        out = empty_f32(rows, cols, A.device)
    rc = _lib.lib().decomp_to_f32_f64(_p(A), ld(A), rows, cols, _p(out), ld(out), _lib.stream_ptr())
    _lib.check(rc, 'decomp_to_f32_f64')
    _count(1)
    return out


def gemm_nt_mask_tf32x3(A_hi, A_lo, B_hi, B_lo, mask32, F=None, FT=None, cwidth=1, skip=None):
    """F = (A . B^T) * mask32 on tcgen05 (TF32 split), written as the TF32 pair ``F`` = (hi, lo) row-major [M, N]
    and / or ``FT`` = (hi, lo) transposed: [N, M], or K-blocked [ceil(M / block), N, block]."""
    M, K = A_hi.shape
    N = B_hi.shape[0]
    Fh, Fl = F if F is not None else (None, None)
    FTh, FTl = FT if FT is not None else (None, None)
    blocked = FTh is not None and FTh.dim() == 3
    ft_block = FTh.shape[2] if blocked else 0
    ldft = 0 if (FTh is None or blocked) else ld(FTh)
    rc = _lib.lib().decomp_gemm_nt_mask_tf32x3(_p(A_hi), _p(A_lo), ld(A_hi), _p(B_hi), _p(B_lo), ld(B_hi), M, N, K,
                                               _p(mask32), ld(mask32) if mask32 is not None else 0, cwidth, _p(Fh), _p(Fl),
                                               ld(Fh) if Fh is not None else 0, _p(FTh), _p(FTl), ldft, ft_block,
                                               _p(skip), _lib.stream_ptr())
    _lib.check(rc, 'decomp_gemm_nt_mask_tf32x3')
    _count(1)


RESIDENT_MAX_ITERS = 32


def lasso_resident_supported(n_real):
    """True if the on-chip-resident Lasso kernel covers this (real) problem width."""
    return bool(_lib.lib().decomp_lasso_resident_supported(int(n_real)))


def lasso_resident(Q_rhs, M, epi, momentum, skip=None):
    """len(momentum) ISTA/FISTA iterations on M rows in one launch; epi.x = w (in/out), epi.out = x (in/out),
    epi.other = c."""
    N = Q_rhs.shape[0]
    mom = (ctypes.c_double * len(momentum))(*momentum)
    rc = _lib.lib().decomp_lasso_resident_f64(_p(Q_rhs), ld(Q_rhs), M, N, ctypes.byref(epi), len(momentum), mom,
                                              _p(skip), _lib.stream_ptr())
    _lib.check(rc, 'decomp_lasso_resident_f64')
    _count(1)


def prox_apply(P, epi, w_hi, w_lo, skip=None):
    """The masked iteration's FP64 pass: z = w + step (other - P), threshold / momentum / convergence as proxq_apply."""
    M, N = P.shape
    rc = _lib.lib().decomp_prox_apply_f64(_p(P), ld(P), ctypes.byref(epi), _p(w_hi), _p(w_lo), ld(w_hi), M, N,
                                          _p(skip), _lib.stream_ptr())
    _lib.check(rc, 'decomp_prox_apply_f64')
    _count(1)


def proxq_apply(P, epi, w_hi, w_lo, skip=None):
    """FP64 threshold / momentum / convergence pass after gemm_nt_tf32x3; writes w_next as (w_hi, w_lo)."""
    M, N = P.shape
    rc = _lib.lib().decomp_proxq_apply_f64(_p(P), ld(P), ctypes.byref(epi), _p(w_hi), _p(w_lo), ld(w_hi), M, N,
                                           _p(skip), _lib.stream_ptr())
    _lib.check(rc, 'decomp_proxq_apply_f64')
    _count(1)
