"""ctypes binding of libdecomp_b200.so (the C ABI declared in include/decomp_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is
raised.  torch is used only to own device memory and to name the CUDA stream.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'lib', 'libdecomp_b200.so')

EPI_STORE, EPI_STORE_MASK, EPI_MU_NUM, EPI_MU_DEN, EPI_PROX, EPI_KL_RATIO, EPI_PROXQ = range(7)
SHRINK_REAL, SHRINK_COMPLEX, SHRINK_POSITIVE = range(3)
EPI_FLAG_COLVEC_IS_THRESHOLD = 1

c_dp = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32


class Epilogue(ctypes.Structure):
    """Mirror of decomp_epilogue_t."""
    _fields_ = [
        ('kind', c_i32), ('shrink', c_i32), ('cwidth', c_i32), ('check', c_i32),
        ('out', c_dp), ('ldo', c_i64),
        ('out2', c_dp), ('ldo2', c_i64),
        ('x', c_dp), ('ldx', c_i64),
        ('other', c_dp), ('ldother', c_i64),
        ('prev', c_dp), ('ldprev', c_i64),
        ('mask', c_dp), ('ldmask', c_i64),
        ('colvec', c_dp), ('colvec2', c_dp), ('rowvec', c_dp), ('step', c_dp),
        ('momentum', ctypes.c_double),
        ('latch', c_dp), ('scratch', c_dp),
        ('latch_value', c_i32), ('flags', c_i32),
    ]


class DecompError(RuntimeError):
    pass


_lib = None


def _declare(lib):
    lib.decomp_last_error.restype = ctypes.c_char_p
    lib.decomp_abi_version.restype = ctypes.c_int
    lib.decomp_probe_dmma_tflops.argtypes = [ctypes.POINTER(ctypes.c_double)]
    lib.decomp_gemm_nt_f64.argtypes = [c_dp, c_i64, c_dp, c_i64, c_i64, c_i64, c_i64,
                                       ctypes.POINTER(Epilogue), c_dp, c_dp]
    lib.decomp_gemm_b2b_masked_supported.argtypes = [c_i64]
    lib.decomp_gemm_b2b_masked_f64.argtypes = [c_dp, c_i64, c_dp, c_i64, c_i64, c_i64, c_i64, ctypes.POINTER(Epilogue), c_dp,
                                               c_dp]
    lib.decomp_gemm_tn_workspace_bytes.argtypes = [c_i64, c_i64, c_i64]
    lib.decomp_gemm_tn_workspace_bytes.restype = ctypes.c_size_t
    lib.decomp_gemm_tn_f64.argtypes = [c_dp, c_i64, c_dp, c_i64, c_i64, c_i64, c_i64, c_dp, c_i64, c_i32,
                                       ctypes.c_double, c_dp, ctypes.c_size_t, c_dp, c_dp]
    lib.decomp_make_rhs_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, c_i32, c_i32, c_dp, c_i64, c_dp, c_dp]
    lib.decomp_row_norms_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, c_i32, c_dp, c_dp]
    lib.decomp_scale_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, c_i32, c_dp, c_i32, c_dp, c_i32, c_dp, c_i64, c_dp]
    lib.decomp_mask_mul_f64.argtypes = [c_dp, c_i64, c_dp, c_i64, c_i64, c_i64, c_i32, c_dp, c_i64, c_dp]
    lib.decomp_col_sums_workspace_bytes.argtypes = [c_i64, c_i64]
    lib.decomp_col_sums_workspace_bytes.restype = ctypes.c_size_t
    lib.decomp_col_sums_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, ctypes.c_double, c_dp, c_dp, ctypes.c_size_t, c_dp]
    lib.decomp_row_sums_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, ctypes.c_double, c_dp, c_dp]
    lib.decomp_gershgorin_step_f64.argtypes = [c_dp, c_i64, c_i64, c_i32, c_dp, c_dp, c_dp, c_dp]
    lib.decomp_normalize_rows_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, c_i32, c_i32, c_dp, c_i64, c_dp, c_i64,
                                              ctypes.c_double, c_dp, c_i32, c_dp, c_dp, c_dp, c_dp]
    lib.decomp_gather_rows_f64.argtypes = [c_dp, c_i64, c_dp, c_i64, c_i64, c_dp, c_i64, c_dp]
    lib.decomp_scatter_rows_f64.argtypes = [c_dp, c_i64, c_dp, c_i64, c_i64, c_dp, c_i64, c_dp]
    lib.decomp_lasso_vectors_f64.argtypes = [c_dp, c_i64, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_dp, c_dp,
                                             c_dp, c_dp]
    lib.decomp_axpby_f64.argtypes = [ctypes.c_double, c_dp, c_i64, ctypes.c_double, c_dp, c_i64, c_i64, c_i64, c_dp, c_i64,
                                     c_dp]
    lib.decomp_svrmu_update_f64.argtypes = [c_dp, c_i64, c_dp, c_i64, c_dp, c_i64, ctypes.c_double, c_i64, c_i64, c_dp,
                                            c_i64, c_dp]
    lib.decomp_lasso_q_f64.argtypes = [c_dp, c_i64, c_i64, c_i32, c_dp, c_dp, c_i64, c_dp]
    lib.decomp_scale_scalar_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, c_dp, c_dp, c_i64, c_dp]
    lib.decomp_split_tf32_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, c_dp, c_dp, c_i64, c_dp]
    lib.decomp_gemm_nt_tf32x3.argtypes = [c_dp, c_dp, c_i64, c_dp, c_dp, c_i64, c_i64, c_i64, c_i64, c_dp, c_i64, c_dp,
                                          c_dp]
    lib.decomp_gemm_nt_tf32x3_splitk_workspace_bytes.argtypes = [c_i64, c_i64, c_i64, c_i64]
    lib.decomp_gemm_nt_tf32x3_splitk_workspace_bytes.restype = ctypes.c_size_t
    lib.decomp_gemm_nt_tf32x3_splitk_f64.argtypes = [c_dp, c_dp, c_i64, c_dp, c_dp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i32,
                                                     c_dp, c_i64, c_dp, ctypes.c_size_t, c_dp, c_dp]
    lib.decomp_nmf_xupdate_tf32x3.argtypes = [c_dp, c_dp, c_i64, c_dp, c_dp, c_i64, c_i64, c_i64, c_i64, c_dp, c_i64, c_dp,
                                              c_i64, c_dp, c_dp, c_i64, c_dp, c_dp, c_i64, c_i64, c_dp, c_dp]
    lib.decomp_gemm_nt_mask_tf32x3.argtypes = [c_dp, c_dp, c_i64, c_dp, c_dp, c_i64, c_i64, c_i64, c_i64, c_dp, c_i64, c_i32,
                                               c_dp, c_dp, c_i64, c_dp, c_dp, c_i64, c_i64, c_dp, c_dp]
    lib.decomp_to_f32_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, c_dp, c_i64, c_dp]
    lib.decomp_split_transpose_tf32_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, c_dp, c_dp, c_i64, c_i64, c_dp]
    lib.decomp_proxq_apply_f64.argtypes = [c_dp, c_i64, ctypes.POINTER(Epilogue), c_dp, c_dp, c_i64, c_i64, c_i64, c_dp,
                                           c_dp]
    lib.decomp_prox_apply_f64.argtypes = lib.decomp_proxq_apply_f64.argtypes
    lib.decomp_lasso_resident_supported.argtypes = [c_i64]
    lib.decomp_lasso_resident_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, ctypes.POINTER(Epilogue), c_i32,
                                              ctypes.POINTER(ctypes.c_double), c_dp, c_dp]
    lib.decomp_mu_update_f64.argtypes = [c_dp, c_i64, c_dp, c_i64, c_dp, c_i64, c_i64, c_i64, c_dp, c_i64, c_dp, c_dp]
    lib.decomp_max_abs_diff_f64.argtypes = [c_dp, c_i64, c_dp, c_i64, c_i64, c_i64, c_i32, ctypes.c_double, c_dp, c_i32,
                                            c_dp, c_dp, c_dp, c_dp]
    lib.decomp_dl_sweep_workspace_bytes.argtypes = [c_i64, c_i64, c_i32]
    lib.decomp_dl_sweep_workspace_bytes.restype = ctypes.c_size_t
    lib.decomp_dl_sweep_f64.argtypes = [c_dp, c_i64, c_dp, c_i64, c_dp, c_i64, c_i64, c_i64, c_i32, c_dp, ctypes.c_size_t, c_dp]
    lib.decomp_dl_atom_weighted_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, c_i32, c_i64, c_dp, c_i64, c_dp]
    lib.decomp_dl_pair_products_t_f64.argtypes = [c_dp, c_i64, c_i64, c_i32, c_dp, c_dp, c_i64, c_dp, c_i64, c_dp]
    lib.decomp_dl_scatter_stats_f64.argtypes = [c_dp, c_i64, c_i64, c_i64, c_i32, c_dp, c_dp, c_i64, ctypes.c_double, c_dp,
                                                c_i64, c_dp]
    lib.decomp_dl_masked_update_phase_f64.argtypes = [c_i32, c_dp, c_i64, c_i64, c_dp, c_i64, c_dp, c_i64, c_i64, c_i64,
                                                      c_i32, c_dp, c_dp, c_dp, c_dp]
    lib.decomp_dl_mirror_f64.argtypes = [c_dp, c_i64, c_i64, c_i32, c_dp]
    lib.decomp_dl_masked_update_f64.argtypes = [c_dp, c_dp, c_i64, c_dp, c_i64, c_i64, c_i64, c_i32, c_dp, c_i64,
                                                c_dp, c_dp]
    lib.decomp_nmf_mu_small_supported.argtypes = [c_i64, c_i64, c_i64, c_i32]
    lib.decomp_nmf_mu_small_workspace_bytes.argtypes = [c_i64, c_i64, c_i64, c_i32]
    lib.decomp_nmf_mu_small_workspace_bytes.restype = ctypes.c_size_t
    lib.decomp_nmf_mu_small_f64.argtypes = [c_dp, c_i64, c_dp, c_i64, c_dp, c_i64, c_dp, c_i64, c_dp, c_i64, c_i64, c_i64, c_i64,
                                            c_i32, ctypes.c_double, c_dp, c_dp, ctypes.c_size_t, c_dp]
    lib.decomp_comm_unique_id.argtypes = [c_dp]
    lib.decomp_comm_init.argtypes = [c_dp, c_i32, c_i32, ctypes.POINTER(ctypes.c_void_p)]
    lib.decomp_comm_destroy.argtypes = [c_dp]
    lib.decomp_comm_allreduce_sum_f64.argtypes = [c_dp, c_dp, c_i64, c_dp]
    lib.decomp_comm_allreduce_min_i32.argtypes = [c_dp, c_dp, c_i64, c_dp]
    lib.decomp_comm_reduce_scatter_sum_f64.argtypes = [c_dp, c_dp, c_dp, c_i64, c_dp]
    lib.decomp_comm_allgather_f64.argtypes = [c_dp, c_dp, c_dp, c_i64, c_dp]
    lib.decomp_staged_upload.argtypes = [c_dp, c_dp, ctypes.c_size_t, c_dp, ctypes.c_size_t, c_i32, c_i32, c_dp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is ctypes.c_int and name not in ('decomp_abi_version',):
            fn.restype = ctypes.c_int


EXPORTS = (
    'decomp_last_error', 'decomp_abi_version', 'decomp_probe_dmma_tflops', 'decomp_gemm_nt_f64', 'decomp_gemm_b2b_masked_supported', 'decomp_gemm_b2b_masked_f64',
    'decomp_gemm_tn_workspace_bytes',
    'decomp_gemm_tn_f64', 'decomp_make_rhs_f64', 'decomp_row_norms_f64', 'decomp_scale_f64', 'decomp_mask_mul_f64',
    'decomp_col_sums_workspace_bytes', 'decomp_col_sums_f64', 'decomp_row_sums_f64', 'decomp_gershgorin_step_f64', 'decomp_normalize_rows_f64',
    'decomp_gather_rows_f64', 'decomp_scatter_rows_f64', 'decomp_lasso_vectors_f64', 'decomp_axpby_f64', 'decomp_svrmu_update_f64', 'decomp_lasso_q_f64', 'decomp_scale_scalar_f64', 'decomp_mu_update_f64', 'decomp_max_abs_diff_f64',
    'decomp_dl_sweep_workspace_bytes', 'decomp_dl_sweep_f64', 'decomp_dl_atom_weighted_f64', 'decomp_dl_pair_products_t_f64', 'decomp_dl_scatter_stats_f64', 'decomp_dl_mirror_f64', 'decomp_dl_masked_update_f64', 'decomp_dl_masked_update_phase_f64',
    'decomp_split_tf32_f64', 'decomp_gemm_nt_tf32x3', 'decomp_proxq_apply_f64',
    'decomp_gemm_nt_tf32x3_splitk_workspace_bytes', 'decomp_gemm_nt_tf32x3_splitk_f64', 'decomp_nmf_xupdate_tf32x3',
    'decomp_split_transpose_tf32_f64', 'decomp_gemm_nt_mask_tf32x3', 'decomp_to_f32_f64', 'decomp_prox_apply_f64',
    'decomp_lasso_resident_supported', 'decomp_lasso_resident_f64', 'decomp_staged_upload',
    'decomp_comm_unique_id', 'decomp_comm_init', 'decomp_comm_destroy', 'decomp_comm_allreduce_sum_f64',
    'decomp_comm_allreduce_min_i32', 'decomp_comm_reduce_scatter_sum_f64', 'decomp_comm_allgather_f64',
    'decomp_nmf_mu_small_supported', 'decomp_nmf_mu_small_workspace_bytes', 'decomp_nmf_mu_small_f64',
)


def lib():
    """The loaded shared library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DecompError(
                'libdecomp_b200.so is missing (%s). Build it with `python -m decomp_b200._build` '
                '(needs nvcc); decomp_b200 has no CPU fallback.' % LIB_PATH)
        loaded = ctypes.CDLL(LIB_PATH)
        _declare(loaded)
        _lib = loaded
    return _lib


def check(rc, what):
    if rc != 0:
        raise DecompError('%s failed (%d): %s' % (what, rc, lib().decomp_last_error().decode()))


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def rview(t):
    """Real [rows, cols*cw] view of a 2-D real or complex device tensor with unit inner stride."""
    if t.is_complex():
        r = torch.view_as_real(t)            # [rows, cols, 2]
        return r.as_strided((t.shape[0], t.shape[1] * 2), (r.stride(0), 1), r.storage_offset())
    return t


def ld(t):
    """Leading dimension in doubles of a 2-D (real view) tensor."""
    assert t.dim() == 2 and (t.shape[1] == 1 or t.stride(1) == 1), 'inner stride must be 1'
    if t.shape[0] > 1:
        return t.stride(0)
    n = max(t.stride(0), t.shape[1])     # single row: any even pitch >= cols will do
    return n + (n & 1)
