"""ctypes binding of libsst.so (include/sst.h).  Fails loudly when the library is missing."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsst.so")

F32, BF16 = 0, 1
GEMM_TN, GEMM_NT_MN = 0, 1
EPI_BIAS, EPI_RELU, EPI_DROPOUT, EPI_MULMASK, EPI_ACCUM = 1, 2, 4, 8, 16


class SstError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("out_dtype", C.c_int32), ("aux_dtype", C.c_int32), ("layout", C.c_int32),
                ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
                ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64), ("ldaux", C.c_int64),
                ("a_rows", C.c_int64), ("a_cols", C.c_int64), ("b_rows", C.c_int64), ("b_cols", C.c_int64),
                ("n_seg", C.c_int32), ("a_row_shift", C.c_int32 * 3), ("a_col0", C.c_int32 * 3),
                ("b_row_shift", C.c_int32 * 3), ("b_col0", C.c_int32 * 3),
                ("epilogue", C.c_int32), ("alpha", C.c_float), ("mask_scale", C.c_float), ("drop_p", C.c_float),
                ("seed", C.c_uint64), ("remap_P", C.c_int32), ("remap_T", C.c_int32), ("remap_j0", C.c_int32),
                ("force_simt", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SstError("libsst.so not built (%s); run __graft_entry__.build() -- there is no fallback path" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.sst_version.restype = C.c_char_p
        _lib.sst_last_error.restype = C.c_char_p
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise SstError("%s failed (%d): %s" % (what, rc, lib().sst_last_error().decode()))


def dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise SstError("unsupported dtype %s" % t.dtype)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_device():
    check(lib().sst_device_check(), "sst_device_check")


def gemm(A, B, Cout, M, N, K, lda, ldb, ldc, layout=GEMM_TN, bias=None, aux=None, ldaux=0, epilogue=0, alpha=1.0,
         mask_scale=1.0, drop_p=0.0, seed=0, n_seg=1, a_row_shift=(0, 0, 0), a_col0=(0, 0, 0),
         b_row_shift=(0, 0, 0), b_col0=(0, 0, 0), a_rows=None, a_cols=None, b_rows=None, b_cols=None,
         remap=(0, 0, 0), force_simt=False):
    d = GemmDesc()
    d.dtype, d.out_dtype, d.layout = dt(A), dt(Cout), layout
    d.aux_dtype = dt(aux) if aux is not None else F32
    d.M, d.N, d.K = M, N, K
    d.lda, d.ldb, d.ldc, d.ldaux = lda, ldb, ldc, ldaux
    if layout == GEMM_TN:
        d.a_rows = a_rows if a_rows is not None else M
        d.a_cols = a_cols if a_cols is not None else K
        d.b_rows, d.b_cols = N, K
    else:
        d.a_rows = a_rows if a_rows is not None else K
        d.a_cols = a_cols if a_cols is not None else M
        d.b_rows = b_rows if b_rows is not None else K
        d.b_cols = b_cols if b_cols is not None else N
    d.n_seg = n_seg
    for i in range(3):
        d.a_row_shift[i], d.a_col0[i] = a_row_shift[i], a_col0[i]
        d.b_row_shift[i], d.b_col0[i] = b_row_shift[i], b_col0[i]
    d.epilogue, d.alpha, d.mask_scale, d.drop_p, d.seed = epilogue, alpha, mask_scale, drop_p, seed
    d.remap_P, d.remap_T, d.remap_j0 = remap
    d.force_simt = 1 if force_simt else 0
    check(lib().sst_gemm(C.byref(d), ptr(A), ptr(B), ptr(Cout), ptr(bias), ptr(aux), stream()), "sst_gemm")
