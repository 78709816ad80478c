"""sst_gemm (include/sst.h "GEMM family") on a B200 against torch fp32 matmul / conv1d on the same (bf16-rounded) operands:
TN and MN-major layouts, conv-tap K segments with halo-row remap, split-K weight gradients, every epilogue bit, both tile
widths, and the CUDA-core twin.  Tolerances: fp32 kernel 1e-5, bf16 operands with fp32 output 2e-3, bf16 output 2e-2
(one bf16 rounding of the result), all relative to max|ref|."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def L():
    import sst_b200  # noqa: F401
    from sst_b200 import lib
    lib.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return lib


def rel(got, ref):
    got, ref = got.double(), ref.double()
    assert bool(torch.isfinite(got).all())
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-30))


def tol_for(dtype, out_dtype):
    if out_dtype == torch.bfloat16:
        return 2e-2
    return 1e-5 if dtype == torch.float32 else 2e-3


TN_SHAPES = [(128, 128, 64), (128, 256, 64), (256, 512, 768), (800, 768, 768), (1000, 2304, 768), (333, 200, 136),
             (7744, 768, 768), (4100, 3072, 768)]


@pytest.mark.parametrize("M,N,K", TN_SHAPES)
@pytest.mark.parametrize("epi_name", ["none", "bias_relu", "bias_accum"])
def test_tn_tensor_core(L, M, N, K, epi_name):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=DEV, generator=g) * 0.5).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    epi = {"none": 0, "bias_relu": L.EPI_BIAS | L.EPI_RELU, "bias_accum": L.EPI_BIAS | L.EPI_ACCUM}[epi_name]
    for out_dtype in (torch.bfloat16, torch.float32):
        C = torch.full((M, N), 7.0, device=DEV, dtype=out_dtype)
        ref = A.float() @ B.float().t()
        if epi & L.EPI_BIAS:
            ref = ref + bias
        if epi & L.EPI_RELU:
            ref = ref.relu()
        if epi & L.EPI_ACCUM:
            ref = ref + 7.0
        L.gemm(A, B, C, M, N, K, K, K, N, bias=bias, epilogue=epi)
        assert rel(C, ref) < tol_for(torch.bfloat16, out_dtype), (out_dtype, epi_name)


@pytest.mark.parametrize("M,N,K", [(800, 768, 3072), (7744, 768, 2304), (333, 768, 44), (1000, 200, 136)])
@pytest.mark.parametrize("simt", [False, True])
def test_tn_with_mn_major_b(L, M, N, K, simt):
    """SST_GEMM_TN_BMN: C = A[M,K] . B[K,N] with B row-major (K, N) -- dx = dy . W on the weight's own layout; K = 44 is the
    CTC head (dy pitch 64, rows of W beyond K zero-filled by TMA)."""
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    lda = (K + 63) // 64 * 64
    Aw = (torch.randn(M, lda, device=DEV, generator=g) * 0.5).bfloat16()
    B = (torch.randn(K, N, device=DEV, generator=g) * 0.5).bfloat16()
    aux = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    for epi in (0, L.EPI_ACCUM, L.EPI_MULMASK):
        C = torch.full((M, N), 2.0, device=DEV, dtype=torch.bfloat16)
        L.gemm(Aw, B, C, M, N, K, lda, N, N, layout=L.GEMM_TN_BMN, a_cols=K, aux=aux if epi == L.EPI_MULMASK else None, ldaux=N,
               epilogue=epi, mask_scale=1.25, force_simt=simt)
        ref = Aw[:, :K].float() @ B.float()
        if epi == L.EPI_ACCUM:
            ref = ref + 2.0
        if epi == L.EPI_MULMASK:
            ref = ref * torch.where(aux.float() > 0, 1.25, 0.0)
        assert rel(C, ref) < 2e-2, (epi, simt)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_tn_cuda_core(L, dtype):
    M, N, K = 300, 136, 72
    A = (torch.randn(M, K, device=DEV) * 0.5).to(dtype)
    B = (torch.randn(N, K, device=DEV) * 0.5).to(dtype)
    bias = torch.randn(N, device=DEV)
    C = torch.full((M, N), 7.0, device=DEV, dtype=dtype)
    L.gemm(A, B, C, M, N, K, K, K, N, bias=bias, epilogue=L.EPI_BIAS | L.EPI_RELU | L.EPI_ACCUM, force_simt=True)
    ref = (A.float() @ B.float().t() + bias).relu() + 7.0
    assert rel(C, ref) < tol_for(dtype, dtype)


def test_alpha_and_padded_pitches(L):
    """alpha scaling, lda/ldb/ldc larger than the logical extents (views into wider matrices)."""
    M, N, K = 500, 96, 768
    Aw = (torch.randn(M, 3 * K, device=DEV) * 0.5).bfloat16()
    Bw = (torch.randn(N, K + 64, device=DEV) * 0.5).bfloat16()
    Cw = torch.zeros(M, 256, device=DEV, dtype=torch.float32)
    A, B = Aw[:, K:2 * K], Bw[:, :K]
    L.gemm(A, B, Cw, M, N, K, 3 * K, K + 64, 256, alpha=0.25, a_cols=3 * K)
    assert rel(Cw[:, :N], 0.25 * (A.float() @ B.float().t())) < 2e-3
    assert float(Cw[:, N:].abs().max()) == 0.0


@pytest.mark.parametrize("simt", [False, True])
def test_dropout_epilogue_keep_rate_and_stream(L, simt):
    """DROPOUT epilogue: v = keep ? v/(1-p) : 0 with keep a pure function of (seed, m*N+n): the tensor-core and the
    CUDA-core kernel draw the same mask; keep-rate within 1 % of 1-p."""
    M, N, K, p = 1024, 768, 128, 0.2
    A = (torch.rand(M, K, device=DEV) + 0.5).bfloat16()
    B = (torch.rand(N, K, device=DEV) + 0.5).bfloat16()
    ref = A.float() @ B.float().t()
    C = torch.empty(M, N, device=DEV, dtype=torch.float32)
    L.gemm(A, B, C, M, N, K, K, K, N, epilogue=L.EPI_DROPOUT, drop_p=p, seed=1234, force_simt=simt)
    kept = C != 0
    assert abs(float(kept.float().mean()) - (1 - p)) < 0.01
    from helpers import philox_keep_mask16
    assert bool((kept.cpu() == philox_keep_mask16(1234, M * N, p).view(M, N)).all()), "mask differs from the host Philox mirror"
    assert rel(C[kept], (ref / (1 - p))[kept]) < 2e-3
    C2 = torch.empty_like(C)
    L.gemm(A, B, C2, M, N, K, K, K, N, epilogue=L.EPI_DROPOUT, drop_p=p, seed=1234, force_simt=not simt)
    assert bool(((C2 != 0) == kept).all()), "tensor-core and CUDA-core kernels must draw the same dropout mask"
    C3 = torch.empty_like(C)
    L.gemm(A, B, C3, M, N, K, K, K, N, epilogue=L.EPI_DROPOUT, drop_p=p, seed=99, force_simt=simt)
    assert 0.1 < float(((C3 != 0) != kept).float().mean()) < 0.5      # another seed, another mask


@pytest.mark.parametrize("aux_dtype", [torch.bfloat16, torch.float32])
def test_mulmask_epilogue(L, aux_dtype):
    """MULMASK: v *= (aux > 0 ? mask_scale : 0) -- the ReLU/dropout backward of the FFN (transformer.py:61)."""
    M, N, K = 600, 3072, 768
    A = (torch.randn(M, K, device=DEV) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=DEV) * 0.5).bfloat16()
    aux = torch.randn(M, N, device=DEV).to(aux_dtype)
    C = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    L.gemm(A, B, C, M, N, K, K, K, N, aux=aux, ldaux=N, epilogue=L.EPI_MULMASK, mask_scale=1.25)
    ref = (A.float() @ B.float().t()) * torch.where(aux.float() > 0, 1.25, 0.0)
    assert rel(C, ref) < 2e-2


@pytest.mark.parametrize("Kr,M,N", [(64, 128, 256), (800, 768, 768), (6400, 768, 3072), (1000, 3072, 768), (777, 200, 328),
                                      (7744, 768, 768), (64000, 768, 2304)])
@pytest.mark.parametrize("accum", [False, True])
def test_mn_major_weight_gradient(L, Kr, M, N, accum):
    """C (+)= A[K,M]^T B[K,N] (split-K, fp32 atomics)."""
    A = (torch.randn(Kr, M, device=DEV) * 0.5).bfloat16()
    B = (torch.randn(Kr, N, device=DEV) * 0.5).bfloat16()
    C = torch.full((M, N), 3.0 if accum else 99.0, device=DEV, dtype=torch.float32)
    L.gemm(A, B, C, M, N, Kr, M, N, N, layout=L.GEMM_NT_MN, epilogue=L.EPI_ACCUM if accum else 0)
    ref = A.float().t() @ B.float() + (3.0 if accum else 0.0)
    assert rel(C, ref) < 2e-3


def _conv_inputs(n, T, Cin, Cout, dtype):
    x = (torch.randn(n, T, Cin, device=DEV) * 0.5).to(dtype)
    w = (torch.randn(Cout, Cin, 3, device=DEV) * 0.1).to(dtype)
    xp = torch.zeros(n, T + 2, Cin, device=DEV, dtype=dtype)
    xp[:, 1:T + 1] = x
    wp = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).contiguous()      # [o][tap*Cin + c]
    return x, w, xp, wp


@pytest.mark.parametrize("n,T,Cin,Cout,stride,dtype,simt", [
    (4, 200, 768, 768, 1, torch.bfloat16, False), (4, 400, 768, 768, 2, torch.bfloat16, False),
    (3, 40, 64, 96, 2, torch.bfloat16, False), (3, 40, 64, 96, 1, torch.float32, True), (3, 40, 64, 96, 2, torch.float32, True)])
def test_conv3_as_segmented_gemm(L, n, T, Cin, Cout, stride, dtype, simt):
    """Conv1d(k=3, pad=1, stride 1|2) (architecture.py:26,28) over the time-padded channels-last layout."""
    x, w, xp, wp = _conv_inputs(n, T, Cin, Cout, dtype)
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), stride=stride, padding=1).transpose(1, 2)
    To = ref.shape[1]
    out = torch.full((n * To, Cout), 5.0, device=DEV, dtype=dtype)
    if stride == 1:
        P = T + 2
        L.gemm(xp, wp, out, n * P, Cout, 3 * Cin, Cin, 3 * Cin, Cout, n_seg=3, a_row_shift=(-1, 0, 1), a_col0=(0, 0, 0),
               a_rows=n * P, a_cols=Cin, remap=(P, To, 1), force_simt=simt)
    else:
        P = T // 2 + 1
        L.gemm(xp, wp, out, n * P, Cout, 3 * Cin, 2 * Cin, 3 * Cin, Cout, n_seg=3, a_row_shift=(0, 0, 1), a_col0=(0, Cin, 0),
               a_rows=n * P, a_cols=2 * Cin, remap=(P, To, 0), force_simt=simt)
    # fp32 CUDA-core kernel vs cuDNN/oneDNN-free torch conv on GPU: accumulation order differs, 1e-4 is the parity bar
    assert rel(out.view(n * To, Cout), ref.reshape(n * To, Cout)) < (2e-2 if dtype == torch.bfloat16 else 1e-4)


@pytest.mark.parametrize("simt", [False, True])
def test_conv3_weight_gradient(L, simt):
    n, T, Cin, Cout = (4, 200, 768, 768) if not simt else (3, 40, 256, 128)
    P = T + 2
    x = torch.zeros(n, P, Cin, device=DEV); x[:, 1:T + 1] = torch.randn(n, T, Cin, device=DEV) * 0.5
    dy = torch.zeros(n, P, Cout, device=DEV); dy[:, 1:T + 1] = torch.randn(n, T, Cout, device=DEV) * 0.5
    x, dy = x.bfloat16(), dy.bfloat16()
    xf, dyf = x.float().view(n * P, Cin), dy.float().view(n * P, Cout)
    ref = torch.zeros(Cout, 3 * Cin, device=DEV)
    for tap in range(3):
        xs = torch.zeros_like(xf)
        if tap == 0:
            xs[1:] = xf[:-1]
        elif tap == 2:
            xs[:-1] = xf[1:]
        else:
            xs = xf
        ref[:, tap * Cin:(tap + 1) * Cin] = dyf.t() @ xs
    out = torch.zeros(Cout, 3 * Cin, device=DEV)
    L.gemm(dy, x, out, Cout, 3 * Cin, n * P, Cout, Cin, 3 * Cin, layout=L.GEMM_NT_MN, n_seg=3, b_row_shift=(-1, 0, 1),
           b_col0=(0, 0, 0), a_rows=n * P, a_cols=Cout, b_rows=n * P, b_cols=Cin, epilogue=L.EPI_ACCUM, force_simt=simt)
    assert rel(out, ref) < 3e-3


def test_argument_errors_are_reported(L):
    A = torch.zeros(128, 72, device=DEV, dtype=torch.bfloat16)     # row pitch 144 B: not a multiple of 16 -> TMA cannot map it
    B = torch.zeros(128, 72, device=DEV, dtype=torch.bfloat16)
    C = torch.zeros(128, 128, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(L.SstError):
        L.gemm(A[:, :68], B[:, :68], C, 128, 128, 68, 68, 68, 128)
    with pytest.raises(L.SstError):
        L.gemm(A, B, C, 128, 128, 72, 72, 72, 128, epilogue=L.EPI_BIAS)      # BIAS without a bias pointer


@pytest.mark.parametrize("M,N,K,grp", [(800, 768, 768, 0), (4100, 3072, 768, 0), (7744, 1536, 64, 768), (333, 256, 136, 128)])
@pytest.mark.parametrize("mode", [1, 2])
def test_fused_column_accumulators(L, M, N, K, grp, mode):
    """SstGemmDesc.col_acc: the bias gradient (mode 1: float column sums, ACCUMULATED) / the BatchNorm batch statistics (mode 2: double
    sums and sums of squares per group of `grp` channels, zeroed by the call) of the stored bf16 result leave from the GEMM's
    epilogue -- they must be what a separate pass over the stored C gives.  Also with a masked epilogue and with a row remap that
    drops halo rows (the dropped rows must not be counted)."""
    g = torch.Generator(device=DEV).manual_seed(M + N)
    A = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=DEV, generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    for remap in ((0, 0, 0), (10, 8, 1)):
        if remap[0] and M % remap[0]:
            continue
        rows_out = M if not remap[0] else M // remap[0] * remap[1]
        C = torch.full((rows_out, N), 7.0, device=DEV, dtype=torch.bfloat16)
        if mode == 1:
            acc = torch.full((N,), 3.0, device=DEV)                    # accumulated on top of what is there
        else:
            acc = torch.full((2 * N,), 123.0, device=DEV, dtype=torch.float64)      # zeroed by the call
        L.gemm(A, B, C, M, N, K, K, K, N, bias=bias, epilogue=L.EPI_BIAS, remap=remap, col_acc=acc, col_acc_mode=mode, col_acc_grp=grp)
        torch.cuda.synchronize()
        stored = C.double()
        assert not bool((C == 7.0).all(dim=1).any())                   # every output row was written
        if mode == 1:
            want = 3.0 + stored.sum(0)
            assert float((acc.double() - want).abs().max()) <= 1e-5 * float(stored.abs().sum(0).max()) + 1e-3
        else:
            G = grp if grp else N
            got = acc.view(N // G, 2, G)
            s, q = stored.sum(0).view(N // G, G), (stored * stored).sum(0).view(N // G, G)
            assert float((got[:, 0] - s).abs().max()) <= 1e-5 * float(stored.abs().sum(0).max()) + 1e-3
            assert float((got[:, 1] - q).abs().max()) <= 1e-5 * float(q.max()) + 1e-3


def test_fused_column_accumulators_need_the_tensor_core_path(L):
    A = torch.randn(64, 64, device=DEV)
    C = torch.empty(64, 64, device=DEV)
    acc = torch.zeros(64, device=DEV)
    with pytest.raises(L.SstError):
        L.gemm(A, A, C, 64, 64, 64, 64, 64, 64, col_acc=acc, col_acc_mode=1)
