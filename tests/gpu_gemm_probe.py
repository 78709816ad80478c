"""GPU probe (not a pytest file): exercises sst_gemm in every mode against torch fp32 matmul on the same
bf16-rounded operands and prints diagnostics.  Usage: python tests/gpu_gemm_probe.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sst_b200
from sst_b200 import lib as L

torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
L.require_device()
print(L.lib().sst_version().decode())


def report(name, got, ref, tol):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-30
    rel = err.max().item() / denom
    ok = rel < tol and torch.isfinite(got).all().item()
    print("%-58s rel_err %.3e  %s" % (name, rel, "OK" if ok else "FAIL"), flush=True)
    if not ok:
        idx = torch.nonzero(err > tol * denom)
        print("   n_bad", idx.shape[0], "of", err.numel(), "first", idx[:5].tolist(), "last", idx[-3:].tolist())
        rows = torch.unique(idx[:, 0]); cols = torch.unique(idx[:, 1])
        print("   bad rows (first 16)", rows[:16].tolist(), "n", rows.numel(), " bad cols (first 16)", cols[:16].tolist(), "n", cols.numel())
        i, j = idx[0].tolist()
        print("   sample got", got[i, j:j + 8].tolist(), "ref", ref[i, j:j + 8].tolist())
    return ok


def tn_case(M, N, K, simt=False, dtype=torch.bfloat16, epi=0, out_dtype=None):
    A = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
    B = (torch.randn(N, K, device=dev) * 0.5).to(dtype)
    out_dtype = out_dtype or dtype
    Cc = torch.full((M, N), 7.0, device=dev, dtype=out_dtype)
    bias = torch.randn(N, device=dev)
    ref = A.float() @ B.float().t()
    e = 0
    if epi & L.EPI_BIAS:
        ref = ref + bias
    if epi & L.EPI_RELU:
        ref = ref.relu()
    if epi & L.EPI_ACCUM:
        ref = ref + 7.0
    L.gemm(A, B, Cc, M, N, K, K, K, N, bias=bias, epilogue=epi, force_simt=simt)
    torch.cuda.synchronize()
    return report("TN %dx%dx%d %s epi=%d simt=%d out=%s" % (M, N, K, dtype, epi, simt, out_dtype), Cc, ref,
                  2e-2 if out_dtype == torch.bfloat16 else (1e-5 if dtype == torch.float32 else 2e-3))


def mn_case(Kr, M, N, simt=False, accum=False):
    A = (torch.randn(Kr, M, device=dev) * 0.5).bfloat16()
    B = (torch.randn(Kr, N, device=dev) * 0.5).bfloat16()
    Cc = torch.full((M, N), 3.0 if accum else 99.0, device=dev, dtype=torch.float32)
    ref = A.float().t() @ B.float() + (3.0 if accum else 0.0)
    L.gemm(A, B, Cc, M, N, Kr, M, N, N, layout=L.GEMM_NT_MN, epilogue=L.EPI_ACCUM if accum else 0, force_simt=simt)
    torch.cuda.synchronize()
    return report("MN K=%d %dx%d accum=%d simt=%d" % (Kr, M, N, accum, simt), Cc, ref, 2e-3)


def conv_case(n, T, Cin, Cout, stride, simt=False, dtype=torch.bfloat16):
    """k=3 pad=1 conv as segmented GEMM over the time-padded channels-last layout (lead 1 / trail 1)."""
    x = (torch.randn(n, T, Cin, device=dev) * 0.5).to(dtype)
    w = (torch.randn(Cout, Cin, 3, device=dev) * 0.1).to(dtype)
    xp = torch.zeros(n, T + 2, Cin, device=dev, dtype=dtype)
    xp[:, 1:T + 1] = x
    wp = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).contiguous()      # [o][tap*Cin + c]
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), stride=stride, padding=1).transpose(1, 2)
    To = ref.shape[1]
    out = torch.full((n * To, Cout), 5.0, device=dev, dtype=dtype)
    if stride == 1:
        P = T + 2
        L.gemm(xp, wp, out, n * P, Cout, 3 * Cin, Cin, 3 * Cin, Cout, n_seg=3, a_row_shift=(-1, 0, 1), a_col0=(0, 0, 0),
               a_rows=n * P, a_cols=Cin, remap=(P, To, 1), force_simt=simt)
    else:
        P = T // 2 + 1
        L.gemm(xp, wp, out, n * P, Cout, 3 * Cin, 2 * Cin, 3 * Cin, Cout, n_seg=3, a_row_shift=(0, 0, 1), a_col0=(0, Cin, 0),
               a_rows=n * P, a_cols=2 * Cin, remap=(P, To, 0), force_simt=simt)
    torch.cuda.synchronize()
    return report("conv3 n=%d T=%d %d->%d s%d simt=%d %s" % (n, T, Cin, Cout, stride, simt, dtype), out.view(n, To, Cout).reshape(n * To, Cout),
                  ref.reshape(n * To, Cout), 2e-2 if dtype == torch.bfloat16 else 1e-5)


def conv_wgrad_case(n, T, Cin, Cout, simt=False):
    """dW[o, tap*Cin+c] = sum_p dY[p,o] * Xpad[p+tap-1, c]  (stride 1, both in the padded layout)."""
    P = T + 2
    x = torch.zeros(n, P, Cin, device=dev); x[:, 1:T + 1] = torch.randn(n, T, Cin, device=dev) * 0.5
    dy = torch.zeros(n, P, Cout, device=dev); dy[:, 1:T + 1] = torch.randn(n, T, Cout, device=dev) * 0.5
    x = x.bfloat16(); dy = dy.bfloat16()
    xf = x.float().view(n * P, Cin); dyf = dy.float().view(n * P, Cout)
    ref = torch.zeros(Cout, 3 * Cin, device=dev)
    for tap in range(3):
        sh = tap - 1
        xs = torch.zeros_like(xf)
        if sh == -1: xs[1:] = xf[:-1]
        elif sh == 1: xs[:-1] = xf[1:]
        else: xs = xf
        ref[:, tap * Cin:(tap + 1) * Cin] = dyf.t() @ xs
    out = torch.zeros(Cout, 3 * Cin, device=dev)
    L.gemm(dy, x, out, Cout, 3 * Cin, n * P, Cout, Cin, 3 * Cin, layout=L.GEMM_NT_MN, n_seg=3, b_row_shift=(-1, 0, 1),
           b_col0=(0, 0, 0), a_rows=n * P, a_cols=Cout, b_rows=n * P, b_cols=Cin, epilogue=L.EPI_ACCUM, force_simt=simt)
    torch.cuda.synchronize()
    return report("conv3 wgrad n=%d T=%d %d->%d simt=%d" % (n, T, Cin, Cout, simt), out, ref, 3e-3)


results = []
# SIMT first (known-simple kernel), fp32 and bf16
results.append(tn_case(200, 96, 80, simt=True, dtype=torch.float32))
results.append(tn_case(200, 96, 80, simt=True, dtype=torch.float32, epi=L.EPI_BIAS | L.EPI_RELU | L.EPI_ACCUM))
results.append(tn_case(300, 136, 72, simt=True))
results.append(mn_case(500, 128, 256, simt=True))
results.append(conv_case(3, 40, 64, 96, 1, simt=True, dtype=torch.float32))
results.append(conv_case(3, 40, 64, 96, 2, simt=True, dtype=torch.float32))
results.append(conv_wgrad_case(3, 40, 256, 128, simt=True))
# tcgen05
for (M, N, K) in [(128, 128, 64), (128, 256, 64), (128, 128, 256), (256, 512, 768), (800, 768, 768), (1000, 2304, 768),
                  (333, 200, 136), (64000, 768, 768)]:
    results.append(tn_case(M, N, K))
results.append(tn_case(800, 768, 768, epi=L.EPI_BIAS | L.EPI_RELU))
results.append(tn_case(800, 768, 768, epi=L.EPI_BIAS | L.EPI_ACCUM))
results.append(tn_case(800, 3072, 768, epi=L.EPI_BIAS, out_dtype=torch.float32))
for (Kr, M, N) in [(64, 128, 256), (256, 128, 256), (800, 768, 768), (6400, 768, 3072), (1000, 3072, 768), (777, 200, 328)]:
    results.append(mn_case(Kr, M, N))
results.append(mn_case(800, 768, 768, accum=True))
results.append(conv_case(4, 200, 768, 768, 1))
results.append(conv_case(4, 400, 768, 768, 2))
results.append(conv_case(3, 40, 64, 96, 2))
results.append(conv_wgrad_case(4, 200, 768, 768))

# quick timing of the big TN GEMM
M, N, K = 64000, 3072, 768
A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); Cc = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(3): L.gemm(A, B, Cc, M, N, K, K, K, N)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): L.gemm(A, B, Cc, M, N, K, K, K, N)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("TN 64000x3072x768: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9))
for _ in range(3): torch.matmul(A, B.t())
torch.cuda.synchronize(); e0.record()
for _ in range(10): torch.matmul(A, B.t())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("cuBLAS same shape:  %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9))
print("ALL OK" if all(results) else "SOME FAILED: %d/%d ok" % (sum(results), len(results)))
