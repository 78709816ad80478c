"""ctypes binding of libsst.so (include/sst.h).  Fails loudly when the library is missing."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SST_LIB", os.path.join(_HERE, "libsst.so"))      # SST_LIB: an instrumented build (tools/gemm_trace.py)

F32, BF16 = 0, 1
GEMM_TN, GEMM_NT_MN, GEMM_TN_BMN = 0, 1, 2
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
                ("force_simt", C.c_int32),
                ("out_seg_cols", C.c_int64), ("out_seg_stride", C.c_int64), ("out_grp_cols", C.c_int64), ("out_grp_off", C.c_int64 * 3),
                ("col_acc", C.c_void_p), ("col_acc_mode", C.c_int32), ("col_acc_grp", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SstError("libsst.so not built (%s); run __graft_entry__.build() -- there is no fallback path" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.sst_version.restype = C.c_char_p
        _lib.sst_last_error.restype = C.c_char_p
        _lib.sst_launch_count.restype = C.c_longlong
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise SstError("%s failed (%d): %s" % (what, rc, lib().sst_last_error().decode()))


class Profiler:
    """Per-kernel-family device timing with CUDA events on the launching stream (bench.py's roofline leg).
    Every C-ABI wrapper below brackets its launch with `_scope(kind, flops=, bytes=)`; while a Profiler is
    installed (`with Profiler() as prof:`) each scope records a start/stop event pair; `summary()` synchronises and
    returns {kind: dict(ms, launches, flops, bytes)}.  With no profiler installed a scope is a no-op."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _profiler
        self._prev, _profiler = _profiler, self
        return self

    def __exit__(self, *exc):
        global _profiler
        _profiler = self._prev

    def summary(self, by_tag=False):
        torch.cuda.synchronize()
        out = {}
        for kind, flops, nbytes, e0, e1, tag in self.records:
            if by_tag:
                kind = kind + " " + tag
            d = out.setdefault(kind, dict(ms=0.0, launches=0, flops=0.0, bytes=0.0))
            d["ms"] += e0.elapsed_time(e1)
            d["launches"] += 1
            d["flops"] += flops
            d["bytes"] += nbytes
        return out


_profiler = None
_NVTX = os.environ.get("SST_NVTX", "0") not in ("", "0")      # SST_NVTX=1: an NVTX range around every C-ABI launch (kernel family
                                                               # name) and every engine stage (nvtx_range), for nsys / ncu --nvtx


class nvtx_range:
    """`with nvtx_range("enc3.fwd"):` -- a no-op unless SST_NVTX=1 (the tracing the reference lacks)."""
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()


class _scope:
    __slots__ = ("kind", "flops", "bytes", "e0", "tag")

    def __init__(self, kind, flops=0.0, bytes=0.0, tag=""):
        self.kind, self.flops, self.bytes, self.tag = kind, flops, bytes, tag

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.kind)
        if _profiler is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        if _profiler is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _profiler.records.append((self.kind, self.flops, self.bytes, self.e0, e1, self.tag))


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
         remap=(0, 0, 0), force_simt=False, out_seg=None, col_acc=None, col_acc_mode=0, col_acc_grp=0):
    """col_acc / col_acc_mode / col_acc_grp: column sums (1, float32 accumulator) or BatchNorm statistics (2, float64 [2*N], zeroed
    by the call) of the stored bf16 result, fused into the tensor-core epilogue (include/sst.h)."""
    d = GemmDesc()
    d.dtype, d.out_dtype, d.layout = dt(A), dt(Cout), layout
    d.aux_dtype = dt(aux) if aux is not None else F32
    d.M, d.N, d.K = M, N, K
    d.lda, d.ldb, d.ldc, d.ldaux = lda, ldb, ldc, ldaux
    if layout == GEMM_TN:
        d.a_rows = a_rows if a_rows is not None else M
        d.a_cols = a_cols if a_cols is not None else K
        d.b_rows, d.b_cols = N, K
    elif layout == GEMM_TN_BMN:
        d.a_rows = a_rows if a_rows is not None else M
        d.a_cols = a_cols if a_cols is not None else K
        d.b_rows = b_rows if b_rows is not None else K
        d.b_cols = b_cols if b_cols is not None else N
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
    if col_acc is not None:
        d.col_acc, d.col_acc_mode, d.col_acc_grp = col_acc.data_ptr(), int(col_acc_mode), int(col_acc_grp)
    if out_seg is not None:                     # (seg_cols, seg_stride, grp_cols, (grp_off0, grp_off1, grp_off2)), element units
        d.out_seg_cols, d.out_seg_stride, d.out_grp_cols = out_seg[0], out_seg[1], out_seg[2]
        for i, o in enumerate(out_seg[3]):
            d.out_grp_off[i] = o
    m_real = M * remap[1] / remap[0] if remap[0] > 0 else M        # halo rows of a remapped conv GEMM are not work
    tc = d.dtype == BF16 and not force_simt
    kind = ("gemm_tcgen05_wgrad" if layout == GEMM_NT_MN else "gemm_tcgen05") if tc else "gemm_cuda_core"
    tag = "%dx%dx%d seg%d epi%d %s" % (M, N, K, n_seg, epilogue, "f32out" if Cout.dtype == torch.float32 else "") if _profiler is not None else ""
    with _scope(kind, 2.0 * m_real * N * K, (m_real * K + N * K) * A.element_size() + m_real * N * Cout.element_size(), tag):
        check(lib().sst_gemm(C.byref(d), ptr(A), ptr(B), ptr(Cout), ptr(bias), ptr(aux), stream()), "sst_gemm")


class AttnDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32),
                ("dh", C.c_int32), ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64), ("ldo", C.c_int64),
                ("causal", C.c_int32), ("mask_q_rows", C.c_int32), ("rel_dist", C.c_int32), ("scale", C.c_float),
                ("drop_p", C.c_float), ("seed", C.c_uint64), ("force_simt", C.c_int32),
                ("q_pad", C.c_void_p), ("k_pad", C.c_void_p),
                ("q_off", C.c_void_p), ("k_off", C.c_void_p), ("q_rows_total", C.c_int64), ("k_rows_total", C.c_int64)]


def _i64(v):
    return C.c_int64(int(v))


def _f(v):
    return C.c_float(float(v))


def _u64(v):
    return C.c_uint64(int(v) & 0xFFFFFFFFFFFFFFFF)


def attn_desc(dtype, B, H, Lq, Lk, dh, ldq, ldk, ldv, ldo, causal, mask_q_rows, rel_dist, scale, drop_p, seed,
              force_simt=False, q_pad=None, k_pad=None, q_off=None, k_off=None, q_rows_total=0, k_rows_total=0):
    """q_off / k_off: int64 (B,) device tensors of first rows for the packed layouts of include/sst.h (with the number of rows
    the packed matrices hold); None = the padded layout."""
    d = AttnDesc()
    d.dtype, d.B, d.H, d.Lq, d.Lk, d.dh = dtype, B, H, Lq, Lk, dh
    d.ldq, d.ldk, d.ldv, d.ldo = ldq, ldk, ldv, ldo
    d.causal, d.mask_q_rows, d.rel_dist = int(causal), int(mask_q_rows), int(rel_dist)
    d.scale, d.drop_p, d.seed, d.force_simt = scale, drop_p, seed & 0xFFFFFFFFFFFFFFFF, int(force_simt)
    d.q_pad = q_pad.data_ptr() if q_pad is not None else None
    d.k_pad = k_pad.data_ptr() if k_pad is not None else None
    d.q_off = q_off.data_ptr() if q_off is not None else None
    d.k_off = k_off.data_ptr() if k_off is not None else None
    d.q_rows_total, d.k_rows_total = int(q_rows_total), int(k_rows_total)
    assert (q_off is None or (q_off.dtype == torch.int64 and q_rows_total > 0)) and (k_off is None or (k_off.dtype == torch.int64 and k_rows_total > 0))
    d._keep = (q_pad, k_pad, q_off, k_off)        # the descriptor holds raw pointers: keep the tensors alive with it
    return d


def attn_work(d):
    """Band-limited ALGORITHMIC flops of one attention call (SURVEY.md 8(a) flop model, Q3): per (b, h, query) 4*k*dh for
    q.k and p.v plus 2*k*dh for the q.E bias, k = mean number of in-band keys; backward = 2x forward minus the absent dE."""
    Lq, Lk, R = d.Lq, d.Lk, d.rel_dist
    if R > 0 and Lk > R:
        kbar = sum(min(Lk - 1, i + R - 1) - max(0, i - R + 1) + 1 for i in range(Lq)) / float(Lq)
    elif d.causal:
        kbar = (Lk + 1) / 2.0
    else:
        kbar = float(Lk)
    rows = d.B * d.H * Lq
    qk_pv = 4.0 * kbar * d.dh * rows
    bias = 2.0 * kbar * d.dh * rows if R > 0 else 0.0
    return qk_pv + bias, 2.0 * qk_pv + bias


def attn_fwd(d, q, k, v, E, q_lens, k_lens, o, lse):
    with _scope("attn_fwd", attn_work(d)[0] if _profiler is not None else 0.0, tag="Lq%d Lk%d R%d" % (d.Lq, d.Lk, d.rel_dist)):
        check(lib().sst_attn_fwd(C.byref(d), ptr(q), ptr(k), ptr(v), ptr(E), ptr(q_lens), ptr(k_lens), ptr(o), ptr(lse),
                                 stream()), "sst_attn_fwd")


_attn_ws = {}


def attn_bwd_workspace(d, device):
    """Scratch for sst_attn_bwd (hand-off tiles between its two kernels): one buffer per device, grown on demand and shared
    by every attention call of the step (calls on one stream are ordered)."""
    lib().sst_attn_bwd_workspace_bytes.restype = C.c_size_t
    need = int(lib().sst_attn_bwd_workspace_bytes(C.byref(d)))
    if need == 0:
        return None, 0
    buf = _attn_ws.get(device)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _attn_ws[device] = buf
    return buf, need


def attn_bwd(d, q, k, v, E, q_lens, k_lens, o, lse, dO, dq, dk, dv, delta, ws=None):
    if ws is None:
        ws, _ = attn_bwd_workspace(d, q.device)
    nbytes = ws.numel() * ws.element_size() if ws is not None else 0
    with _scope("attn_bwd", attn_work(d)[1] if _profiler is not None else 0.0, tag="Lq%d Lk%d R%d" % (d.Lq, d.Lk, d.rel_dist)):
        check(lib().sst_attn_bwd(C.byref(d), ptr(q), ptr(k), ptr(v), ptr(E), ptr(q_lens), ptr(k_lens), ptr(o), ptr(lse),
                                 ptr(dO), ptr(dq), ptr(dk), ptr(dv), ptr(delta), ptr(ws), C.c_size_t(nbytes), stream()),
              "sst_attn_bwd")


def layernorm_fwd(dtype, rows, D, x, r, drop_p, seed, gamma, beta, y, s_out, mean, rstd, eps=1e-5):
    # algorithmic bytes: read x, r; write y, s (saved pre-norm sum)
    with _scope("layernorm_fwd", bytes=4.0 * rows * D * (2 if dtype == BF16 else 4), tag="rows%d" % rows):
        check(lib().sst_layernorm_fwd(dtype, _i64(rows), D, ptr(x), ptr(r), _f(drop_p), _u64(seed), ptr(gamma), ptr(beta),
                                      ptr(y), ptr(s_out), ptr(mean), ptr(rstd), _f(eps), stream()), "sst_layernorm_fwd")


def layernorm_bwd(dtype, rows, D, dy, s, mean, rstd, gamma, ds, dr, drop_p, seed, dgamma, dbeta):
    # read dy, s; write ds (+ dr when dropout is on)
    with _scope("layernorm_bwd", bytes=(3.0 + (1.0 if drop_p > 0 else 0.0)) * rows * D * (2 if dtype == BF16 else 4), tag="rows%d" % rows):
        check(lib().sst_layernorm_bwd(dtype, _i64(rows), D, ptr(dy), ptr(s), ptr(mean), ptr(rstd), ptr(gamma), ptr(ds),
                                      ptr(dr), _f(drop_p), _u64(seed), ptr(dgamma), ptr(dbeta), stream()), "sst_layernorm_bwd")


def gelu_dropout_fwd(dtype, rows, cols, x, ldx, drop_p, seed, y, ldy):
    with _scope("gelu_fwd", bytes=2.0 * rows * cols * (2 if dtype == BF16 else 4), tag="rows%d" % rows):      # read x, write y
        check(lib().sst_gelu_dropout_fwd(dtype, _i64(rows), cols, ptr(x), _i64(ldx), _f(drop_p), _u64(seed), ptr(y), _i64(ldy), stream()),
              "sst_gelu_dropout_fwd")


def gelu_dropout_bwd(dtype, rows, cols, dy, lddy, x, ldx, drop_p, seed, dx, lddx):
    with _scope("gelu_bwd", bytes=3.0 * rows * cols * (2 if dtype == BF16 else 4), tag="rows%d" % rows):      # read dy, x; write dx
        check(lib().sst_gelu_dropout_bwd(dtype, _i64(rows), cols, ptr(dy), _i64(lddy), ptr(x), _i64(ldx), _f(drop_p), _u64(seed), ptr(dx),
                                         _i64(lddx), stream()), "sst_gelu_dropout_bwd")


def colstats(dtype, x, rows, Cc, ld, stats):
    with _scope("bn_colstats", bytes=1.0 * rows * Cc * (2 if dtype == BF16 else 4), tag="rows%d" % rows):
        check(lib().sst_colstats(dtype, ptr(x), _i64(rows), Cc, _i64(ld), ptr(stats), stream()), "sst_colstats")


def colsum_accum(dtype, x, rows, Cc, ld, out):
    with _scope("bias_colsum", bytes=1.0 * rows * Cc * (2 if dtype == BF16 else 4), tag="rows%d C%d" % (rows, Cc)):
        check(lib().sst_colsum_accum(dtype, ptr(x), _i64(rows), Cc, _i64(ld), ptr(out), stream()), "sst_colsum_accum")


def bn_finalize(stats, count, Cc, eps, momentum, mean, invstd, running_mean, running_var, training):
    check(lib().sst_bn_finalize(ptr(stats), _i64(count), Cc, _f(eps), _f(momentum), ptr(mean), ptr(invstd),
                                ptr(running_mean), ptr(running_var), int(training), stream()), "sst_bn_finalize")


def bn_apply(dtype, n_chunks, T, Cc, xa, lda, sa, xb, ldb, sb, relu, out, lead, trail):
    """sa / sb = (mean, invstd, gamma, beta) tuples; xb / sb may be None."""
    nb = (None, None, None, None) if sb is None else sb
    with _scope("bn_apply", bytes=(2.0 + (1.0 if xb is not None else 0.0)) * n_chunks * T * Cc * (2 if dtype == BF16 else 4),
                tag="T%d %s" % (T, "2br" if xb is not None else "1br")):
        check(lib().sst_bn_apply(dtype, _i64(n_chunks), T, Cc, ptr(xa), _i64(lda), ptr(sa[0]), ptr(sa[1]), ptr(sa[2]), ptr(sa[3]),
                                 ptr(xb), _i64(ldb), ptr(nb[0]), ptr(nb[1]), ptr(nb[2]), ptr(nb[3]), int(relu), ptr(out),
                                 lead, trail, stream()), "sst_bn_apply")


def bn_bwd(dtype, n_chunks, T, Cc, dout, ld_dout, y, y_lead, y_trail, relu,
           xa, lda, mean_a, invstd_a, gamma_a, dxa, ld_dxa, lead_a, trail_a, dgamma_a, dbeta_a,
           xb, ldb, mean_b, invstd_b, gamma_b, dxb, ld_dxb, lead_b, trail_b, dgamma_b, dbeta_b, red, beta_a=None, beta_b=None):
    # two passes (reduce, apply): dout, xa (+xb) (+y when the ReLU mask is not recomputed) read twice; dxa (+dxb) written
    n_in = 2.0 + (1.0 if xb is not None else 0.0) + (1.0 if (relu and y is not None) else 0.0)
    with _scope("bn_bwd", bytes=(2.0 * n_in + 1.0 + (1.0 if xb is not None else 0.0))
                * n_chunks * T * Cc * (2 if dtype == BF16 else 4), tag="T%d %s" % (T, "2br" if xb is not None else "1br")):
        check(lib().sst_bn_bwd(dtype, _i64(n_chunks), T, Cc, ptr(dout), _i64(ld_dout), ptr(y), y_lead, y_trail, int(relu),
                               ptr(xa), _i64(lda), ptr(mean_a), ptr(invstd_a), ptr(gamma_a), ptr(beta_a), ptr(dxa), _i64(ld_dxa),
                               lead_a, trail_a, ptr(dgamma_a), ptr(dbeta_a),
                               ptr(xb), _i64(ldb), ptr(mean_b), ptr(invstd_b), ptr(gamma_b), ptr(beta_b), ptr(dxb), _i64(ld_dxb),
                               lead_b, trail_b, ptr(dgamma_b), ptr(dbeta_b), ptr(red), stream()), "sst_bn_bwd")


def ctc_loss(logits_dtype, grad_dtype, B, L, Cc, blank, logits, ld, targets, Smax, in_lens, tgt_lens, gcoef, lp_ws, alpha_ws,
             nll, grad, ldg, loss_out):
    with _scope("ctc"):
        check(lib().sst_ctc_loss(logits_dtype, grad_dtype, B, L, Cc, blank, ptr(logits), _i64(ld), ptr(targets), Smax, ptr(in_lens),
                                 ptr(tgt_lens), _f(gcoef), ptr(lp_ws), ptr(alpha_ws), ptr(nll), ptr(grad), _i64(ldg), ptr(loss_out),
                                 stream()), "sst_ctc_loss")


def ce_sumexp_loss(logits_dtype, grad_dtype, rows, S, Cc, logits, ld, target, ignore, eps, n_valid, gcoef, row_ws, grad, ldg,
                   loss_out):
    with _scope("ce_sumexp"):
        check(lib().sst_ce_sumexp_loss(logits_dtype, grad_dtype, _i64(rows), S, Cc, ptr(logits), _i64(ld), ptr(target), ignore,
                                       _f(eps), _i64(n_valid), _f(gcoef), ptr(row_ws), ptr(grad), _i64(ldg), ptr(loss_out),
                                       stream()), "sst_ce_sumexp_loss")


def ctc_greedy(logits_dtype, B, L, Cc, blank, logits, ld, in_lens, out_ids, out_lens):
    check(lib().sst_ctc_greedy(logits_dtype, B, L, Cc, blank, ptr(logits), _i64(ld), ptr(in_lens), ptr(out_ids), ptr(out_lens),
                               stream()), "sst_ctc_greedy")


def greedy_pick(logits, row_stride, B, Cc, tokens, tok_stride_b, tok_stride_pos, pos, eos, done, n_done):
    """tokens int64 (element strides given), done (B,) uint8, n_done (1,) int32 -- all on the device."""
    check(lib().sst_greedy_pick(dt(logits), B, Cc, ptr(logits), _i64(row_stride), ptr(tokens), _i64(tok_stride_b), _i64(tok_stride_pos),
                                int(pos), int(eos), ptr(done), ptr(n_done), stream()), "sst_greedy_pick")


def shift_left(x, n_chunks, T, Cc, r):
    check(lib().sst_shift_left(ptr(x), _i64(n_chunks), T, Cc, r, stream()), "sst_shift_left")


def im2col_first(out_dtype, x, col, n_chunks, Tin):
    check(lib().sst_im2col_first(out_dtype, ptr(x), ptr(col), _i64(n_chunks), Tin, stream()), "sst_im2col_first")


def gather_rows_pad(dtype, inp, out, offs, lens, B, Lmax, D, fill):
    check(lib().sst_gather_rows_pad(dtype, ptr(inp), ptr(out), ptr(offs), ptr(lens), B, Lmax, D, _f(fill), stream()),
          "sst_gather_rows_pad")


def scatter_rows(dtype, dout, din, offs, lens, B, Lmax, D):
    check(lib().sst_scatter_rows(dtype, ptr(dout), ptr(din), ptr(offs), ptr(lens), B, Lmax, D, stream()), "sst_scatter_rows")


def embed_posenc_fwd(out_dtype, y, W, pe, out, B, S, D, drop_p, seed):
    check(lib().sst_embed_posenc_fwd(out_dtype, ptr(y), ptr(W), ptr(pe), ptr(out), B, S, D, _f(drop_p), _u64(seed), stream()),
          "sst_embed_posenc_fwd")


def embed_bwd(dtype, y, dout, dW, B, S, D, pad_idx, drop_p, seed):
    check(lib().sst_embed_bwd(dtype, ptr(y), ptr(dout), ptr(dW), B, S, D, pad_idx, _f(drop_p), _u64(seed), stream()),
          "sst_embed_bwd")


class PermuteItem(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("out", C.c_void_p),
                ("d0", C.c_int64), ("d1", C.c_int64), ("d2", C.c_int64), ("s0", C.c_int64), ("s1", C.c_int64), ("s2", C.c_int64),
                ("o0", C.c_int64), ("o1", C.c_int64), ("o2", C.c_int64),
                ("in_dtype", C.c_int32), ("out_dtype", C.c_int32), ("accumulate", C.c_int32),
                ("mode", C.c_int32), ("first_block", C.c_int32), ("nblocks", C.c_int32)]


_permute_log = None          # list while a PermutePlan is recording


class PermutePlan:
    """Records every permute3_cast issued inside `with plan.record():` (they still run) and replays them as ONE launch
    (sst_permute3_cast_batch) while the tensors they touched keep their addresses."""

    def __init__(self):
        self.calls, self.table, self.n, self.blocks, self.hazard = [], None, 0, 0, None

    def record(self):
        plan = self

        class _Rec:
            def __enter__(self_):
                global _permute_log
                plan.calls = []
                _permute_log = plan.calls

            def __exit__(self_, *exc):
                global _permute_log
                _permute_log = None
                if exc[0] is None:
                    plan._finish()
        return _Rec()

    def _finish(self):
        n = len(self.calls)
        arr = (PermuteItem * max(n, 1))()
        for i, (inp, out, dims, si, so, acc) in enumerate(self.calls):
            it = arr[i]
            it.inp, it.out = inp.data_ptr(), out.data_ptr()
            it.d0, it.d1, it.d2 = dims
            it.s0, it.s1, it.s2 = si
            it.o0, it.o1, it.o2 = so
            it.in_dtype, it.out_dtype, it.accumulate = dt(inp), dt(out), int(acc)
        self.n = n
        if n == 0:
            return
        # one launch has no ordering between its items: a table in which one item reads (or accumulates into) what another
        # writes cannot be replayed -- leave it unplanned (valid() stays False and the caller keeps issuing single launches)
        def region(t, dims, strides):
            st = [abs(x) for d, x in zip(dims, strides) if d > 1]
            dm = [d for d in dims if d > 1]
            ext = 1 + sum((d - 1) * x for d, x in zip(dm, st))
            es = t.element_size()
            pitch = max(st) if st else 1                                       # row pitch of a column-sliced matrix view
            inner = 1 + sum((d - 1) * x for d, x in zip(dm, st) if x != pitch)  # elements a "row" of the view spans
            return t.data_ptr(), t.data_ptr() + ext * es, pitch * es, inner * es

        def disjoint(r1, r2):
            if r1[1] <= r2[0] or r2[1] <= r1[0]:
                return True
            # same row pitch, each view narrower than a row, column windows apart: interleaved slices of one matrix
            if r1[2] == r2[2] and r1[3] <= r1[2] and r2[3] <= r2[2]:
                a0, b0 = r1[0] % r1[2], r2[0] % r2[2]
                return a0 + r1[3] <= b0 or b0 + r2[3] <= a0
            return False
        ins = [region(c[0], c[2], c[3]) for c in self.calls]
        outs = [region(c[1], c[2], c[4]) for c in self.calls]
        for i in range(n):
            for j in range(n):
                if j != i and not (disjoint(ins[i], outs[j]) and (j < i or disjoint(outs[i], outs[j]))):
                    self.hazard = (i, j)
                    return
        nb = C.c_int(0)
        check(lib().sst_permute3_plan(arr, n, C.byref(nb)), "sst_permute3_plan")
        self.blocks = nb.value
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.table = host.to(self.calls[0][0].device)
        self.ptrs = [(c[0], c[0].data_ptr(), c[1], c[1].data_ptr()) for c in self.calls]

    def valid(self):
        return self.table is not None and all(a.data_ptr() == pa and b.data_ptr() == pb for a, pa, b, pb in self.ptrs)

    def replay(self):
        with _scope("weight_repack"):
            check(lib().sst_permute3_cast_batch(ptr(self.table), self.n, self.blocks, stream()), "sst_permute3_cast_batch")


def permute3_cast(inp, out, dims, in_strides, out_strides, accumulate=False):
    if _permute_log is not None:
        _permute_log.append((inp, out, tuple(dims), tuple(in_strides), tuple(out_strides), accumulate))
    check(lib().sst_permute3_cast(dt(inp), dt(out), ptr(inp), ptr(out), _i64(dims[0]), _i64(dims[1]), _i64(dims[2]),
                                  _i64(in_strides[0]), _i64(in_strides[1]), _i64(in_strides[2]),
                                  _i64(out_strides[0]), _i64(out_strides[1]), _i64(out_strides[2]), int(accumulate), stream()),
          "sst_permute3_cast")


def adamw(p, g, m, v, n, lr, beta1, beta2, eps, wd, step, p_bf16=None):
    with _scope("adamw", bytes=(28.0 + (2.0 if p_bf16 is not None else 0.0)) * n):    # read p, g, m, v; write p, m, v (+ bf16 shadow)
        check(lib().sst_adamw(ptr(p), ptr(g), ptr(m), ptr(v), _i64(n), _f(lr), _f(beta1), _f(beta2), _f(eps), _f(wd), _i64(step),
                              ptr(p_bf16), stream()), "sst_adamw")


def adamw_dev(p, g, m, v, n, hyper, beta1, beta2, eps, wd, p_bf16=None):
    """AdamW step whose (lr, 1 - beta1^t, sqrt(1 - beta2^t)) are read from the device float[3] `hyper` (CUDA-graph replays)."""
    with _scope("adamw", bytes=(28.0 + (2.0 if p_bf16 is not None else 0.0)) * n):
        check(lib().sst_adamw_dev(ptr(p), ptr(g), ptr(m), ptr(v), _i64(n), ptr(hyper), _f(beta1), _f(beta2), _f(eps), _f(wd),
                                  ptr(p_bf16), stream()), "sst_adamw_dev")


_salt_tensor = None


def set_dropout_salt(t):
    """Register (None: unregister) a device uint64 (int64 tensor of one element) that every dropout kernel adds to its seed."""
    global _salt_tensor
    if t is not None:
        assert t.is_cuda and t.dtype == torch.int64 and t.numel() == 1
    _salt_tensor = t                              # libsst.so keeps the raw pointer: keep the tensor alive
    check(lib().sst_set_dropout_salt(ptr(t)), "sst_set_dropout_salt")


def write_scalars(dst, fmt, *values):
    """struct.pack(fmt, *values) into the device tensor `dst` as a kernel parameter (captured by value at the call)."""
    import struct
    blob = struct.pack(fmt, *values)
    assert len(blob) % 4 == 0 and len(blob) <= 64 and dst.numel() * dst.element_size() >= len(blob)
    check(lib().sst_write_scalars(ptr(dst), blob, len(blob), stream()), "sst_write_scalars")


def nccl_lib_path():
    """The libnccl.so.2 PyTorch itself uses (nvidia-nccl wheel), so that sst_comm_* and torch.distributed share one NCCL."""
    env = os.environ.get("SST_NCCL_LIB")
    if env:
        return env
    try:
        import nvidia.nccl as _n
        for base in list(getattr(_n, "__path__", [])) + ([os.path.dirname(_n.__file__)] if getattr(_n, "__file__", None) else []):
            p = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.isfile(p):
                return p
    except Exception:
        pass
    return "libnccl.so.2"


def comm_nccl_version():
    return int(lib().sst_comm_nccl_version(nccl_lib_path().encode()))


class Comm:
    """sst_comm_* communicator (include/sst.h): `exchange(obj)` must return rank 0's `obj` on every rank (any host channel)."""

    def __init__(self, rank, world, exchange):
        path = nccl_lib_path().encode()
        uid = C.create_string_buffer(128)
        if rank == 0:
            check(lib().sst_comm_unique_id(uid, path), "sst_comm_unique_id")
        blob = exchange(bytes(uid.raw) if rank == 0 else None)
        self.handle = C.c_void_p()
        check(lib().sst_comm_init(blob, rank, world, path, C.byref(self.handle)), "sst_comm_init")
        self.rank, self.world = rank, world

    def allreduce(self, t, average=True):
        """In place over the ranks, enqueued on the current stream."""
        check(lib().sst_comm_allreduce_bucket(self.handle, ptr(t), _i64(t.numel()), dt(t), int(average), stream()), "sst_comm_allreduce_bucket")

    def destroy(self):
        if self.handle:
            check(lib().sst_comm_destroy(self.handle), "sst_comm_destroy")
            self.handle = C.c_void_p()


def launch_count():
    return int(lib().sst_launch_count())
