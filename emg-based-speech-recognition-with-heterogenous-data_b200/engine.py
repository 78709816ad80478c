"""Host-side orchestration of the hot path: explicit forward AND backward over libsst.so kernels.

Mirrors, op for op, the reference modules it replaces (paths relative to /root/reference/speech_recognition/):
  ResBlock x3 + w_raw_in          architecture.py:22-48, :54-59, :109-112
  decollate / pad(42) / masks     architecture.py:116-121, data_utils.py:176-185
  TransformerEncoderLayer         transformer.py:47-64   (post-LN, ReLU FFN, rel-pos MHA :162-210, :260-403)
  TransformerDecoderLayer         transformer.py:108-134 (causal self-attn + cross-attn, no rel-pos)
  embedding_tgt + pos_decoder     architecture.py:126-127, transformer.py:431-435
  w_aux / w_out heads             architecture.py:139
  CTC + label-smoothed CE mix     recognition_model.py:93-107, LabelSmoothingLoss.py:13-15
PyTorch supplies device memory (torch.empty) and the current stream only; there is no autograd in here and no
torch math op on an activation.  Activation layout: token-major rows (b*L + t, D), compute dtype fp32 (parity
mode) or bf16 (tensor-core mode); conv activations are channels-last with one zero halo row on each side of every
1600-sample chunk so that a k=3 convolution is a row-shifted GEMM (include/sst.h).
"""
import math

import torch

from . import lib as L

PAD = 42


class Ctx(dict):
    """Saved-for-backward tensors of one forward pass (plain attribute dict)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class Engine:
    def __init__(self, params, buffers, cfg, dtype=torch.float32, device="cuda"):
        """params: name -> fp32 CUDA tensor (reference names); buffers: BN running stats + pos_decoder.pe."""
        L.require_device()
        self.P = params
        self.Bf = buffers
        self.cfg = dict(cfg)
        self.dtype = dtype
        self.dt = L.F32 if dtype == torch.float32 else L.BF16
        self.dev = torch.device(device)
        self.D = cfg["d_model"]
        self.F = cfg["d_ff"]
        self.H = cfg["n_heads"]
        self.dh = self.D // self.H
        self.Hd = cfg.get("n_heads_dec", self.H)          # FLAGS.n_heads_decoder (architecture.py:17); head dims other than 96
        self.dhd = self.D // self.Hd                      # run on the CUDA-core attention kernels
        self.gelu = cfg.get("activation", "relu") == "gelu"   # exact-erf GELU instead of the shipped ReLU
        self.R = cfg["rel_dist"]
        self.n_enc = cfg["n_enc"]
        self.n_dec = cfg["n_dec"]
        self.n_out_enc = params["w_aux.weight"].shape[0]
        self.n_out_dec = params["w_out.weight"].shape[0]
        self.LDH = 64                      # pitch of the (padded) head-logit matrices
        self.pk = {}
        self.force_simt = False
        self.shadow = {}                   # name -> bf16 view of the parameter kept current by the optimizer (optional)
        self._packed_version = None
        self._bn_scratch = torch.empty(3 * self.D, dtype=torch.float64, device=self.dev)
        self._side = None                  # side stream for work that can run under the decoder (CTC)
        assert self.D % 64 == 0 and self.dh % 8 == 0 and self.dhd % 8 == 0 and self.dhd * self.Hd == self.D

    # ------------------------------------------------------------------------------------------------ utils
    def empty(self, *shape, dtype=None):
        return torch.empty(*shape, dtype=dtype or self.dtype, device=self.dev)

    def zeros(self, *shape, dtype=None):
        return torch.zeros(*shape, dtype=dtype or self.dtype, device=self.dev)

    def gemm(self, A, B, C, M, N, K, lda, ldb, ldc, **kw):
        L.gemm(A, B, C, M, N, K, lda, ldb, ldc, force_simt=self.force_simt, **kw)

    # ------------------------------------------------------------------------------------------------ packing
    def _cast2d(self, w2d, transpose=False):
        N, K = w2d.shape
        if transpose:
            Np = N if N % 8 == 0 else (N + 63) // 64 * 64      # TMA needs 16-byte row pitches (44/43-class heads)
            out = (self.zeros(K, Np) if Np != N else self.empty(K, Np))[:, :N]      # pad columns (if any) stay zero
            L.permute3_cast(w2d, out, (1, K, N), (0, 1, K), (0, Np, 1))
        else:
            if self.dtype == torch.float32:
                return w2d
            out = self.empty(N, K)
            L.permute3_cast(w2d, out, (1, N, K), (0, K, 1), (0, K, 1))
        return out

    def _pack_linear(self, key, w2d, name=None):
        """GEMM operand of a (N_out, K_in) weight: the bf16 shadow view the optimizer maintains (train.FlatState) when there is
        one, else a cast copy.  No transposed copy: input gradients read the same matrix MN-major (SST_GEMM_TN_BMN)."""
        sh = self.shadow.get(name) if name is not None else None
        self.pk[key] = sh.view(w2d.shape) if (sh is not None and self.dtype == torch.bfloat16) else self._cast2d(w2d)
        # input gradients (and the out-projection forward) multiply by the transposed matrix.  SST_GEMM_TN_BMN could read the
        # matrix above MN-major instead, but measured 20-25 % slower than a K-major operand on the big GEMMs (B200, cfg2), so
        # a transposed copy is kept (tiled transpose from the bf16 matrix)
        # (transposed from the shadow / the fp32 master, never from the cast copy above: the repack is replayed as one
        # launch without ordering between its permutes)
        self.pk[key + ".T"] = self._cast2d(self.pk[key] if self.pk[key] is not w2d and sh is not None and self.dtype == torch.bfloat16
                                           else w2d, transpose=True)

    def _pack_heads_in(self, key, ws, W=None, WT=None):
        """ws: list of (H, D, dh) per-head projection weights -> W (len*H*dh, D) and its transpose (D, len*H*dh); W / WT may be
        views (row block / column block) of a wider operand several layers share."""
        H, D, dh = ws[0].shape
        HD = H * dh
        n = len(ws)
        W = self.empty(n * HD, D) if W is None else W
        for s, w in enumerate(ws):
            L.permute3_cast(w, W[s * HD:], (H, dh, D), (D * dh, 1, dh), (dh * D, D, 1))
        WT = self.empty(D, n * HD) if WT is None else WT     # built from the masters too (no read of W: one unordered launch replays this)
        for s, w in enumerate(ws):
            L.permute3_cast(w, WT[:, s * HD:], (H, D, dh), (D * dh, dh, 1), (dh, WT.stride(0), 1))
        self.pk[key] = W
        self.pk[key + ".T"] = WT

    def _pack_conv3(self, key, w, stride):
        Co, Ci, _ = w.shape
        W = self.empty(Co, 3 * Ci)
        L.permute3_cast(w, W, (Co, 3, Ci), (Ci * 3, 1, 3), (3 * Ci, Ci, 1))
        self.pk[key] = W
        flat = w.reshape(-1)
        if stride == 1:
            Wd = self.empty(Ci, 3 * Co)           # segment s = tap k, read with row shift 1-k
            L.permute3_cast(w, Wd, (Ci, 3, Co), (3, 1, Ci * 3), (3 * Co, Co, 1))
            self.pk[key + ".d"] = Wd
        else:
            We = self.empty(Ci, Co)               # even input positions: tap 1
            L.permute3_cast(flat[1:], We, (1, Ci, Co), (0, 3, Ci * 3), (0, Co, 1))
            Wo = self.empty(Ci, 2 * Co)           # odd positions: [tap 2 | tap 0], the latter read from the next row
            L.permute3_cast(flat[2:], Wo, (1, Ci, Co), (0, 3, Ci * 3), (0, 2 * Co, 1))
            L.permute3_cast(flat, Wo[:, Co:], (1, Ci, Co), (0, 3, Ci * 3), (0, 2 * Co, 1))
            self.pk[key + ".de"] = We
            self.pk[key + ".do"] = Wo

    def pack(self, version=None):
        """(Re)build the GEMM-operand forms of every weight.  Call after the weights change."""
        if version is not None and version == self._packed_version:
            return
        # Every repack after the first is ONE launch: the ~230 permutes below are recorded once (sources = parameter
        # storage, destinations = the persistent operand buffers in self.pk) and replayed as a table while those addresses
        # and the set of bf16 shadows stay the same.
        plan = getattr(self, "_pack_plan", None)
        shadow_key = tuple(sorted((k, v.data_ptr()) for k, v in self.shadow.items()))
        if plan is not None and shadow_key == self._pack_shadow_key and plan.valid():
            plan.replay()
            self._packed_version = version
            return
        plan = L.PermutePlan()
        with plan.record():
            self._pack_all()
        self._pack_plan, self._pack_shadow_key = plan, shadow_key
        self._packed_version = version

    def _pack_all(self):
        P, D, C = self.P, self.D, self.D
        self.pk = {}
        # first ResBlock: conv1 (k3) and residual_path (k1) over 8 channels as one (2C, 32) operand
        wcat = self.zeros(2 * C, 32)
        L.permute3_cast(P["conv_blocks.0.conv1.weight"], wcat, (C, 3, 8), (24, 1, 3), (32, 8, 1))
        L.permute3_cast(P["conv_blocks.0.residual_path.weight"], wcat[C:, 24:], (1, C, 8), (0, 8, 1), (0, 32, 1))
        self.pk["first.W"] = wcat
        bcat = torch.empty(2 * C, dtype=torch.float32, device=self.dev)
        L.permute3_cast(P["conv_blocks.0.conv1.bias"], bcat, (1, 1, C), (0, 0, 1), (0, 0, 1))
        L.permute3_cast(P["conv_blocks.0.residual_path.bias"], bcat[C:], (1, 1, C), (0, 0, 1), (0, 0, 1))
        self.pk["first.b"] = bcat
        self._pack_conv3("conv_blocks.0.conv2", P["conv_blocks.0.conv2.weight"], 1)
        for i in (1, 2):
            p = "conv_blocks.%d" % i
            self._pack_conv3(p + ".conv1", P[p + ".conv1.weight"], 2)
            self._pack_conv3(p + ".conv2", P[p + ".conv2.weight"], 1)
            self._pack_linear(p + ".res", P[p + ".residual_path.weight"].view(C, C), p + ".residual_path.weight")
        self._pack_linear("w_raw_in", P["w_raw_in.weight"], "w_raw_in.weight")
        for i in range(self.n_enc):
            p = "transformerEncoder.layers.%d" % i
            a = p + ".self_attn"
            self._pack_heads_in(a + ".qkv", [P[a + ".w_q"], P[a + ".w_k"], P[a + ".w_v"]])
            self._pack_linear(a + ".o", P[a + ".w_o"].view(D, D), a + ".w_o")       # (H*dh, D) = (K, N): read MN-major in forward
            E = P[a + ".relative_positional.embeddings"]
            self.pk[a + ".E"] = self._cast2d(E.view(-1, self.dh))
            self._pack_linear(p + ".linear1", P[p + ".linear1.weight"], p + ".linear1.weight")
            self._pack_linear(p + ".linear2", P[p + ".linear2.weight"], p + ".linear2.weight")
        # cross-attention key / value projections of ALL decoder layers side by side: the encoder memory is projected by one GEMM
        # (N = n_dec * 2D) and its gradient collected by one (K = n_dec * 2D) instead of n_dec short ones that each re-read and
        # re-write the (frames, D) accumulator; a layer's own operand is a row / column block of these
        if self.n_dec > 0:
            kv_all = self.empty(self.n_dec * 2 * D, D)
            kv_all_T = self.empty(D, self.n_dec * 2 * D)
        for i in range(self.n_dec):
            p = "transformerDecoder.layers.%d" % i
            a = p + ".self_attn"
            self._pack_heads_in(a + ".qkv", [P[a + ".w_q"], P[a + ".w_k"], P[a + ".w_v"]])
            self._pack_linear(a + ".o", P[a + ".w_o"].view(D, D), a + ".w_o")
            m = p + ".multihead_attn"
            self._pack_heads_in(m + ".q", [P[m + ".w_q"]])
            self._pack_heads_in(m + ".kv", [P[m + ".w_k"], P[m + ".w_v"]], W=kv_all[i * 2 * D:(i + 1) * 2 * D],
                                WT=kv_all_T[:, i * 2 * D:(i + 1) * 2 * D])
            self._pack_linear(m + ".o", P[m + ".w_o"].view(D, D), m + ".w_o")
            self._pack_linear(p + ".linear1", P[p + ".linear1.weight"], p + ".linear1.weight")
            self._pack_linear(p + ".linear2", P[p + ".linear2.weight"], p + ".linear2.weight")
        if self.n_dec > 0:
            self.pk["dec.kv_all"], self.pk["dec.kv_all.T"] = kv_all, kv_all_T
        self._pack_linear("w_aux", P["w_aux.weight"], "w_aux.weight")
        self._pack_linear("w_out", P["w_out.weight"], "w_out.weight")

    # ------------------------------------------------------------------------------------------------ building blocks
    def _linear_fwd(self, x, M, key, bias=None, relu=False, drop_p=0.0, seed=0, out=None, out_dtype=None, N=None, ldc=None,
                    w_kn=False):
        """y = x W^T for a packed (N, K) weight; w_kn: the packed matrix is (K, N) (the out-projection w_o) and is read MN-major."""
        W = self.pk[key]
        if w_kn:
            K, Nw = W.shape
        else:
            Nw, K = W.shape
        N = N or Nw
        ldc = ldc or N
        if out is None:
            out = self.empty(M, ldc, dtype=out_dtype)
        epi = (L.EPI_BIAS if bias is not None else 0) | (L.EPI_RELU if relu else 0) | (L.EPI_DROPOUT if drop_p > 0 else 0)
        self.gemm(x, W, out, M, N, K, x.stride(0), W.stride(0), ldc, bias=bias, epilogue=epi, drop_p=drop_p, seed=seed,
                  layout=L.GEMM_TN_BMN if w_kn else L.GEMM_TN)
        return out

    def _fused_cols(self, what="ffn"):
        """Which column reductions ride in the tensor-core GEMM epilogue (SstGemmDesc.col_acc) instead of a separate pass.
        Measured on cfg2 (one box, back to back, ms/step): none 51.19 | linear1.bias only 50.49 | BatchNorm statistics only 51.42 |
        both 51.17.  The bias gradient pays (the 64000 x 3072 colsum pass costs 0.52 ms, the epilogue 0.17); the BatchNorm
        statistics do NOT: one epilogue warp only sees 32 rows, so every channel receives rows/32 = 8000 double-precision atomics
        per conv output (+1.9 ms of GEMM time against 0.55 ms of colstats, whose blocks reduce thousands of rows first).  Default
        therefore 'ffn'; SST_FUSED_COLS = 0 | ffn | bn | 1 selects for experiments.  The fp32 parity mode and the CUDA-core
        cross-check always keep the separate passes."""
        import os
        sel = os.environ.get("SST_FUSED_COLS", "ffn")
        return self.dtype == torch.bfloat16 and not self.force_simt and sel in ("1", what)

    def _linear_bwd(self, dy, x, M, key, G, wname, bname=None, dx_out=None, accum_dx=False, aux=None, mask_scale=1.0,
                    need_dx=True, K_dy=None, dx_colsum=None):
        """dW (+)= dy^T x into G[wname] (reference layout (N,K)); db += colsum(dy); dx (+)= dy W (optional relu/dropout mask).
        dx_colsum: fp32 accumulator that receives the column sums of dx from the same GEMM's epilogue (the bias gradient of the
        layer below)."""
        W = self.pk[key]
        N, K = W.shape
        if G is not None:
            self.gemm(dy, x, G[wname], N, K, M, dy.stride(0), x.stride(0), K, layout=L.GEMM_NT_MN, epilogue=L.EPI_ACCUM,
                      a_cols=dy.stride(0) if dy.stride(0) >= N else N, b_cols=K)
            if bname is not None:
                L.colsum_accum(self.dt, dy, M, N, dy.stride(0), G[bname])
        if not need_dx:
            return None
        WT = self.pk[key + ".T"]                # (K, N)
        if dx_out is None:
            dx_out = self.empty(M, K)
        epi = (L.EPI_ACCUM if accum_dx else 0) | (L.EPI_MULMASK if aux is not None else 0)
        self.gemm(dy, WT, dx_out, M, K, N, dy.stride(0), WT.stride(0), dx_out.stride(0), aux=aux, ldaux=aux.stride(0) if aux is not None else 0,
                  epilogue=epi, mask_scale=mask_scale, a_cols=N, col_acc=dx_colsum, col_acc_mode=1 if dx_colsum is not None else 0)
        return dx_out

    def _ffn1_fwd(self, x, M, pfx, p, seed):
        """linear1 + activation + dropout (transformer.py:61).  ReLU rides in the GEMM epilogue; GELU is the separate
        bandwidth-bound kernel and keeps the pre-activation for backward.  Returns (h, saved pre-activation or None)."""
        if not self.gelu:
            return self._linear_fwd(x, M, pfx + ".linear1", bias=self.P[pfx + ".linear1.bias"], relu=True, drop_p=p, seed=seed), None
        pre = self._linear_fwd(x, M, pfx + ".linear1", bias=self.P[pfx + ".linear1.bias"])
        h = self.empty(M, self.F)
        L.gelu_dropout_fwd(self.dt, M, self.F, pre, self.F, p, seed, h, self.F)
        return h, pre

    def _ffn2_bwd(self, dy, h, pre, M, pfx, G, p, seed):
        """gradient w.r.t. linear1's output: (dy W2) masked by the activation derivative and the dropout keep factors."""
        if not self.gelu:
            keep_scale = 1.0 / (1.0 - p) if p > 0 else 1.0
            # the masked input gradient IS linear1's output gradient: its column sums (linear1.bias) leave from the same epilogue
            fused = self._fused_cols() and self.F % 32 == 0
            dh = self._linear_bwd(dy, h, M, pfx + ".linear2", G, pfx + ".linear2.weight", pfx + ".linear2.bias", aux=h, mask_scale=keep_scale,
                                  dx_colsum=G[pfx + ".linear1.bias"] if fused else None)
            return dh, fused
        dh = self._linear_bwd(dy, h, M, pfx + ".linear2", G, pfx + ".linear2.weight", pfx + ".linear2.bias")
        L.gelu_dropout_bwd(self.dt, M, self.F, dh, self.F, pre, self.F, p, seed, dh, self.F)
        return dh, False

    def _ln_fwd(self, x, r, M, prefix, p, seed):
        """returns y; r's buffer is overwritten with the pre-norm sum s (saved for backward)."""
        y = self.empty(M, self.D)
        mean = self.empty(M, dtype=torch.float32)
        rstd = self.empty(M, dtype=torch.float32)
        L.layernorm_fwd(self.dt, M, self.D, x, r, p, seed, self.P[prefix + ".weight"], self.P[prefix + ".bias"], y, r, mean, rstd)
        return y, (r, mean, rstd, p, seed)

    def _ln_bwd(self, dy, saved, M, prefix, G):
        s, mean, rstd, p, seed = saved
        ds = self.empty(M, self.D)
        dr = self.empty(M, self.D) if p > 0 else None
        L.layernorm_bwd(self.dt, M, self.D, dy, s, mean, rstd, self.P[prefix + ".weight"], ds, dr, p, seed,
                        G[prefix + ".weight"], G[prefix + ".bias"])
        return ds, (dr if p > 0 else ds)

    def _attn_desc(self, B, Lq, Lk, ldq, ldk, ldv, causal, mask_q_rows, R, p, seed, q_pad=None, k_pad=None, dec=False,
                   q_off=None, k_off=None, q_rows=0, k_rows=0):
        H, dh = (self.Hd, self.dhd) if dec else (self.H, self.dh)
        return L.attn_desc(self.dt, B, H, Lq, Lk, dh, ldq, ldk, ldv, self.D, causal, mask_q_rows, R,
                           1.0 / math.sqrt(dh), p, seed, self.force_simt, q_pad=q_pad, k_pad=k_pad,
                           q_off=q_off, k_off=k_off, q_rows_total=q_rows, k_rows_total=k_rows)

    # ------------------------------------------------------------------------------------------------ conv front-end
    def _bn_stats(self, x, rows, ld, prefix, training, stats=None):
        """stats: float64 [2*C] batch sums / sums of squares the producing GEMM's epilogue already accumulated (SstGemmDesc.col_acc)."""
        C = self.D
        mean = self.empty(C, dtype=torch.float32)
        invstd = self.empty(C, dtype=torch.float32)
        rm, rv = self.Bf[prefix + ".running_mean"], self.Bf[prefix + ".running_var"]
        if training:
            if stats is None:
                stats = self._bn_scratch
                L.colstats(self.dt, x, rows, C, ld, stats)
            L.bn_finalize(stats, rows, C, 1e-5, 0.1, mean, invstd, rm, rv, True)
            self.Bf[prefix + ".num_batches_tracked"] += 1
        else:
            L.bn_finalize(None, rows, C, 1e-5, 0.1, mean, invstd, rm, rv, False)
        return mean, invstd, self.P[prefix + ".weight"], self.P[prefix + ".bias"]

    def _resblock_fwd(self, i, inp, n, T_in, training, last):
        C = self.D
        T = T_in // 2
        rows = n * T
        pfx = "conv_blocks.%d" % i
        c = Ctx(i=i, n=n, T_in=T_in, T=T, inp=inp)
        # training-mode BatchNorm statistics of the three conv outputs come out of the conv GEMMs' epilogues (col_acc mode 2)
        fuse = training and self._fused_cols("bn")
        st = (lambda k: self.empty(2 * k * C, dtype=torch.float64)) if fuse else (lambda k: None)
        mode = 2 if fuse else 0
        if i == 0:
            col = self.empty(rows, 32)
            L.im2col_first(self.dt, inp, col, n, T_in)
            ycat = self.empty(rows, 2 * C)
            st_cat = st(2)                               # [bn1: sums C | squares C][res_norm: sums C | squares C]
            self.gemm(col, self.pk["first.W"], ycat, rows, 2 * C, 32, 32, 32, 2 * C, bias=self.pk["first.b"], epilogue=L.EPI_BIAS,
                      col_acc=st_cat, col_acc_mode=mode, col_acc_grp=C)
            yc1, ld1, yr, ldr = ycat, 2 * C, ycat[:, C:], 2 * C
            st1, str_ = (st_cat[:2 * C], st_cat[2 * C:]) if fuse else (None, None)
            c.col = col
        else:
            P = T + 1                                    # pair-rows per chunk of the (T_in + 2)-row padded input
            yc1 = self.empty(rows, C)
            st1, str_ = st(1), st(1)
            self.gemm(inp, self.pk[pfx + ".conv1"], yc1, n * P, C, 3 * C, 2 * C, 3 * C, C, n_seg=3, a_row_shift=(0, 0, 1),
                      a_col0=(0, C, 0), a_rows=n * P, a_cols=2 * C, remap=(P, T, 0), bias=self.P[pfx + ".conv1.bias"],
                      epilogue=L.EPI_BIAS, col_acc=st1, col_acc_mode=mode)
            yr = self.empty(rows, C)
            self.gemm(inp, self.pk[pfx + ".res"], yr, n * P, C, C, 2 * C, C, C, a_col0=(C, 0, 0), a_rows=n * P, a_cols=2 * C,
                      remap=(P, T, 0), bias=self.P[pfx + ".residual_path.bias"], epilogue=L.EPI_BIAS, col_acc=str_, col_acc_mode=mode)
            ld1, ldr = C, C
        bn1 = self._bn_stats(yc1, rows, ld1, pfx + ".bn1", training, st1)
        h1p = self.empty(n, T + 2, C)
        L.bn_apply(self.dt, n, T, C, yc1, ld1, bn1, None, 0, None, True, h1p, 1, 1)
        yc2 = self.empty(rows, C)
        st2 = st(1)
        self.gemm(h1p, self.pk[pfx + ".conv2"], yc2, n * (T + 2), C, 3 * C, C, 3 * C, C, n_seg=3, a_row_shift=(-1, 0, 1),
                  a_rows=n * (T + 2), a_cols=C, remap=(T + 2, T, 1), bias=self.P[pfx + ".conv2.bias"], epilogue=L.EPI_BIAS,
                  col_acc=st2, col_acc_mode=mode)
        bn2 = self._bn_stats(yc2, rows, C, pfx + ".bn2", training, st2)
        bnr = self._bn_stats(yr, rows, ldr, pfx + ".res_norm", training, str_)
        lead = 0 if last else 1
        out = self.empty(n, T + 2 * lead, C)
        L.bn_apply(self.dt, n, T, C, yc2, C, bn2, yr, ldr, bnr, True, out, lead, lead)
        c.update(yc1=yc1, ld1=ld1, yr=yr, ldr=ldr, bn1=bn1, h1p=h1p, yc2=yc2, bn2=bn2, bnr=bnr, out=out, lead=lead)
        return out, c

    def _resblock_bwd(self, c, dout, G):
        """dout: compact (n*T, C) gradient of the block output.  Returns the compact gradient of the block input (None for block 0)."""
        C, n, T, i = self.D, c.n, c.T, c.i
        rows = n * T
        pfx = "conv_blocks.%d" % i
        red = self._bn_scratch
        # out = relu(bn2(yc2) + res_norm(yr))
        dyc2p = self.empty(n, T + 2, C)
        if i == 0:
            dycat = self.empty(rows, 2 * C)
            dyr, ld_dyr, lr, tr = dycat[:, C:], 2 * C, 0, 0
        else:
            dyr = self.empty(n, T + 1, C)
            ld_dyr, lr, tr = C, 0, 1
        # y=None: the ReLU mask is recomputed from the BN inputs instead of reading the block output back
        L.bn_bwd(self.dt, n, T, C, dout, C, None, c.lead, c.lead, True,
                 c.yc2, C, c.bn2[0], c.bn2[1], c.bn2[2], dyc2p, C, 1, 1, G[pfx + ".bn2.weight"], G[pfx + ".bn2.bias"],
                 c.yr, c.ldr, c.bnr[0], c.bnr[1], c.bnr[2], dyr, ld_dyr, lr, tr, G[pfx + ".res_norm.weight"], G[pfx + ".res_norm.bias"], red,
                 beta_a=c.bn2[3], beta_b=c.bnr[3])
        # conv2 (k3, s1): weight gradient in packed (C, 3C) form, then unpack-accumulate into (C, C, 3)
        Pp = T + 2
        dW = self.empty(C, 3 * C, dtype=torch.float32)
        self.gemm(dyc2p, c.h1p, dW, C, 3 * C, n * Pp, C, C, 3 * C, layout=L.GEMM_NT_MN, n_seg=3, b_row_shift=(-1, 0, 1),
                  a_rows=n * Pp, a_cols=C, b_rows=n * Pp, b_cols=C)
        L.permute3_cast(dW, G[pfx + ".conv2.weight"], (C, C, 3), (3 * C, 1, C), (3 * C, 3, 1), accumulate=True)
        dh1 = self.empty(rows, C)
        self.gemm(dyc2p, self.pk[pfx + ".conv2.d"], dh1, n * Pp, C, 3 * C, C, 3 * C, C, n_seg=3, a_row_shift=(1, 0, -1),
                  a_rows=n * Pp, a_cols=C, remap=(Pp, T, 1))
        # h1p = relu(bn1(yc1))
        if i == 0:
            dyc1, ld_dyc1, l1, t1 = dycat, 2 * C, 0, 0
        else:
            dyc1 = self.empty(n, T + 1, C)
            ld_dyc1, l1, t1 = C, 0, 1
        L.bn_bwd(self.dt, n, T, C, dh1, C, None, 1, 1, True,
                 c.yc1, c.ld1, c.bn1[0], c.bn1[1], c.bn1[2], dyc1, ld_dyc1, l1, t1, G[pfx + ".bn1.weight"], G[pfx + ".bn1.bias"],
                 None, 0, None, None, None, None, 0, 0, 0, None, None, red, beta_a=c.bn1[3])
        if i == 0:
            dWc = self.empty(2 * C, 32, dtype=torch.float32)
            self.gemm(dycat, c.col, dWc, 2 * C, 32, rows, 2 * C, 32, 32, layout=L.GEMM_NT_MN, a_rows=rows, a_cols=2 * C,
                      b_rows=rows, b_cols=32)
            L.permute3_cast(dWc, G[pfx + ".conv1.weight"], (C, 8, 3), (32, 1, 8), (24, 3, 1), accumulate=True)
            L.permute3_cast(dWc[C:, 24:], G[pfx + ".residual_path.weight"], (1, C, 8), (0, 32, 1), (0, 8, 1), accumulate=True)
            return None
        P = T + 1
        inp = c.inp                                        # (n, T_in + 2, C) == pair view (n*P, 2C)
        dW = self.empty(C, 3 * C, dtype=torch.float32)
        self.gemm(dyc1, inp, dW, C, 3 * C, n * P, C, 2 * C, 3 * C, layout=L.GEMM_NT_MN, n_seg=3, b_row_shift=(0, 0, 1),
                  b_col0=(0, C, 0), a_rows=n * P, a_cols=C, b_rows=n * P, b_cols=2 * C)
        L.permute3_cast(dW, G[pfx + ".conv1.weight"], (C, C, 3), (3 * C, 1, C), (3 * C, 3, 1), accumulate=True)
        self.gemm(dyr, inp, G[pfx + ".residual_path.weight"], C, C, n * P, C, 2 * C, C, layout=L.GEMM_NT_MN, b_col0=(C, 0, 0),
                  a_rows=n * P, a_cols=C, b_rows=n * P, b_cols=2 * C, epilogue=L.EPI_ACCUM)
        dinp = self.empty(n * c.T_in, C)                   # compact; as pair view (n*T, 2C): [even | odd] input positions
        self.gemm(dyc1, self.pk[pfx + ".conv1.de"], dinp, n * P, C, C, C, C, 2 * C, a_rows=n * P, a_cols=C, remap=(P, T, 0))
        self.gemm(dyr, self.pk[pfx + ".res.T"], dinp, n * P, C, C, C, C, 2 * C, a_rows=n * P, a_cols=C, remap=(P, T, 0),
                  epilogue=L.EPI_ACCUM)
        self.gemm(dyc1, self.pk[pfx + ".conv1.do"], dinp.view(-1)[C:], n * P, C, 2 * C, C, 2 * C, 2 * C, n_seg=2,
                  a_row_shift=(0, 1, 0), a_rows=n * P, a_cols=C, remap=(P, T, 0))
        return dinp

    # ------------------------------------------------------------------------------------------------ encoder
    def _enc_layer_fwd(self, x, B, Lx, lens, i, training, seeds, M=None, off=None):
        """M / off: rows and per-utterance first rows of a PACKED batch (SstAttnDesc.q_off); default = the padded (B*Lx) layout."""
        D = self.D
        M = M if M is not None else B * Lx
        p = self.cfg["dropout"] if training else 0.0
        pfx = "transformerEncoder.layers.%d" % i
        a = pfx + ".self_attn"
        qkv = self._linear_fwd(x, M, a + ".qkv")
        o = self.empty(M, D)
        lse = self.empty(2 * B * self.H * Lx, dtype=torch.float32)
        ad = self._attn_desc(B, Lx, Lx, 3 * D, 3 * D, 3 * D, False, True, self.R, p, seeds(), q_off=off, k_off=off, q_rows=M, k_rows=M)
        L.attn_fwd(ad, qkv, qkv[:, D:], qkv[:, 2 * D:], self.pk[a + ".E"], lens, lens, o, lse)
        y = self._linear_fwd(o, M, a + ".o.T")
        x1, ln1 = self._ln_fwd(x, y, M, pfx + ".norm1", p, seeds())
        s_ffn = seeds()
        h, pre = self._ffn1_fwd(x1, M, pfx, p, s_ffn)
        y2 = self._linear_fwd(h, M, pfx + ".linear2", bias=self.P[pfx + ".linear2.bias"])
        x2, ln2 = self._ln_fwd(x1, y2, M, pfx + ".norm2", p, seeds())
        c = Ctx(x=x, qkv=qkv, o=o, lse=lse, ad=ad, ln1=ln1, x1=x1, h=h, pre=pre, s_ffn=s_ffn, ln2=ln2, p=p)
        return x2, c

    def _enc_layer_bwd(self, c, dx2, B, Lx, lens, i, G, M=None):
        D = self.D
        M = M if M is not None else B * Lx
        pfx = "transformerEncoder.layers.%d" % i
        a = pfx + ".self_attn"
        ds2, dy2 = self._ln_bwd(dx2, c.ln2, M, pfx + ".norm2", G)
        dh, b1_done = self._ffn2_bwd(dy2, c.h, c.pre, M, pfx, G, c.p, c.s_ffn)
        # dx1 = ds2 + dh W1   (accumulated in place into ds2)
        self._linear_bwd(dh, c.x1, M, pfx + ".linear1", G, pfx + ".linear1.weight", None if b1_done else pfx + ".linear1.bias",
                         dx_out=ds2, accum_dx=True)
        ds1, dy = self._ln_bwd(ds2, c.ln1, M, pfx + ".norm1", G)
        # out-projection y = o W_o^T-form: packed key ".o" holds (H*dh, D) = w_o itself, ".o.T" holds (D, H*dh)
        # weight gradient directly in the parameter layout: d w_o (H*dh, D) = o^T dy
        self.gemm(c.o, dy, G[a + ".w_o"], D, D, M, D, D, D, layout=L.GEMM_NT_MN, epilogue=L.EPI_ACCUM)
        dO = self.empty(M, D)
        self.gemm(dy, self.pk[a + ".o"], dO, M, D, D, D, D, D)
        dqkv = self.empty(M, 3 * D)
        delta = self.empty(B * self.H * Lx, dtype=torch.float32)
        L.attn_bwd(c.ad, c.qkv, c.qkv[:, D:], c.qkv[:, 2 * D:], self.pk[a + ".E"], lens, lens, c.o, c.lse, dO,
                   dqkv, dqkv[:, D:], dqkv[:, 2 * D:], delta)
        self._qkv_wgrad(dqkv, c.x, M, [a + ".w_q", a + ".w_k", a + ".w_v"], G)
        self.gemm(dqkv, self.pk[a + ".qkv.T"], ds1, M, D, 3 * D, 3 * D, 3 * D, D, epilogue=L.EPI_ACCUM)
        return ds1

    def _qkv_wgrad(self, dproj, x, M, names, G, dec=False):
        """Weight gradients of the fused per-head projections, accumulated straight into the reference's (H, D, dh) layout of
        w_q / w_k / w_v: C^T form  dW^T (D, len*H*dh) = x^T dproj  with segmented output columns (head h of tensor s lands at
        G[names[s]] + h*D*dh, row pitch dh) -- no packed temporary, no permute launches."""
        D = self.D
        H, dh = (self.Hd, self.dhd) if dec else (self.H, self.dh)
        n = len(names)
        base = G[names[0]]
        offs = [0, 0, 0]
        for s, name in enumerate(names):
            g = G[name]
            assert g.dtype == torch.float32 and g.is_contiguous() and (g.data_ptr() - base.data_ptr()) % 4 == 0
            offs[s] = (g.data_ptr() - base.data_ptr()) // 4
        self.gemm(x, dproj, base, D, n * D, M, x.stride(0), dproj.stride(0), dh, layout=L.GEMM_NT_MN, epilogue=L.EPI_ACCUM,
                  a_cols=D, b_cols=n * D, out_seg=(dh, D * dh, H * dh, tuple(offs)))

    # ------------------------------------------------------------------------------------------------ decoder
    def _dec_layer_fwd(self, t, mem, B, S, Lm, tgt_lens, mem_lens, i, training, seeds, tgt_pad=None, Mm=None, mem_off=None, cross_kv=None):
        """Mm / mem_off: rows and per-utterance first rows of a PACKED encoder memory (SstAttnDesc.k_off); cross_kv: this layer's
        already projected memory keys / values (Mm, 2D) (search loops project the memory once, not once per step)."""
        D, M = self.D, B * S
        Mm = Mm if Mm is not None else B * Lm
        p = self.cfg["dropout"] if training else 0.0
        pfx = "transformerDecoder.layers.%d" % i
        a, m = pfx + ".self_attn", pfx + ".multihead_attn"
        qkv = self._linear_fwd(t, M, a + ".qkv")
        o1 = self.empty(M, D)
        lse1 = self.empty(2 * B * self.Hd * S, dtype=torch.float32)
        ad1 = self._attn_desc(B, S, S, 3 * D, 3 * D, 3 * D, True, True, 0, p, seeds(), q_pad=tgt_pad, k_pad=tgt_pad, dec=True)
        L.attn_fwd(ad1, qkv, qkv[:, D:], qkv[:, 2 * D:], None, tgt_lens, tgt_lens, o1, lse1)
        y = self._linear_fwd(o1, M, a + ".o.T")
        t1, ln1 = self._ln_fwd(t, y, M, pfx + ".norm1", p, seeds())
        q = self._linear_fwd(t1, M, m + ".q")
        kv = cross_kv if cross_kv is not None else self._linear_fwd(mem, Mm, m + ".kv")
        o2 = self.empty(M, D)
        lse2 = self.empty(2 * B * self.Hd * S, dtype=torch.float32)
        ad2 = self._attn_desc(B, S, Lm, D, kv.stride(0), kv.stride(0), False, False, 0, p, seeds(), dec=True, k_off=mem_off, k_rows=Mm)
        L.attn_fwd(ad2, q, kv, kv[:, D:], None, None, mem_lens, o2, lse2)
        y2 = self._linear_fwd(o2, M, m + ".o.T")
        t2, ln2 = self._ln_fwd(t1, y2, M, pfx + ".norm2", p, seeds())
        s_ffn = seeds()
        h, pre = self._ffn1_fwd(t2, M, pfx, p, s_ffn)
        y3 = self._linear_fwd(h, M, pfx + ".linear2", bias=self.P[pfx + ".linear2.bias"])
        t3, ln3 = self._ln_fwd(t2, y3, M, pfx + ".norm3", p, seeds())
        c = Ctx(t=t, qkv=qkv, o1=o1, lse1=lse1, ad1=ad1, ln1=ln1, t1=t1, q=q, kv=kv, o2=o2, lse2=lse2, ad2=ad2, ln2=ln2, t2=t2,
                h=h, pre=pre, s_ffn=s_ffn, ln3=ln3, p=p)
        return t3, c

    def _dec_layer_bwd(self, c, dt3, mem, dmem, B, S, Lm, tgt_lens, mem_lens, i, G, Mm=None, dkv=None):
        """dkv: this layer's (Mm, 2D) block of the shared key / value gradient buffer -- the caller then folds all layers into
        the memory gradient with ONE GEMM (Engine.backward); without it the layer adds its own share to dmem."""
        D, M = self.D, B * S
        Mm = Mm if Mm is not None else B * Lm
        pfx = "transformerDecoder.layers.%d" % i
        a, m = pfx + ".self_attn", pfx + ".multihead_attn"
        ds3, dy3 = self._ln_bwd(dt3, c.ln3, M, pfx + ".norm3", G)
        dh, b1_done = self._ffn2_bwd(dy3, c.h, c.pre, M, pfx, G, c.p, c.s_ffn)
        self._linear_bwd(dh, c.t2, M, pfx + ".linear1", G, pfx + ".linear1.weight", None if b1_done else pfx + ".linear1.bias",
                         dx_out=ds3, accum_dx=True)
        ds2, dy2 = self._ln_bwd(ds3, c.ln2, M, pfx + ".norm2", G)
        # cross attention
        self.gemm(c.o2, dy2, G[m + ".w_o"], D, D, M, D, D, D, layout=L.GEMM_NT_MN, epilogue=L.EPI_ACCUM)
        dO2 = self.empty(M, D)
        self.gemm(dy2, self.pk[m + ".o"], dO2, M, D, D, D, D, D)
        dq = self.empty(M, D)
        own = dkv is None
        if own:
            dkv = self.empty(Mm, c.kv.stride(0))[:, :2 * D]          # dk / dv are written with the pitches of k / v
        assert dkv.stride(0) == c.kv.stride(0)
        delta = self.empty(B * self.Hd * S, dtype=torch.float32)
        L.attn_bwd(c.ad2, c.q, c.kv, c.kv[:, D:], None, None, mem_lens, c.o2, c.lse2, dO2, dq, dkv, dkv[:, D:], delta)
        self._qkv_wgrad(dq, c.t1, M, [m + ".w_q"], G, dec=True)
        self.gemm(dq, self.pk[m + ".q.T"], ds2, M, D, D, D, D, D, epilogue=L.EPI_ACCUM)
        self._qkv_wgrad(dkv, mem, Mm, [m + ".w_k", m + ".w_v"], G, dec=True)
        if own:
            WT = self.pk[m + ".kv.T"]
            self.gemm(dkv, WT, dmem, Mm, D, 2 * D, dkv.stride(0), WT.stride(0), D, epilogue=L.EPI_ACCUM)
        ds1, dy1 = self._ln_bwd(ds2, c.ln1, M, pfx + ".norm1", G)
        # self attention
        self.gemm(c.o1, dy1, G[a + ".w_o"], D, D, M, D, D, D, layout=L.GEMM_NT_MN, epilogue=L.EPI_ACCUM)
        dO1 = self.empty(M, D)
        self.gemm(dy1, self.pk[a + ".o"], dO1, M, D, D, D, D, D)
        dqkv = self.empty(M, 3 * D)
        L.attn_bwd(c.ad1, c.qkv, c.qkv[:, D:], c.qkv[:, 2 * D:], None, tgt_lens, tgt_lens, c.o1, c.lse1, dO1,
                   dqkv, dqkv[:, D:], dqkv[:, 2 * D:], delta)
        self._qkv_wgrad(dqkv, c.t, M, [a + ".w_q", a + ".w_k", a + ".w_v"], G, dec=True)
        self.gemm(dqkv, self.pk[a + ".qkv.T"], ds1, M, D, 3 * D, 3 * D, 3 * D, D, epilogue=L.EPI_ACCUM)
        return ds1

    # ------------------------------------------------------------------------------------------------ whole model
    class _Seeds:
        def __init__(self, base):
            self.base, self.n = int(base), 0

        def __call__(self):
            self.n += 1
            return (self.base * 1000003 + self.n * 7919) & 0xFFFFFFFFFFFFFFFF

    def batch_meta(self, lengths):
        """Device copies of the per-utterance frame counts and first rows (int32 lengths, int64 offsets).  encode() makes them itself
        unless they are handed in: a CUDA-graph capture cannot contain the host-to-device copy."""
        offs = [0]
        for l in lengths[:-1]:
            offs.append(offs[-1] + l)
        return (torch.tensor(lengths, dtype=torch.int32).to(self.dev, non_blocking=True),
                torch.tensor(offs, dtype=torch.int64).to(self.dev, non_blocking=True))

    def encode(self, x_raw, lengths, training, seed=0, save=True, packed=False, meta=None):
        """x_raw: (n, 1600, 8) fp32 CUDA (already shifted); lengths: list[int] frames per utterance.
        Returns (x_enc (B*Lmax, D), ctx).  packed=True: a ragged batch is NOT padded to (B, Lmax) -- the
        transformer runs on the sum(lengths) rows decollate_tensor (data_utils.py:176-185) yields, utterance b at rows
        [offs[b], offs[b] + lengths[b]), the attention kernels taking the offsets (SstAttnDesc.q_off / k_off); x_enc is then
        (sum(lengths), D).  Every real position computes what it computes in the padded layout: padded keys have probability
        exactly 0 there, and everything else in the transformer is row-local."""
        assert x_raw.dtype == torch.float32 and x_raw.is_cuda and x_raw.is_contiguous()
        n, T0, Cc = x_raw.shape
        assert Cc == 8 and T0 % 8 == 0
        seeds = Engine._Seeds(seed)
        ctx = Ctx(n=n, blocks=[], layers=[], training=training)
        a, T = x_raw, T0
        for i in range(3):
            with L.nvtx_range("conv%d.fwd" % i):
                a, c = self._resblock_fwd(i, a, n, T, training, last=(i == 2))
            T //= 2
            ctx.blocks.append(c)
        rows3 = n * T
        a3 = a.view(rows3, self.D)
        xlin = self._linear_fwd(a3, rows3, "w_raw_in", bias=self.P["w_raw_in.bias"])
        B, Lmax, total = len(lengths), max(lengths), sum(lengths)
        assert total <= rows3, "lengths exceed the available frames (data_utils.py:182)"
        lens_dev, offs_dev = meta if meta is not None else self.batch_meta(lengths)
        ragged = not (all(l == Lmax for l in lengths) and total == rows3)
        packed = bool(packed and ragged and min(lengths) >= 1)
        M, off = B * Lmax, None
        if ragged:
            ctx.offs = offs_dev
            if packed:
                x, M, off = xlin[:total], total, offs_dev       # the decollated rows as they are
            else:
                x = self.empty(B * Lmax, self.D)
                L.gather_rows_pad(self.dt, xlin, x, offs_dev, lens_dev, B, Lmax, self.D, float(PAD))
        else:
            x = xlin
        ctx.update(a3=a3, rows3=rows3, B=B, Lmax=Lmax, lens=lens_dev, ragged=ragged, lengths=list(lengths), packed=packed, M=M, off=off)
        for i in range(self.n_enc):
            with L.nvtx_range("enc%d.fwd" % i):
                x, c = self._enc_layer_fwd(x, B, Lmax, lens_dev, i, training, seeds, M=M, off=off)
            ctx.layers.append(c)
        ctx.x_enc = x
        ctx.seeds = seeds
        return x, ctx

    def enc_head(self, x_enc, M, ctx=None):
        """w_aux: fp32 logits in a pitch-64 matrix (columns >= 44 undefined), always in the padded (B*Lmax) row layout the
        reference returns and the CTC kernels index: a packed batch (`ctx.packed`) is re-padded here, 64 columns only (padded
        positions read 0; the reference leaves the network's response to the 42-filled rows there, which nothing consumes)."""
        logits = self._linear_fwd(x_enc, M, "w_aux", bias=self.P["w_aux.bias"], out_dtype=torch.float32, ldc=self.LDH)
        if ctx is not None and ctx.packed:
            padded = self.empty(ctx.B * ctx.Lmax, self.LDH, dtype=torch.float32)
            L.gather_rows_pad(L.F32, logits, padded, ctx.off, ctx.lens, ctx.B, ctx.Lmax, self.LDH, 0.0)
            return padded
        return logits

    def ctc_greedy(self, x_raw, lengths):
        """Inference (BASELINE.json config 5): encoder forward in eval mode + CTC best-path decode on the device.
        Returns (ids int32 (B, Lmax) padded with -1, lens int32 (B,)) -- still on the device, no host sync."""
        x_enc, ctx = self.encode(x_raw, lengths, training=False, packed=self.cfg.get("packed", True))
        logits = self.enc_head(x_enc, ctx.M, ctx)
        ids = torch.empty(ctx.B, ctx.Lmax, dtype=torch.int32, device=self.dev)
        lens = torch.empty(ctx.B, dtype=torch.int32, device=self.dev)
        L.ctc_greedy(L.F32, ctx.B, ctx.Lmax, self.n_out_enc, self.n_out_enc - 1, logits, self.LDH, ctx.lens, ids, lens)
        return ids, lens

    # ------------------------------------------------------------------------------------------------ incremental decoding
    def greedy_cached(self, mem, mem_lens, B, Lm, max_seq_length, start_tok, eos_tok, check_every=4):
        """Greedy attention-decoder search with key/value caches: the decoder runs on ONE new position
        per step; self-attention keys/values of earlier positions and the cross-attention keys/values of the memory are
        computed once.  Equivalent to re-running the decoder on the whole prefix (greedy_search.py:19-38) as long as no PAD id
        is generated (a PAD position is masked as a key AND as a query row there, which makes earlier rows depend on the
        prefix length) -- the caller checks and falls back.  Arg-max, append and the stop test stay on the device; the host
        looks at the `done` flags every `check_every` steps.  Returns the (B, n) int64 device tensor of prefixes."""
        D, H = self.D, self.Hd
        Tmax = max_seq_length - 1                                   # decoder input positions
        tokens = torch.full((max_seq_length, B), PAD, dtype=torch.int64, device=self.dev)    # position-major: one step's ids are contiguous
        tokens[0] = start_tok
        done = torch.zeros(B, dtype=torch.uint8, device=self.dev)
        n_done = torch.zeros(1, dtype=torch.int32, device=self.dev)
        n_done_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        cross, cache = [], []
        for i in range(self.n_dec):
            m = "transformerDecoder.layers.%d.multihead_attn" % i
            cross.append(self._linear_fwd(mem, B * Lm, m + ".kv"))
            cache.append(self.zeros(B * Tmax, 2 * D))
        klens = torch.arange(1, Tmax + 1, dtype=torch.int32, device=self.dev).view(Tmax, 1).expand(Tmax, B).contiguous()
        # per-step buffers, allocated once: the loop below launches libsst.so kernels only
        t0 = self.empty(B, D)
        o1, o2 = self.empty(B, D), self.empty(B, D)
        lse = self.empty(2 * B * H, dtype=torch.float32)
        n = 1
        for s in range(Tmax):
            L.embed_posenc_fwd(self.dt, tokens[s], self.P["embedding_tgt.weight"], self.Bf["pos_decoder.pe"], t0, B, 1, D, 0.0, 0)
            t = t0
            klen = klens[s]                                       # s + 1 keys for every sample
            for i in range(self.n_dec):
                pfx = "transformerDecoder.layers.%d" % i
                a, m = pfx + ".self_attn", pfx + ".multihead_attn"
                qkv = self._linear_fwd(t, B, a + ".qkv")
                kv_s = cache[i].view(B, Tmax, 2 * D)[:, s]          # rows b*Tmax + s
                L.permute3_cast(qkv[:, D:], kv_s, (1, B, 2 * D), (0, 3 * D, 1), (0, Tmax * 2 * D, 1))
                ad1 = self._attn_desc(B, 1, Tmax, 3 * D, 2 * D, 2 * D, False, False, 0, 0.0, 0, dec=True)
                L.attn_fwd(ad1, qkv, cache[i], cache[i][:, D:], None, None, klen, o1, lse)
                y1 = self._linear_fwd(o1, B, a + ".o.T")
                t1, _ = self._ln_fwd(t, y1, B, pfx + ".norm1", 0.0, 0)
                q = self._linear_fwd(t1, B, m + ".q")
                ad2 = self._attn_desc(B, 1, Lm, D, 2 * D, 2 * D, False, False, 0, 0.0, 0, dec=True)
                L.attn_fwd(ad2, q, cross[i], cross[i][:, D:], None, None, mem_lens, o2, lse)
                y2 = self._linear_fwd(o2, B, m + ".o.T")
                t2, _ = self._ln_fwd(t1, y2, B, pfx + ".norm2", 0.0, 0)
                hid, _ = self._ffn1_fwd(t2, B, pfx, 0.0, 0)
                y3 = self._linear_fwd(hid, B, pfx + ".linear2", bias=self.P[pfx + ".linear2.bias"])
                t, _ = self._ln_fwd(t2, y3, B, pfx + ".norm3", 0.0, 0)
            logits = self.dec_head(t, B)
            # arg-max, append and the stop latch: one libsst.so kernel (the bit-exact-critical index op of the search)
            L.greedy_pick(logits, self.LDH, B, self.n_out_dec, tokens, 1, B, s + 1, eos_tok, done, n_done)
            n = s + 2
            if n >= max_seq_length:
                break
            if (s + 1) % check_every == 0:
                n_done_host.copy_(n_done, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                if int(n_done_host[0]) == B:
                    break
        return tokens[:n].t().contiguous()

    def decode(self, y, tgt_lens, mem, mem_lens, B, Lm, training, seeds, ctx=None, tgt_pad=None, Mm=None, mem_off=None, cross=None):
        """y: (B, S) int64 CUDA; returns x_dec (B*S, D).  Target padding is either a suffix (`tgt_lens`, the training
        batches of pad_sequence) or an arbitrary per-position mask `tgt_pad` (uint8 (B, S); greedy prefixes, where a
        generated PAD id can sit anywhere: architecture.py:174 masks by `tgt == pad`)."""
        S = y.shape[1]
        p_pos = self.cfg["dropout_pos"] if training else 0.0
        t = self.empty(B * S, self.D)
        s_emb = seeds()
        L.embed_posenc_fwd(self.dt, y, self.P["embedding_tgt.weight"], self.Bf["pos_decoder.pe"], t, B, S, self.D, p_pos, s_emb)
        layers = []
        if cross is None and self.n_dec > 0:
            Mm_ = Mm if Mm is not None else B * Lm
            kv_all = self._linear_fwd(mem, Mm_, "dec.kv_all")            # (Mm, n_dec * 2D): every layer's keys / values in one GEMM
            cross = [kv_all[:, i * 2 * self.D:(i + 1) * 2 * self.D] for i in range(self.n_dec)]
        for i in range(self.n_dec):
            with L.nvtx_range("dec%d.fwd" % i):
                t, c = self._dec_layer_fwd(t, mem, B, S, Lm, tgt_lens, mem_lens, i, training, seeds, tgt_pad, Mm=Mm, mem_off=mem_off,
                                           cross_kv=cross[i] if cross is not None else None)
            layers.append(c)
        if ctx is not None:
            ctx.update(y=y, S=S, tgt_lens=tgt_lens, dec_layers=layers, p_pos=p_pos, s_emb=s_emb, x_dec=t)
        return t

    def project_memory(self, mem, Mm):
        """Cross-attention keys / values of every decoder layer for an encoder memory of Mm rows: [(Mm, 2D)] * n_dec."""
        return [self._linear_fwd(mem, Mm, "transformerDecoder.layers.%d.multihead_attn.kv" % i) for i in range(self.n_dec)]

    def dec_head(self, x_dec, M):
        return self._linear_fwd(x_dec, M, "w_out", bias=self.P["w_out.bias"], out_dtype=torch.float32, ldc=self.LDH)

    def forward(self, x_raw, lengths, y=None, tgt_lens=None, training=True, seed=0, ctc=None, packed=None, meta=None):
        """Model.forward_training (architecture.py:101-139).  Returns (enc_logits (B*L, 64) fp32, dec_logits (B*S, 64) fp32 | None, ctx).
        `ctc` = (targets (B, Smax) int64, target lengths int32, coefficient): start the CTC loss + gradient on the side stream as
        soon as the encoder logits exist, so that its serial alpha/beta recursion runs under the decoder forward."""
        x_enc, ctx = self.encode(x_raw, lengths, training, seed, packed=self.cfg.get("packed", True) if packed is None else packed,
                                 meta=meta)
        B, Lmax = ctx.B, ctx.Lmax
        has_dec = y is not None and y.numel() > 0
        cross = None
        if has_dec and self.n_dec > 0:
            # the one big GEMM of the decoder (every layer's cross-attention keys / values) goes BEFORE the CTC kernels are put on
            # the side stream: their 64 blocks hold 176 KB of shared memory each, and a persistent GEMM that shares the machine
            # with them runs on the remaining SMs only -- the small per-layer decoder kernels are the better company
            kv_all = self._linear_fwd(x_enc, ctx.M, "dec.kv_all")
            cross = [kv_all[:, i * 2 * self.D:(i + 1) * 2 * self.D] for i in range(self.n_dec)]
        ctx.enc_logits = self.enc_head(x_enc, ctx.M, ctx)
        ctx.dec_logits = None
        ctx.loss_out = torch.zeros(3, dtype=torch.float32, device=self.dev)
        if ctc is not None:
            self._ctc_async(ctx, *ctc)
        if has_dec:
            if tgt_lens is None:
                tgt_lens = (y != PAD).sum(1).to(torch.int32)
            x_dec = self.decode(y, tgt_lens, x_enc, ctx.lens, B, Lmax, training, ctx.seeds, ctx, Mm=ctx.M, mem_off=ctx.off, cross=cross)
            ctx.dec_logits = self.dec_head(x_dec, B * y.shape[1])
        return ctx.enc_logits, ctx.dec_logits, ctx

    def _ctc(self, ctx, ctc_targets, ctc_tgt_lens, coef):
        B, Lx = ctx.B, ctx.Lmax
        Smax = ctc_targets.shape[1]
        ctx.ctc_ws = (self.empty(B * Lx * self.n_out_enc, dtype=torch.float32),
                      self.empty(2 * B * Lx * (2 * Smax + 1), dtype=torch.float32),      # alpha and beta lattices
                      self.empty(B, dtype=torch.float32))
        ctx.d_enc_logits = self.empty(B * Lx, self.LDH)
        L.ctc_loss(L.F32, self.dt, B, Lx, self.n_out_enc, self.n_out_enc - 1, ctx.enc_logits, self.LDH, ctc_targets, Smax, ctx.lens,
                   ctc_tgt_lens, coef, ctx.ctc_ws[0], ctx.ctc_ws[1], ctx.ctc_ws[2], ctx.d_enc_logits, self.LDH, ctx.loss_out[2:])

    def _ctc_async(self, ctx, ctc_targets, ctc_tgt_lens, coef):
        """CTC on the side stream; every tensor it touches stays referenced by ctx until the main stream has joined."""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        main = torch.cuda.current_stream()
        # allocate on the main stream (its caching-allocator pool), run on the side stream
        B, Lx = ctx.B, ctx.Lmax
        Smax = ctc_targets.shape[1]
        ctx.ctc_ws = (self.empty(B * Lx * self.n_out_enc, dtype=torch.float32),
                      self.empty(2 * B * Lx * (2 * Smax + 1), dtype=torch.float32),
                      self.empty(B, dtype=torch.float32))
        ctx.d_enc_logits = self.empty(B * Lx, self.LDH)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            L.ctc_loss(L.F32, self.dt, B, Lx, self.n_out_enc, self.n_out_enc - 1, ctx.enc_logits, self.LDH, ctc_targets, Smax, ctx.lens,
                       ctc_tgt_lens, coef, ctx.ctc_ws[0], ctx.ctc_ws[1], ctx.ctc_ws[2], ctx.d_enc_logits, self.LDH, ctx.loss_out[2:])
            ctx.ctc_done = torch.cuda.Event()
            ctx.ctc_done.record(self._side)

    def losses(self, ctx, ctc_targets, ctc_tgt_lens, dec_target, n_valid, alpha, eps_ls=0.1, want_grad=True):
        """recognition_model.py:93-107 on the saved logits.  Writes d(loss)/d(logits) (compute dtype, pitch 64) into ctx and
        returns a float32[3] device tensor (loss, loss_dec, loss_enc)."""
        B = ctx.B
        out = ctx.loss_out
        has_dec = ctx.dec_logits is not None
        if ctx.get("ctc_done") is not None:
            torch.cuda.current_stream().wait_event(ctx.ctc_done)        # started in forward() on the side stream
        else:
            self._ctc(ctx, ctc_targets, ctc_tgt_lens, alpha if has_dec else 1.0)
        if has_dec:
            S = ctx.S
            rows = B * S
            ws = self.empty(2 * rows, dtype=torch.float32)
            ctx.d_dec_logits = self.empty(rows, self.LDH)
            L.ce_sumexp_loss(L.F32, self.dt, rows, S, self.n_out_dec, ctx.dec_logits, self.LDH, dec_target, PAD, eps_ls, n_valid,
                             1.0 - alpha, ws, ctx.d_dec_logits, self.LDH, out[1:])
        ctx.loss_mix = (alpha, has_dec)
        return out

    def backward(self, ctx, G, d_enc_logits=None, d_dec_logits=None, on_stage=None):
        """Accumulates every parameter gradient into G[name] (fp32, reference layout).  `on_stage(label)` is invoked when all
        gradients of a stage are final ("heads", "dec<i>", "embed", "enc<i>", "w_raw_in", "conv<i>") so that the caller can start reducing them."""
        on_stage = on_stage or (lambda label: None)
        B, Lx, D = ctx.B, ctx.Lmax, self.D
        M = ctx.M                                       # B * Lx, or sum(lengths) for a packed batch
        d_enc_logits = d_enc_logits if d_enc_logits is not None else ctx.d_enc_logits
        d_dec_logits = d_dec_logits if d_dec_logits is not None else ctx.get("d_dec_logits")
        if ctx.packed:                                  # the loss kernels index (b, t): bring their gradient to the packed rows
            d_packed = self.empty(M, self.LDH)
            L.scatter_rows(self.dt, d_enc_logits, d_packed, ctx.off, ctx.lens, B, Lx, self.LDH)
            d_enc_logits = d_packed
        # CTC head
        dx = self._linear_bwd(d_enc_logits, ctx.x_enc, M, "w_aux", G, "w_aux.weight", "w_aux.bias")
        if ctx.dec_logits is not None and d_dec_logits is not None:
            S = ctx.S
            Md = B * S
            dt_ = self._linear_bwd(d_dec_logits, ctx.x_dec, Md, "w_out", G, "w_out.weight", "w_out.bias")
            on_stage("heads")
            shared = self.n_dec > 0 and ctx.dec_layers[0].kv.stride(0) == self.n_dec * 2 * D
            dkv_all = self.empty(M, self.n_dec * 2 * D) if shared else None
            for i in reversed(range(self.n_dec)):
                with L.nvtx_range("dec%d.bwd" % i):
                    dt_ = self._dec_layer_bwd(ctx.dec_layers[i], dt_, ctx.x_enc, dx, B, S, Lx, ctx.tgt_lens, ctx.lens, i, G, Mm=M,
                                              dkv=dkv_all[:, i * 2 * D:(i + 1) * 2 * D] if shared else None)
                on_stage("dec%d" % i)
            if shared:       # gradient of the encoder memory through every layer's key / value projection: one K = n_dec * 2D GEMM
                WT = self.pk["dec.kv_all.T"]
                self.gemm(dkv_all, WT, dx, M, D, self.n_dec * 2 * D, dkv_all.stride(0), WT.stride(0), D, epilogue=L.EPI_ACCUM)
            L.embed_bwd(self.dt, ctx.y, dt_, G["embedding_tgt.weight"], B, S, D, PAD, ctx.p_pos, ctx.s_emb)
        on_stage("embed")
        for i in reversed(range(self.n_enc)):
            with L.nvtx_range("enc%d.bwd" % i):
                dx = self._enc_layer_bwd(ctx.layers[i], dx, B, Lx, ctx.lens, i, G, M=M)
            on_stage("enc%d" % i)
        if ctx.packed:
            # dx already is the gradient of w_raw_in's first sum(lengths) output rows; the rest of the last chunk (the 42-filled
            # tail) has gradient 0: run the layer's backward over the real rows only and leave zeros behind them
            da = self.zeros(ctx.rows3, D)
            self._linear_bwd(dx, ctx.a3, M, "w_raw_in", G, "w_raw_in.weight", "w_raw_in.bias", dx_out=da)
        else:
            if ctx.ragged:
                dxlin = self.zeros(ctx.rows3, D)
                L.scatter_rows(self.dt, dx, dxlin, ctx.offs, ctx.lens, B, Lx, D)
            else:
                dxlin = dx
            da = self._linear_bwd(dxlin, ctx.a3, ctx.rows3, "w_raw_in", G, "w_raw_in.weight", "w_raw_in.bias")
        on_stage("w_raw_in")
        for c in reversed(ctx.blocks):
            with L.nvtx_range("conv%d.bwd" % c.i):
                da = self._resblock_bwd(c, da, G)
            on_stage("conv%d" % c.i)
