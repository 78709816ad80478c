"""Drop-in mirror of the reference `speech_recognition/architecture.py` Model API (architecture.py:51-188).

Same constructor, same `forward(length_raw_signal, device, x_raw, y, mode, part, memory)` contract and side effects
(in-place time-shift augmentation consuming one `random.randrange(8)` per training forward, architecture.py:104-108;
the cached `src_key_padding_mask` reused by later `part='decoder'` calls, :165,:176), same state_dict keys/shapes
 and -- because the sub-modules are created in the reference's order with torch's own initialisers
-- the same weights for a given `torch.manual_seed`.  The nn.Modules below are parameter containers only: every
forward/backward FLOP runs in libsst.so through `engine.Engine`; without the library or off sm_100 the forward raises.
"""
import copy
import math
import random

import torch
from torch import nn

from . import lib as L
from .engine import Engine, PAD

try:
    from absl import flags
    FLAGS = flags.FLAGS

    def _define(kind, name, default, doc):
        if name not in FLAGS:
            getattr(flags, "DEFINE_" + kind)(name, default, doc)

    # architecture.py:12-20 and recognition_model.py:38; only defined when the reference modules have not already done so
    _define("integer", "model_size", 768, "number of hidden dimensions")
    _define("integer", "feed_forward_layer_size", 3072, "feed-forward dimensions")
    _define("integer", "num_layers_encoder", 6, "number of encoder layers")
    _define("integer", "num_layers_decoder", 6, "number of decoder layers")
    _define("integer", "n_heads_encoder", 8, "number of heads encoder")
    _define("integer", "n_heads_decoder", 8, "number of heads decoder")
    _define("integer", "relative_distance", 300, "relative positional distance")
    _define("float", "dropout_model", .2, "dropout")
    _define("float", "dropout_pos_emb", .2, "dropout")
    _define("integer", "pad", 42, "Padding value according to the position on phoneme inventory")
    _define("string", "sst_dtype", "fp32", "compute dtype of the B200 path: fp32 (parity) or bf16 (tensor cores)")
    _define("string", "sst_ffn_activation", "relu", "feed-forward activation: relu (the reference as shipped, transformer.py:45,61) or gelu "
            "(exact erf; the variant of its logs_to_save/GELU_* runs)")
    _define("boolean", "sst_packed", True, "run ragged batches packed (sum(lengths) rows, attention by row offsets) instead of padding "
            "every utterance to the longest (data_utils.py:176-185 + pad_sequence, architecture.py:116-117)")
except ImportError:                                               # pragma: no cover
    FLAGS = None


def configure(**kw):
    """Set the model flags programmatically (the reference parses them from sys.argv, recognition_model.py:403):
    configure(model_size=768, num_layers_encoder=6, relative_distance=100, dropout_model=0.2, sst_dtype='bf16', ...)."""
    FLAGS.unparse_flags()
    FLAGS(["sst_b200"] + ["--%s=%s" % (k, v) for k, v in kw.items()], known_only=True)


def _flag(name, default):
    try:
        return getattr(FLAGS, name)
    except Exception:
        return default


class _Container(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("parameter container: the math of this module runs inside libsst.so (sst_b200.engine)")


class ResBlock(_Container):
    """architecture.py:22-35 (parameters only)."""

    def __init__(self, num_ins, num_outs, stride=1):
        super().__init__()
        self.conv1 = nn.Conv1d(num_ins, num_outs, 3, padding=1, stride=stride)
        self.bn1 = nn.BatchNorm1d(num_outs)
        self.conv2 = nn.Conv1d(num_outs, num_outs, 3, padding=1)
        self.bn2 = nn.BatchNorm1d(num_outs)
        self.residual_path = nn.Conv1d(num_ins, num_outs, 1, stride=stride)
        self.res_norm = nn.BatchNorm1d(num_outs)


class LearnedRelativePositionalEmbedding(_Container):
    """transformer.py:233-258 (unmasked, per-head): embeddings (H, 2R-1, dh, 1) ~ N(0, dh^-0.5); never trained (Q2)."""

    def __init__(self, max_relative_pos, num_heads, embedding_dim):
        super().__init__()
        self.embeddings = nn.Parameter(torch.zeros(num_heads, 2 * max_relative_pos - 1, embedding_dim, 1))
        nn.init.normal_(self.embeddings, mean=0.0, std=embedding_dim ** (-0.5))


class MultiHeadAttention(_Container):
    """transformer.py:138-160 (parameters only)."""

    def __init__(self, d_model, n_head, relative_positional, relative_positional_distance):
        super().__init__()
        d_qkv = d_model // n_head
        assert d_qkv * n_head == d_model, 'd_model must be divisible by n_head'
        self.w_q = nn.Parameter(torch.Tensor(n_head, d_model, d_qkv))
        self.w_k = nn.Parameter(torch.Tensor(n_head, d_model, d_qkv))
        self.w_v = nn.Parameter(torch.Tensor(n_head, d_model, d_qkv))
        self.w_o = nn.Parameter(torch.Tensor(n_head, d_qkv, d_model))
        for w in (self.w_q, self.w_k, self.w_v, self.w_o):
            nn.init.xavier_normal_(w)
        if relative_positional:
            self.relative_positional = LearnedRelativePositionalEmbedding(relative_positional_distance, n_head, d_qkv)
        else:
            self.relative_positional = None


class TransformerEncoderLayer(_Container):
    """transformer.py:32-45."""

    def __init__(self, d_model, nhead, dim_feedforward, relative_positional_distance):
        super().__init__()
        self.self_attn = MultiHeadAttention(d_model, nhead, True, relative_positional_distance)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)


class TransformerDecoderLayer(_Container):
    """transformer.py:89-106."""

    def __init__(self, d_model, nhead, dim_feedforward, relative_positional_distance):
        super().__init__()
        self.self_attn = MultiHeadAttention(d_model, nhead, False, relative_positional_distance)
        self.multihead_attn = MultiHeadAttention(d_model, nhead, False, relative_positional_distance)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)


class _LayerStack(_Container):
    """nn.TransformerEncoder/Decoder as used by architecture.py:68-69: N deep copies of one layer under `.layers`."""

    def __init__(self, layer, n):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(n)])


class PositionalEncoding(_Container):
    """transformer.py:408-422: sinusoid buffer `pe` (max_len, 1, d)."""

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe.unsqueeze(0).transpose(0, 1))


class _StepFn(torch.autograd.Function):
    """Connects the engine's explicit forward/backward to torch autograd so that the reference training loop
    (recognition_model.py:90-114: model(...) -> F.ctc_loss / LabelSmoothingLoss -> loss.backward()) works unchanged."""

    @staticmethod
    def forward(ctx, model, x_raw, y, lengths, want_memory, *params):
        eng = model.engine()
        # the encoder memory handed back to a search loop keeps the reference's padded (B, Lmax, D) shape; a training step
        # (nothing but the logits leaves) runs ragged batches packed
        enc_logits, dec_logits, c = eng.forward(x_raw, lengths, y, None, training=model.training, seed=model._next_seed(),
                                                packed=False if want_memory else None)
        ctx.model, ctx.c = model, c
        model._cache_memory(c)
        out_enc = model._unpad_logits(enc_logits, c.B, c.Lmax, eng.n_out_enc)
        if want_memory:
            ctx.mark_non_differentiable(c.x_enc)
            return c.x_enc.view(c.B, c.Lmax, eng.D), out_enc
        out_dec = model._unpad_logits(dec_logits, c.B, c.S, eng.n_out_dec)
        return out_enc, out_dec

    @staticmethod
    def backward(ctx, g0, g1):
        model, c = ctx.model, ctx.c
        eng = model.engine()
        if c.dec_logits is None:
            g_enc, g_dec = g1, None
        else:
            g_enc, g_dec = g0, g1
        names = model._param_names
        G = {n: torch.zeros_like(p) for n, p in eng.P.items()}
        d_enc = model._pad_grad(g_enc, c.B * c.Lmax, eng.n_out_enc)
        d_dec = model._pad_grad(g_dec, c.B * c.S, eng.n_out_dec) if c.dec_logits is not None else None
        eng.backward(c, G, d_enc, d_dec)
        grads = tuple(G[n] if model._trainable[n] else None for n in names)
        return (None, None, None, None, None) + grads


class Model(nn.Module):
    def __init__(self, num_features, num_outs_enc, num_outs_dec, device, dtype=None):
        super().__init__()
        D = _flag("model_size", 768)
        Fd = _flag("feed_forward_layer_size", 3072)
        R = _flag("relative_distance", 300)
        self.cfg = dict(d_model=D, d_ff=Fd, n_enc=_flag("num_layers_encoder", 6), n_dec=_flag("num_layers_decoder", 6),
                        n_heads=_flag("n_heads_encoder", 8), n_heads_dec=_flag("n_heads_decoder", 8), rel_dist=R,
                        dropout=_flag("dropout_model", .2), dropout_pos=_flag("dropout_pos_emb", .2),
                        activation=str(_flag("sst_ffn_activation", "relu")).lower(), packed=bool(_flag("sst_packed", True)))
        self.pad = _flag("pad", PAD)
        self.conv_blocks = nn.Sequential(ResBlock(8, D, 2), ResBlock(D, D, 2), ResBlock(D, D, 2))
        self.w_raw_in = nn.Linear(D, D)
        self.emg_projection = nn.Linear(num_features, D)           # constructed but unused, as in the reference (Q14)
        self.embedding_tgt = nn.Embedding(num_outs_dec, D, padding_idx=self.pad)
        self.pos_decoder = PositionalEncoding(D)
        encoder_layer = TransformerEncoderLayer(D, self.cfg["n_heads"], Fd, R)
        decoder_layer = TransformerDecoderLayer(D, self.cfg["n_heads_dec"], Fd, R)
        self.transformerEncoder = _LayerStack(encoder_layer, self.cfg["n_enc"])
        self.transformerDecoder = _LayerStack(decoder_layer, self.cfg["n_dec"])
        self.w_aux = nn.Linear(D, num_outs_enc)
        self.w_out = nn.Linear(D, num_outs_dec)
        self.device = device
        if dtype is None:
            dtype = torch.bfloat16 if str(_flag("sst_dtype", "fp32")).lower() in ("bf16", "bfloat16") else torch.float32
        self.compute_dtype = dtype
        self.tgt_key_padding_mask = None
        self.src_key_padding_mask = None
        self.memory_key_padding_mask = None
        self.tgt_mask = None
        self._engine = None
        self._weights_version = 0
        self._seed_base = random.getrandbits(48)
        self._seed_ctr = 0
        self._mem_lens = None
        self._mem_shape = None
        self._weights_hooks = []           # called by weights_changed() (sst_b200.train.Trainer: refresh its bf16 shadow)

    # ---------------------------------------------------------------------------------------------- engine plumbing
    def weights_changed(self):
        """Force a rebuild of the packed GEMM operands at the next forward.  Not needed for the usual ways of changing
        weights: in-place updates (torch optimizers, load_state_dict, p.data.copy_) are seen through the tensors' version
        counters, sst_b200.train.Trainer repacks itself; only writes through raw pointers need this call."""
        self._weights_version += 1
        for hook in self._weights_hooks:
            hook()

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.weights_changed()
        return r

    def _next_seed(self):
        self._seed_ctr += 1
        return (self._seed_base + self._seed_ctr * 2654435761) & 0xFFFFFFFFFFFF

    def engine(self):
        if self._engine is None:
            params = dict(self.named_parameters())
            bufs = dict(self.named_buffers())
            for n, p in params.items():
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise L.SstError("sst_b200.Model must live on a CUDA device in fp32 (parameter %s); call .to('cuda')" % n)
            self._param_names = list(params)
            self._param_list = list(params.values())
            self._trainable = {n: not ("relative_positional" in n or n.startswith("emg_projection")) for n in params}
            self._engine = Engine({n: p.data for n, p in params.items()}, bufs, self.cfg, dtype=self.compute_dtype)
        return self._engine

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._engine = None                                      # parameter storage moved: rebuild views lazily
        return r

    def _packed_engine(self):
        """The engine with GEMM operands that match the CURRENT parameter values.  A torch optimizer step, load_state_dict
        or any other in-place write bumps the parameter tensors' version counters; the reference loop calls nothing between
        optim.step() and the next forward (recognition_model.py:115-118, :126), so neither may this path."""
        eng = self.engine()
        eng.pack((self._weights_version, tuple(p._version for p in self._param_list)))
        return eng

    def _unpad_logits(self, logits, B, Lx, C):
        out = torch.empty(B, Lx, C, dtype=torch.float32, device=logits.device)
        L.permute3_cast(logits, out, (1, B * Lx, C), (0, logits.stride(0), 1), (0, C, 1))
        return out

    def _pad_grad(self, g, rows, C):
        eng = self.engine()
        out = torch.zeros(rows, eng.LDH, dtype=eng.dtype, device=g.device)
        g = g.contiguous().float()
        L.permute3_cast(g, out, (1, rows, C), (0, C, 1), (0, eng.LDH, 1))
        return out

    def _cache_memory(self, c):
        self._mem_lens, self._mem_shape = c.lens, (c.B, c.Lmax)
        ar = torch.arange(c.Lmax, device=c.lens.device)
        self.src_key_padding_mask = ar[None, :] >= c.lens[:, None]         # == (x[:,:,0] == 42), architecture.py:85-88,121
        self.memory_key_padding_mask = self.src_key_padding_mask

    # ---------------------------------------------------------------------------------------------- reference API
    def create_tgt_padding_mask(self, tgt):
        return tgt == self.pad

    def create_src_padding_mask(self, src):
        return src == self.pad

    def forward(self, length_raw_signal, device, x_raw=None, y=None, mode='default', part=None, memory=None):
        if mode == "default":
            return self.forward_training(x_raw=x_raw, y=y, length_raw_signal=length_raw_signal, device=device)
        if part == 'encoder':
            return self.forward_search(part=part, length_raw_signal=length_raw_signal, x_raw=x_raw, device=device)
        if part == 'decoder':
            return self.forward_search(length_raw_signal=length_raw_signal, part=part, y=y, memory=memory, device=device)

    def _shift(self, x_raw):
        if self.training:
            r = random.randrange(8)                               # same RNG call as architecture.py:105
            if r > 0:
                L.shift_left(x_raw, x_raw.shape[0], x_raw.shape[1], x_raw.shape[2], r)

    def _check_input(self, x_raw):
        if not (x_raw.is_cuda and x_raw.dtype == torch.float32 and x_raw.is_contiguous()):
            raise L.SstError("x_raw must be a contiguous fp32 CUDA tensor of shape (n, 1600, 8)")

    def forward_training(self, length_raw_signal, device, x_raw=None, y=None):
        self._check_input(x_raw)
        self._shift(x_raw)
        self._packed_engine()
        lengths = [int(v) for v in length_raw_signal]
        self.tgt_key_padding_mask = self.create_tgt_padding_mask(y)
        params = [p for _, p in self.named_parameters()]
        return _StepFn.apply(self, x_raw, y.contiguous(), lengths, False, *params)

    def forward_search(self, part, length_raw_signal, device, x_raw=None, y=None, memory=None):
        eng = self._packed_engine()
        if part == 'encoder':
            self._check_input(x_raw)
            self._shift(x_raw)
            lengths = [int(v) for v in length_raw_signal]
            if torch.is_grad_enabled() and self.training:
                params = [p for _, p in self.named_parameters()]
                return _StepFn.apply(self, x_raw, None, lengths, True, *params)
            x_enc, c = eng.encode(x_raw, lengths, self.training, self._next_seed())
            self._cache_memory(c)
            logits = eng.enc_head(x_enc, c.B * c.Lmax)
            return x_enc.view(c.B, c.Lmax, eng.D), self._unpad_logits(logits, c.B, c.Lmax, eng.n_out_enc)
        if part == 'decoder':
            if self._mem_lens is None:
                raise L.SstError("part='decoder' needs a preceding part='encoder' call (cached src_key_padding_mask)")
            B0, Lm = self._mem_shape
            if memory.dim() != 3 or memory.shape[1] != Lm or memory.shape[2] != eng.D or memory.shape[0] % B0 != 0 \
                    or memory.dtype != eng.dtype:
                raise L.SstError("memory must be the tensor returned by the part='encoder' call (or memory.repeat(k, 1, 1) of it)")
            # BeamSearch.py:111-114 decodes all hypotheses of an utterance as one batch of memory.repeat(n_hyp, 1, 1); the cached
            # src_key_padding_mask (architecture.py:176) broadcasts over them, i.e. the memory lengths repeat the same way
            B = memory.shape[0]
            mem_lens = self._mem_lens if B == B0 else self._mem_lens.repeat(B // B0)
            if y.shape[0] != B:
                raise L.SstError("y has batch %d, memory has batch %d" % (y.shape[0], B))
            mem = memory.contiguous().reshape(B * Lm, eng.D)
            y = y.contiguous()
            self.tgt_key_padding_mask = self.create_tgt_padding_mask(y)
            # search-time prefixes are generated, not padded: a PAD id may sit anywhere, so mask per position (:174,:178-183)
            tgt_pad = self.tgt_key_padding_mask.to(torch.uint8).contiguous()
            seeds = Engine._Seeds(self._next_seed())
            x_dec = eng.decode(y, None, mem, mem_lens, B, Lm, self.training, seeds, tgt_pad=tgt_pad)
            logits = eng.dec_head(x_dec, B * y.shape[1])
            return self._unpad_logits(logits, B, y.shape[1], eng.n_out_dec)
