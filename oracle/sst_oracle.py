"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference hot path.

Plain PyTorch-CPU / numpy restatement of the Silent Speech Transformer forward pass,
hybrid CTC/attention loss and AdamW step, written functionally over a state_dict that
uses the reference's parameter names.  Every function cites the reference file:line it
follows (paths relative to /root/reference/speech_recognition/).  It exists so that the
`-m gpu` parity tests, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` arm have something to check and time against on a box where
/root/reference is absent.  The product package never imports it: the product path is
CUDA-only and raises when libsst.so is missing.

Parity pinning: the reference ships no tests / golden vectors, so
this restatement is pinned by running the UNMODIFIED reference in the build container
(`oracle/ref_harness.py`) on the same seeded inputs/weights: `tests/test_oracle_vs_reference.py`
(live, build container only) and the committed fixtures `tests/golden/*.npz` produced by
`oracle/make_golden.py` (checked everywhere).
"""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F

PAD = 42          # recognition_model.py:38  (FLAGS.pad; also the *value* used to pad EMG, data_utils.py:170)
SOS = 41          # data_utils.py:19  '<S>'
EOS = 40          # data_utils.py:19  '</S>'
N_PHONES = 43     # len(phoneme_inventory), data_utils.py:19
BLANK = 43        # recognition_model.py:98  blank = n_phones
CHUNK = 1600      # recognition_model.py:77   200*8 raw samples per conv chunk
NEG = -1e8        # transformer.py:181-196, :354-357


def make_cfg(d_model=768, d_ff=3072, n_enc=6, n_dec=6, n_heads=8, rel_dist=100,
             dropout=0.0, dropout_pos=0.0, alpha=0.2, eps_ls=0.1):
    """architecture.py:12-20 + recognition_model.py:50 flags as a plain dict."""
    return dict(d_model=d_model, d_ff=d_ff, n_enc=n_enc, n_dec=n_dec, n_heads=n_heads,
                rel_dist=rel_dist, dropout=dropout, dropout_pos=dropout_pos, alpha=alpha,
                eps_ls=eps_ls)


# ----------------------------------------------------------------------------------------
# synthetic weights / inputs (deterministic, independent of the reference constructor)
# ----------------------------------------------------------------------------------------
def positional_table(d_model, max_len=5000):
    """transformer.py:414-421 (buffer `pos_decoder.pe`, shape (max_len,1,d))."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous()


def state_dict_spec(cfg, num_features=112, n_out_enc=44, n_out_dec=43):
    """Ordered (name, shape, kind) list == reference `Model.state_dict()` keys
    (architecture.py:51-78)."""
    D, Fd, H, R = cfg["d_model"], cfg["d_ff"], cfg["n_heads"], cfg["rel_dist"]
    dh = D // H
    spec = []

    def bn(prefix):
        spec.extend([(prefix + ".weight", (D,), "gamma"), (prefix + ".bias", (D,), "beta"),
                     (prefix + ".running_mean", (D,), "zeros"), (prefix + ".running_var", (D,), "ones"),
                     (prefix + ".num_batches_tracked", (), "count")])

    for i in range(3):
        cin = 8 if i == 0 else D
        p = "conv_blocks.%d" % i
        spec.extend([(p + ".conv1.weight", (D, cin, 3), "fan"), (p + ".conv1.bias", (D,), "fan:%d" % (cin * 3))])
        bn(p + ".bn1")
        spec.extend([(p + ".conv2.weight", (D, D, 3), "fan"), (p + ".conv2.bias", (D,), "fan:%d" % (D * 3))])
        bn(p + ".bn2")
        spec.extend([(p + ".residual_path.weight", (D, cin, 1), "fan"),
                     (p + ".residual_path.bias", (D,), "fan:%d" % cin)])
        bn(p + ".res_norm")
    spec.extend([("w_raw_in.weight", (D, D), "fan"), ("w_raw_in.bias", (D,), "fan:%d" % D),
                 ("emg_projection.weight", (D, num_features), "fan"),
                 ("emg_projection.bias", (D,), "fan:%d" % num_features),
                 ("embedding_tgt.weight", (n_out_dec, D), "embed"),
                 ("pos_decoder.pe", (5000, 1, D), "pe")])

    def mha(prefix, relpos):
        H = cfg["n_heads"] if relpos else cfg.get("n_heads_dec", cfg["n_heads"])      # architecture.py:16-17
        dh = D // H
        for w in ("w_q", "w_k", "w_v"):
            spec.append((prefix + "." + w, (H, D, dh), "xavier"))
        spec.append((prefix + ".w_o", (H, dh, D), "xavier"))
        if relpos:
            spec.append((prefix + ".relative_positional.embeddings", (H, 2 * R - 1, dh, 1), "relpos"))

    def ffn_norms(prefix, n_norm):
        spec.extend([(prefix + ".linear1.weight", (Fd, D), "fan"), (prefix + ".linear1.bias", (Fd,), "fan:%d" % D),
                     (prefix + ".linear2.weight", (D, Fd), "fan"), (prefix + ".linear2.bias", (D,), "fan:%d" % Fd)])
        for j in range(1, n_norm + 1):
            spec.extend([(prefix + ".norm%d.weight" % j, (D,), "gamma"), (prefix + ".norm%d.bias" % j, (D,), "beta")])

    for i in range(cfg["n_enc"]):
        p = "transformerEncoder.layers.%d" % i
        mha(p + ".self_attn", True)
        ffn_norms(p, 2)
    for i in range(cfg["n_dec"]):
        p = "transformerDecoder.layers.%d" % i
        mha(p + ".self_attn", False)
        mha(p + ".multihead_attn", False)
        ffn_norms(p, 3)
    spec.extend([("w_aux.weight", (n_out_enc, D), "fan"), ("w_aux.bias", (n_out_enc,), "fan:%d" % D),
                 ("w_out.weight", (n_out_dec, D), "fan"), ("w_out.bias", (n_out_dec,), "fan:%d" % D)])
    return spec


def synthetic_state_dict(cfg, seed=0):
    """Deterministic weights with the reference's init *distributions* (not its RNG stream):
    kaiming-uniform conv/linear (torch defaults), xavier-normal attention (transformer.py:150-153),
    N(0, dh^-0.5) rel-pos (transformer.py:256-258), N(0,1) embedding with zero pad row
    (architecture.py:63).  Norm affine params are jittered around (1, 0) so that parity tests
    exercise them.  One torch.Generator per tensor name => independent of key order."""
    sd = {}
    for idx, (name, shape, kind) in enumerate(state_dict_spec(cfg)):
        g = torch.Generator().manual_seed(seed * 100003 + idx * 7919 + 17)
        if kind == "fan" or kind.startswith("fan:"):
            fan_in = int(kind.split(":")[1]) if ":" in kind else int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "gamma":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "beta":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "zeros":
            t = torch.zeros(shape)
        elif kind == "ones":
            t = torch.ones(shape)
        elif kind == "count":
            t = torch.zeros((), dtype=torch.int64)
        elif kind == "embed":
            t = torch.randn(shape, generator=g)
            t[PAD] = 0.0
        elif kind == "pe":
            t = positional_table(shape[2], shape[0])
        elif kind == "xavier":
            # nn.init.xavier_normal_ on a 3-D tensor: fan_in = size(1)*rf, fan_out = size(0)*rf, rf = size(2)
            fan_in, fan_out = shape[1] * shape[2], shape[0] * shape[2]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / (fan_in + fan_out))
        elif kind == "relpos":
            t = torch.randn(shape, generator=g) * (shape[2] ** -0.5)
        else:
            raise ValueError(kind)
        sd[name] = t
    return sd


def synthetic_batch(n_utt=4, frames=200, tgt_len=30, seed=1234, ragged=None, tgt_lens=None):
    """SURVEY.md §8(d) inputs in the `collate_raw` output contract (read_emg.py:463-504):
    dict with raw_emg (list of (8*frames_i, 8) f32), lengths (frames), phonemes_int
    (list of int64 incl. <S>/</S>), phonemes_int_lengths."""
    g = torch.Generator().manual_seed(seed)
    lengths = list(ragged) if ragged is not None else [frames] * n_utt
    tls = list(tgt_lens) if tgt_lens is not None else [tgt_len] * len(lengths)
    raw, phon = [], []
    for L, tl in zip(lengths, tls):
        x = torch.randn(8 * L, 8, generator=g) * 5.0
        raw.append(x.clamp_(-50.0, 50.0))
        ids = torch.randint(0, 40, (tl,), generator=g, dtype=torch.int64)
        phon.append(torch.cat([torch.tensor([SOS]), ids, torch.tensor([EOS])]))
    return dict(raw_emg=raw, lengths=lengths, phonemes_int=phon,
                phonemes_int_lengths=[int(p.numel()) for p in phon])


# ----------------------------------------------------------------------------------------
# boundary helpers
# ----------------------------------------------------------------------------------------
def combine_fixed_length(tensor_list, length=CHUNK):
    """data_utils.py:165-174: concat along time, pad the tail with the VALUE 42.0, view (n,length,8)."""
    total = sum(t.size(0) for t in tensor_list)
    tl = list(tensor_list)
    if total % length != 0:
        pad_length = length - (total % length)
        tl.append(torch.full((pad_length, *tl[0].size()[1:]), float(PAD), dtype=tl[0].dtype))
        total += pad_length
    tensor = torch.cat(tl, 0)
    return tensor.view(total // length, length, *tensor.size()[1:])


def decollate_tensor(tensor, lengths):
    """data_utils.py:176-185."""
    b, s, d = tensor.size()
    tensor = tensor.reshape(b * s, d)
    out, idx = [], 0
    for L in lengths:
        assert idx + L <= b * s
        out.append(tensor[idx:idx + L])
        idx += L
    return out


def shift_left_(x_raw, r):
    """architecture.py:104-108 (in place, per 1600-chunk, zero fill)."""
    if r > 0:
        x_raw[:, :-r, :] = x_raw[:, r:, :].clone()
        x_raw[:, -r:, :] = 0
    return x_raw


# ----------------------------------------------------------------------------------------
# conv front-end
# ----------------------------------------------------------------------------------------
def batch_norm(x, sd, prefix, training, stats_out=None, momentum=0.1, eps=1e-5):
    """nn.BatchNorm1d (architecture.py:27,29,33) on (n,C,T): batch stats in training mode
    (biased var for normalisation, unbiased for the running buffer), running stats in eval.
    Uses the same F.batch_norm entry the reference's nn.BatchNorm1d calls; the updated buffers are
    returned through `stats_out` instead of mutating the state_dict."""
    rm = sd[prefix + ".running_mean"].detach().clone()
    rv = sd[prefix + ".running_var"].detach().clone()
    y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training, momentum, eps)
    if training and stats_out is not None:
        stats_out[prefix + ".running_mean"] = rm
        stats_out[prefix + ".running_var"] = rv
        stats_out[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    return y


RELU_MASKS = None
"""Test hook (tests/test_engine_gpu.py::test_bf16_gradients_with_aligned_relu_masks): when set to a dict
{site: bool tensor}, the ReLU at `site` multiplies by that mask instead of by (x > 0).  Sites and layouts:
'conv_blocks.i.relu1' / '.relu2' (n, C, T); 'transformerEncoder.layers.i.relu' (L, B, F);
'transformerDecoder.layers.i.relu' (S, B, F).  None (the default) = the reference's F.relu."""


def _relu(x, site):
    if RELU_MASKS is None:
        return F.relu(x)
    return x * RELU_MASKS[site].to(x.dtype)


def _ffn_act(x, site, cfg):
    """transformer.py:45,61: ReLU as shipped; cfg['activation'] == 'gelu' is the exact-erf variant north_star names (no reference
    code: the oracle for it is torch's F.gelu)."""
    return F.gelu(x) if cfg.get("activation", "relu") == "gelu" else _relu(x, site)


def res_block(x, sd, prefix, stride, training, stats_out=None):
    """ResBlock.forward, architecture.py:37-48."""
    inp = x
    x = F.conv1d(x, sd[prefix + ".conv1.weight"], sd[prefix + ".conv1.bias"], stride=stride, padding=1)
    x = _relu(batch_norm(x, sd, prefix + ".bn1", training, stats_out), prefix + ".relu1")
    x = F.conv1d(x, sd[prefix + ".conv2.weight"], sd[prefix + ".conv2.bias"], stride=1, padding=1)
    x = batch_norm(x, sd, prefix + ".bn2", training, stats_out)
    res = F.conv1d(inp, sd[prefix + ".residual_path.weight"], sd[prefix + ".residual_path.bias"], stride=stride)
    res = batch_norm(res, sd, prefix + ".res_norm", training, stats_out)
    return _relu(x + res, prefix + ".relu2")


def conv_frontend(x_raw, sd, training, stats_out=None):
    """architecture.py:109-112: (n,1600,8) -> transpose -> 3 ResBlocks (stride 2) -> transpose -> w_raw_in."""
    x = x_raw.transpose(1, 2)
    for i in range(3):
        x = res_block(x, sd, "conv_blocks.%d" % i, 2, training, stats_out)
    x = x.transpose(1, 2)
    return F.linear(x, sd["w_raw_in.weight"], sd["w_raw_in.bias"])


# ----------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------
def relpos_logits_as_written(q_lbd, emb, R, H):
    """LearnedRelativePositionalEmbedding.forward (unmasked, per-head embeddings),
    transformer.py:260-295 -> :297-325 (pad under no_grad, narrow), :327-360 (einsum, -=1e8),
    :362-395 (pad/view skew).  q_lbd: (L, B*H, dh); emb: (H, 2R-1, dh, 1).  Returns (B*H, L, L)."""
    L = q_lbd.shape[0]
    pad_length = max(L - R, 0)
    start_pos = max(R - L, 0)
    with torch.no_grad():
        padded = F.pad(emb, (0, 0, 0, 0, pad_length, pad_length))
    used = padded.narrow(-3, start_pos, 2 * L - 1)[..., 0]            # (H, 2L-1, dh)
    q4 = q_lbd.view(L, -1, H, q_lbd.shape[-1])
    pl = torch.einsum("lbhd,hmd->lbhm", q4, used)
    pl = pl.contiguous().view(L, -1, pl.shape[-1])
    if L > R:
        pl[:, :, :pad_length] -= 1e8
        pl[:, :, -pad_length:] -= 1e8
    x = F.pad(pl, (0, 1))
    x = x.transpose(0, 1)
    bh = x.shape[0]
    x = x.contiguous().view(bh, L * 2 * L)
    x = F.pad(x, (0, L - 1))
    x = x.view(bh, L + 1, 2 * L - 1)
    return x[:, :L, L - 1:]


def relpos_logits_closed_form(q_bhld, emb, R):
    """Closed form of the above: bias[b,h,i,j] = q[b,h,i,:].E[h, j-i+R-1, :] if |j-i|<R else -1e8."""
    B, H, L, dh = q_bhld.shape
    E = emb[..., 0]                                                     # (H, 2R-1, dh)
    i = torch.arange(L)[:, None]
    j = torch.arange(L)[None, :]
    rel = j - i
    inband = rel.abs() < R
    idx = (rel + R - 1).clamp(0, 2 * R - 2)                             # (L, L)
    qe = torch.einsum("bhid,hmd->bhim", q_bhld, E)                      # (B,H,L,2R-1)
    bias = torch.gather(qe, 3, idx[None, None].expand(B, H, L, L))
    return torch.where(inband[None, None], bias, torch.full_like(bias, NEG))


def multi_head_attention(query, key, value, sd, prefix, cfg, training, relpos,
                         tgt_key_padding_mask=None, tgt_mask=None,
                         src_key_padding_mask=None, memory_key_padding_mask=None, as_written=True):
    """MultiHeadAttention.forward, transformer.py:162-210.  Inputs (T, B, D)."""
    H = cfg["n_heads"]
    w_q, w_k, w_v, w_o = (sd[prefix + "." + n] for n in ("w_q", "w_k", "w_v", "w_o"))
    dh = w_q.shape[-1]
    q = torch.einsum("tbf,hfa->bhta", query, w_q)
    k = torch.einsum("tbf,hfa->bhta", key, w_k)
    v = torch.einsum("tbf,hfa->bhta", value, w_v)
    logits = torch.einsum("bhqa,bhka->bhqk", q, k) / (dh ** 0.5)
    if tgt_mask is not None:
        logits = logits.masked_fill(tgt_mask == float("-inf"), NEG)
    if tgt_key_padding_mask is not None:
        logits = logits.masked_fill(tgt_key_padding_mask.unsqueeze(1).unsqueeze(2), NEG)
        logits = logits.masked_fill(tgt_key_padding_mask.unsqueeze(1).unsqueeze(3), NEG)
    if src_key_padding_mask is not None:
        logits = logits.masked_fill(src_key_padding_mask.unsqueeze(1).unsqueeze(2), NEG)
        logits = logits.masked_fill(src_key_padding_mask.unsqueeze(1).unsqueeze(3), NEG)
    if memory_key_padding_mask is not None:
        logits = logits.masked_fill(memory_key_padding_mask.unsqueeze(1).unsqueeze(2), NEG)
    if relpos:
        emb = sd[prefix + ".relative_positional.embeddings"]
        if as_written:
            q_pos = q.permute(2, 0, 1, 3)
            l, b, h, d = q_pos.size()
            pos = relpos_logits_as_written(q_pos.reshape(l, b * h, d), emb, cfg["rel_dist"], H)
            logits = logits + pos.view(b, h, l, l)
        else:
            logits = logits + relpos_logits_closed_form(q, emb.detach(), cfg["rel_dist"])
    probs = F.softmax(logits, dim=-1)
    probs = F.dropout(probs, cfg["dropout"], training)
    o = torch.einsum("bhqk,bhka->bhqa", probs, v)
    return torch.einsum("bhta,haf->tbf", o, w_o)


def layer_norm(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def encoder_layer(src, sd, prefix, cfg, training, kpm, as_written=True):
    """TransformerEncoderLayer.forward, transformer.py:47-64 (post-LN, ReLU FFN)."""
    p = cfg["dropout"]
    src2 = multi_head_attention(src, src, src, sd, prefix + ".self_attn", cfg, training, True,
                                src_key_padding_mask=kpm, as_written=as_written)
    src = layer_norm(src + F.dropout(src2, p, training), sd, prefix + ".norm1")
    h = F.dropout(_ffn_act(F.linear(src, sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"]), prefix + ".relu", cfg), p, training)
    src2 = F.linear(h, sd[prefix + ".linear2.weight"], sd[prefix + ".linear2.bias"])
    return layer_norm(src + F.dropout(src2, p, training), sd, prefix + ".norm2")


def decoder_layer(tgt, memory, sd, prefix, cfg, training, tgt_mask, tgt_kpm, mem_kpm):
    """TransformerDecoderLayer.forward, transformer.py:108-134 (no rel-pos: :92-93)."""
    p = cfg["dropout"]
    t2 = multi_head_attention(tgt, tgt, tgt, sd, prefix + ".self_attn", cfg, training, False,
                              tgt_key_padding_mask=tgt_kpm, tgt_mask=tgt_mask)
    tgt = layer_norm(tgt + F.dropout(t2, p, training), sd, prefix + ".norm1")
    t2 = multi_head_attention(tgt, memory, memory, sd, prefix + ".multihead_attn", cfg, training, False,
                              memory_key_padding_mask=mem_kpm)
    tgt = layer_norm(tgt + F.dropout(t2, p, training), sd, prefix + ".norm2")
    h = F.dropout(_ffn_act(F.linear(tgt, sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"]), prefix + ".relu", cfg), p, training)
    t2 = F.linear(h, sd[prefix + ".linear2.weight"], sd[prefix + ".linear2.bias"])
    return layer_norm(tgt + F.dropout(t2, p, training), sd, prefix + ".norm3")


# ----------------------------------------------------------------------------------------
# Model.forward restatements
# ----------------------------------------------------------------------------------------
def encode(sd, cfg, x_raw, lengths, training, stats_out=None, as_written=True):
    """Encoder half of forward_training / forward_search(part='encoder'),
    architecture.py:109-121,129-131 / :145-171.  Returns (x_encoder (B,L,D), src_kpm (B,L))."""
    x = conv_frontend(x_raw, sd, training, stats_out)
    xs = decollate_tensor(x, lengths)
    x = torch.nn.utils.rnn.pad_sequence(xs, batch_first=True, padding_value=float(PAD))
    kpm = x[:, :, 0] == PAD
    x = x.transpose(0, 1)
    for i in range(cfg["n_enc"]):
        x = encoder_layer(x, sd, "transformerEncoder.layers.%d" % i, cfg, training, kpm, as_written)
    return x.transpose(0, 1), kpm


def decode(sd, cfg, y, memory_bld, mem_kpm, training):
    """Decoder half, architecture.py:119-139 / :173-188.  y: (B,S) int64; returns x_decoder (B,S,D).
    PositionalEncoding is applied to the batch-first tensor (Q10): adds pe[b]/D to sample b."""
    D = cfg["d_model"]
    tgt_kpm = y == PAD
    S = y.shape[1]
    tgt_mask = torch.triu(torch.full((S, S), float("-inf")), diagonal=1)
    tgt = F.embedding(y, sd["embedding_tgt.weight"], padding_idx=PAD)
    assert tgt.size(0) < 5000
    tgt = tgt + (1.0 / D) * sd["pos_decoder.pe"][:tgt.size(0), :]
    tgt = F.dropout(tgt, cfg["dropout_pos"], training)
    tgt = tgt.transpose(0, 1)
    mem = memory_bld.transpose(0, 1)
    for i in range(cfg["n_dec"]):
        tgt = decoder_layer(tgt, mem, sd, "transformerDecoder.layers.%d" % i, cfg, training,
                            tgt_mask, tgt_kpm, mem_kpm)
    return tgt.transpose(0, 1)


def forward_training(sd, cfg, x_raw, y, lengths, training=True, shift_r=0, stats_out=None, as_written=True):
    """Model.forward_training, architecture.py:101-139.  Returns (w_aux(enc) (B,L,44), w_out(dec) (B,S,43))."""
    if training:
        shift_left_(x_raw, shift_r)
    x_enc, kpm = encode(sd, cfg, x_raw, lengths, training, stats_out, as_written)
    out_enc = F.linear(x_enc, sd["w_aux.weight"], sd["w_aux.bias"])
    if cfg["n_dec"] == 0 or y is None:
        return out_enc, None
    x_dec = decode(sd, cfg, y, x_enc, kpm, training)
    return out_enc, F.linear(x_dec, sd["w_out.weight"], sd["w_out.bias"])


def label_smoothing_loss(logits_bcs, target, eps=0.1):
    """LabelSmoothingLoss.forward, LabelSmoothingLoss.py:13-15 (Q11): input (B,C,S), S = target length."""
    ce = F.cross_entropy(logits_bcs, target, ignore_index=PAD)
    return (1 - eps) * ce + (eps / logits_bcs.shape[2]) * torch.sum(torch.exp(logits_bcs))


def ctc_loss_ref(out_enc_blc, targets, in_lens, tgt_lens):
    """recognition_model.py:93-98: log_softmax over classes, (L,B,C), blank=43, reduction 'mean'."""
    lp = F.log_softmax(out_enc_blc, 2).transpose(1, 0)
    return F.ctc_loss(lp, targets, in_lens, tgt_lens, blank=BLANK)


def make_targets(batch):
    """recognition_model.py:85-87,95-97: decoder input/target and CTC target from phonemes_int."""
    pad_seq = torch.nn.utils.rnn.pad_sequence
    target = pad_seq(batch["phonemes_int"], batch_first=True, padding_value=PAD)
    tgt_in, tgt_out = target[:, :-1], target[:, 1:]
    ctc_lens = [n - 2 for n in batch["phonemes_int_lengths"]]
    ctc_tgt = pad_seq([p[1:-1] for p in batch["phonemes_int"]], batch_first=True, padding_value=PAD)
    return tgt_in, tgt_out, ctc_tgt, ctc_lens


def train_step_losses(sd, cfg, batch, training=True, shift_r=0, stats_out=None, as_written=True):
    """The arithmetic of recognition_model.py:76-107 for one micro-batch.  Returns dict of tensors."""
    X = combine_fixed_length(batch["raw_emg"], CHUNK).clone()
    tgt_in, tgt_out, ctc_tgt, ctc_lens = make_targets(batch)
    out_enc, out_dec = forward_training(sd, cfg, X, tgt_in if cfg["n_dec"] > 0 else None,
                                        batch["lengths"], training, shift_r, stats_out, as_written)
    loss_enc = ctc_loss_ref(out_enc, ctc_tgt, batch["lengths"], ctc_lens)
    res = dict(out_enc=out_enc, loss_enc=loss_enc)
    if out_dec is not None:
        loss_dec = label_smoothing_loss(out_dec.permute(0, 2, 1), tgt_out, cfg["eps_ls"])
        res.update(out_dec=out_dec, loss_dec=loss_dec,
                   loss=(1 - cfg["alpha"]) * loss_dec + cfg["alpha"] * loss_enc)
    else:
        res["loss"] = loss_enc
    return res


def trainable_names(sd, cfg):
    """Names that receive a gradient in the reference: everything floating except BN buffers,
    pos_decoder.pe, the rel-pos embeddings (Q2) and emg_projection (Q14)."""
    out = []
    for name, t in sd.items():
        if not t.is_floating_point():
            continue
        if name.endswith("running_mean") or name.endswith("running_var") or name == "pos_decoder.pe":
            continue
        if "relative_positional" in name or name.startswith("emg_projection"):
            continue
        if cfg["n_dec"] == 0 and (name.startswith("transformerDecoder") or name.startswith("w_out")
                                  or name.startswith("embedding_tgt")):
            continue
        out.append(name)
    return out


def loss_and_grads(sd, cfg, batch, training=True, shift_r=0, as_written=True):
    """forward + backward (recognition_model.py:107,114); returns (results, grads by name, new BN buffers)."""
    sdg = {}
    names = set(trainable_names(sd, cfg))
    for name, t in sd.items():
        sdg[name] = t.clone().requires_grad_(True) if name in names else t
    stats = {}
    res = train_step_losses(sdg, cfg, batch, training, shift_r, stats, as_written)
    res["loss"].backward()
    grads = {n: sdg[n].grad for n in names if sdg[n].grad is not None}
    return {k: v.detach() for k, v in res.items()}, grads, stats


def lr_schedule(iteration, target_lr=3e-4, warmup=1500):
    """recognition_model.py:57-64: linear warm-up; returns None when lr is left unchanged."""
    iteration = iteration + 1
    if iteration <= warmup:
        return iteration * target_lr / warmup
    return None


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.01):
    """torch.optim.AdamW defaults (recognition_model.py:293), single tensor, in place; `step` is 1-based."""
    p.mul_(1 - lr * wd)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# ----------------------------------------------------------------------------------------
# explicit CTC alpha/beta (numpy float64) -- independent of torch's ctc_loss kernel
# ----------------------------------------------------------------------------------------
def ctc_alpha_beta(logits_lc, target, blank=BLANK):
    """One utterance.  logits_lc: (T, C) raw logits (log-softmax applied here, recognition_model.py:93);
    target: 1-D int array (length S >= 0).  Returns (nll, d nll / d logits (T,C)).  Graves 2006 eq. 6-16,
    the recursion F.ctc_loss (recognition_model.py:98) implements."""
    x = np.asarray(logits_lc, dtype=np.float64)
    T, C = x.shape
    lp = x - x.max(1, keepdims=True)
    lp = lp - np.log(np.exp(lp).sum(1, keepdims=True))
    S = len(target)
    ext = np.full(2 * S + 1, blank, dtype=np.int64)
    ext[1::2] = target
    n = 2 * S + 1
    NEGINF = -np.inf

    def lse(*a):
        m = max(a)
        if m == NEGINF:
            return NEGINF
        return m + math.log(sum(math.exp(v - m) for v in a))

    alpha = np.full((T, n), NEGINF)
    alpha[0, 0] = lp[0, ext[0]]
    if n > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, T):
        for s in range(n):
            a = [alpha[t - 1, s]]
            if s >= 1:
                a.append(alpha[t - 1, s - 1])
            if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                a.append(alpha[t - 1, s - 2])
            alpha[t, s] = lse(*a) + lp[t, ext[s]]
    beta = np.full((T, n), NEGINF)
    beta[T - 1, n - 1] = lp[T - 1, ext[n - 1]]
    if n > 1:
        beta[T - 1, n - 2] = lp[T - 1, ext[n - 2]]
    for t in range(T - 2, -1, -1):
        for s in range(n):
            a = [beta[t + 1, s]]
            if s + 1 < n:
                a.append(beta[t + 1, s + 1])
            if s + 2 < n and ext[s] != blank and ext[s] != ext[s + 2]:
                a.append(beta[t + 1, s + 2])
            beta[t, s] = lse(*a) + lp[t, ext[s]]
    ll = lse(alpha[T - 1, n - 1], alpha[T - 1, n - 2]) if n > 1 else alpha[T - 1, 0]
    nll = -ll
    grad = np.exp(lp)                                  # softmax
    occ = np.zeros((T, C))
    for t in range(T):
        for s in range(n):
            v = alpha[t, s] + beta[t, s]
            if v > NEGINF:
                occ[t, ext[s]] += math.exp(v - lp[t, ext[s]] - ll)
    return nll, grad - occ


def ctc_greedy_collapse(out_enc_blc, lengths, blank=BLANK):
    """CTC best-path decode (BASELINE.json config 5): argmax, merge repeats, drop blanks."""
    am = out_enc_blc.argmax(-1)
    res = []
    for b, L in enumerate(lengths):
        seq, prev = [], -1
        for c in am[b, :L].tolist():
            if c != prev and c != blank:
                seq.append(c)
            prev = c
        res.append(seq)
    return res


def greedy_decode(sd, cfg, x_raw, lengths, max_seq_length, margins=None):
    """run_greedy, greedy_search.py:7-53 restated on token ids: encoder once (eval mode), then the full
    decoder on the growing prefix, argmax of the last step, stop when every sample has emitted </S> or
    the prefix reaches max_seq_length.  Returns list of id lists (starting with <S>) and the padded
    (B, max_seq_length) int32 tensor of greedy_search.py:41-47."""
    with torch.no_grad():
        memory, kpm = encode(sd, cfg, x_raw, lengths, False)
        B = memory.shape[0]
        dec_input = torch.full((B, 1), SOS, dtype=torch.int64)
        seqs = [[SOS] for _ in range(B)]
        while True:
            x_dec = decode(sd, cfg, dec_input, memory, kpm, False)
            logits = F.linear(x_dec, sd["w_out.weight"], sd["w_out.bias"])
            pred = torch.argmax(F.softmax(logits, dim=2), dim=2)[:, -1]
            if margins is not None:           # top-1 minus top-2 logit of the step, per sample (decode parity is only
                top2 = logits[:, -1, :].topk(2, dim=1).values    # meaningful where this exceeds the numeric tolerance)
                margins.append((top2[:, 0] - top2[:, 1]).clone())
            for i in range(B):
                if seqs[i][-1] != EOS:
                    seqs[i].append(int(pred[i]))
            dec_input = torch.cat((dec_input, pred.reshape(B, 1)), dim=1)
            if all(EOS in s for s in seqs) or dec_input.shape[1] >= max_seq_length:
                break
        out = torch.full((B, max_seq_length), PAD, dtype=torch.int32)
        for i, s in enumerate(seqs):
            out[i, :len(s)] = torch.tensor(s, dtype=torch.int32)
    return seqs, out
