"""Greedy attention-decoder search with the reference's interface (greedy_search.py:7-53): encoder once, then the decoder
on the growing prefix, argmax of the last position, until every sample has produced </S> or the prefix reaches the target
length.  Same return value: (list of space-joined phone strings, (B, max_seq_length) int32 id tensor padded with 42).

All model math runs in libsst.so through `Model.forward(mode='greedy_search', part=...)`.  Differences from the reference
are host-side only: the arg-max is taken on the logits (softmax is monotone, greedy_search.py:22-23) and the per-step
"has everybody finished" test is one device reduction instead of a Python loop over strings (:26-37)."""
import torch

from . import lib as L
from .data_utils import PAD

# data_utils.py:19 -- 40 phones + </S> (40), <S> (41), <PAD> (42)
phoneme_inventory = ['AA', 'AE', 'AH', 'AO', 'AW', 'AY', 'B', 'CH', 'D', 'DH', 'EH', 'ER', 'EY', 'F', 'G', 'HH', 'IH', 'IX', 'IY',
                     'JH', 'K', 'L', 'M', 'N', 'NG', 'OW', 'OY', 'P', 'R', 'S', 'SH', 'T', 'TH', 'UH', 'UW', 'V', 'W', 'Y', 'Z',
                     'ZH', '</S>', '<S>', '<PAD>']
EOS, SOS = 40, 41


class PhoneTransform(object):
    """data_utils.py:281-291."""

    def __init__(self):
        self.phoneme_inventory = phoneme_inventory
        self.vocabulary_size = len(self.phoneme_inventory)

    def phone_to_int(self, phone):
        return [self.phoneme_inventory.index(c) for c in phone]

    def int_to_phone(self, ints):
        return ''.join(self.phoneme_inventory[int(i)] for i in ints)


def greedy_ids(model, length_raw_signal, X_raw, max_seq_length, device, start_tok=SOS):
    """The id-level search.  Returns a (B, n) int64 CPU tensor of generated prefixes (column 0 = <S>), n <= max_seq_length;
    positions after a sample's first </S> hold what the decoder kept predicting (the reference keeps feeding them too,
    greedy_search.py:33-34) and are dropped by the caller."""
    memory, _ = model(length_raw_signal, device, mode='greedy_search', part='encoder', x_raw=X_raw)
    B = memory.shape[0]
    dev = memory.device
    tokens = torch.full((B, max_seq_length), PAD, dtype=torch.int64, device=dev)
    tokens[:, 0] = start_tok
    done = torch.zeros(B, dtype=torch.uint8, device=dev)
    n_done = torch.zeros(1, dtype=torch.int32, device=dev)
    n = 1
    with torch.no_grad():
        while True:
            step_logits = model(length_raw_signal, device, mode='greedy_search', part='decoder', y=tokens[:, :n], memory=memory)
            # arg-max of the last position, append, stop latch: sst_greedy_pick (softmax is monotone, greedy_search.py:22-23)
            C = step_logits.shape[2]
            L.greedy_pick(step_logits[:, n - 1], n * C, B, C, tokens, max_seq_length, 1, n, EOS, done, n_done)
            n += 1
            if n >= max_seq_length or int(n_done[0]) == B:     # the reference looks at every step, too (greedy_search.py:37)
                break
    return tokens[:, :n].cpu()


def greedy_ids_cached(model, length_raw_signal, X_raw, max_seq_length, device, start_tok=SOS):
    """Same search with key/value caches (one new decoder position per step, stop test on the device).  Falls back to the
    prefix re-run when a PAD id was generated, the one case where the two are not equivalent (Engine.greedy_cached)."""
    memory, _ = model(length_raw_signal, device, mode='greedy_search', part='encoder', x_raw=X_raw)
    eng = model._packed_engine()
    B, Lm = model._mem_shape
    with torch.no_grad():
        ids = eng.greedy_cached(memory.reshape(B * Lm, eng.D), model._mem_lens, B, Lm, max_seq_length, start_tok, EOS)
    ids = ids.cpu()
    first_eos = torch.where((ids == EOS).any(1), (ids == EOS).float().argmax(1), torch.full((B,), ids.shape[1]))
    pad_before_eos = ((ids == PAD) & (torch.arange(ids.shape[1])[None, :] <= first_eos[:, None])).any()
    if bool(pad_before_eos):
        return greedy_ids(model, length_raw_signal, X_raw, max_seq_length, device, start_tok)
    # the prefix re-run stops as soon as every sample has produced </S>; the cached loop looks every few steps only
    n = int(min(ids.shape[1], max(int(first_eos.max()) + 1, 2))) if bool((ids == EOS).any(1).all()) else ids.shape[1]
    return ids[:, :n]


def run_greedy(model, length_raw_signal, X_raw, tgt, vocab_size, device, cached=True):
    batch_len = tgt.shape[0]
    max_seq_length = tgt.shape[1] + 1                      # +1 for the removed <S> (greedy_search.py:11)
    search = greedy_ids_cached if cached else greedy_ids
    ids = search(model, length_raw_signal, X_raw, max_seq_length, device, start_tok=vocab_size - 2)
    pt = PhoneTransform()
    seqs = []
    for b in range(batch_len):
        row = ids[b].tolist()
        seq = [row[0]]
        for tok in row[1:]:
            if seq[-1] == EOS:                              # nothing is appended after </S> (:29)
                break
            seq.append(tok)
        seqs.append(seq)
    new_word_seq_idx = torch.full((batch_len, max_seq_length), PAD, dtype=torch.int32)
    for b, seq in enumerate(seqs):
        new_word_seq_idx[b, :len(seq)] = torch.tensor(seq, dtype=torch.int32)
    phones_seq = [' '.join(pt.int_to_phone([t]) for t in seq) for seq in seqs]
    return phones_seq, new_word_seq_idx.to(device)


def run_ctc_greedy(model, length_raw_signal, X_raw, device=None):
    """CTC best-path decode on the encoder's w_aux head (BASELINE.json config 5; no counterpart in the reference, whose only
    greedy search is the attention decoder above -- SURVEY.md Q16).  Returns (list of space-joined phone strings, list of id lists)."""
    model._check_input(X_raw)
    was_training = model.training
    model.eval()
    try:
        eng = model._packed_engine()
        ids, lens = eng.ctc_greedy(X_raw, [int(v) for v in length_raw_signal])
    finally:
        model.train(was_training)
    ids, lens = ids.cpu(), lens.cpu()
    seqs = [ids[b, :int(lens[b])].tolist() for b in range(ids.shape[0])]
    pt = PhoneTransform()
    return [' '.join(pt.int_to_phone([t]) for t in s) for s in seqs], seqs
