"""Batched decoder step for the reference's beam search (/root/reference/speech_recognition/BeamSearch.py:104-147).

The reference decodes all hypotheses of ONE utterance as a batch:

    memory_stub = memory.repeat(hypos.histories.shape[0], 1, 1)                                      # BeamSearch.py:111
    step_logits = model(length_raw_signal, device, mode='beam_search', part='decoder',
                        y=hypos.histories, memory=memory_stub)[:, -1, :-2]                           # BeamSearch.py:114

i.e. every step copies the (1, Lm, D) encoder memory n_hyp times and every decoder layer projects n_hyp * Lm identical rows to
cross-attention keys / values again.  `BeamDecoder` keeps that call shape (histories in, logits out, same numbers bit for bit)
but projects the memory ONCE per utterance and lets all hypotheses attend to the same key / value rows
(SstAttnDesc.k_off = 0 for every entry: include/sst.h).  The hypotheses' prefixes are still re-run in full each step: the
decoder's positional term is indexed by the hypothesis' position IN THE BATCH (transformer.py:431-435), so a
self-attention cache would not survive the re-ordering of the beam without changing the reference's numbers.
PrefixTree / KenLM scoring stay on the host with the caller (out of scope: no kenlm, no LM binary in the container).
"""
import torch

from . import lib as L
from .engine import Engine


class BeamDecoder:
    def __init__(self, model, memory, index=0):
        """model: sst_b200.architecture.Model after its part='encoder' call; memory: the (B, Lm, D) tensor that call returned;
        index: which utterance of it the search runs on (BeamSearch.run_single_bs gets batches of one)."""
        eng = model._packed_engine()
        if model._mem_lens is None or memory.dim() != 3 or memory.shape[2] != eng.D or memory.dtype != eng.dtype:
            raise L.SstError("BeamDecoder needs the memory returned by model(..., part='encoder')")
        self.model, self.eng = model, eng
        self.Lm = int(memory.shape[1])
        self.mem_len = model._mem_lens[index:index + 1]                     # int32 (1,) on the device
        mem = memory[index].contiguous()
        self.cross = eng.project_memory(mem, self.Lm)                       # once per utterance, shared by every hypothesis and step
        self._zeros = torch.zeros(0, dtype=torch.int64, device=memory.device)

    def __call__(self, histories):
        """histories: (n_hyp, t) int64 CUDA.  Returns the (n_hyp, t, num_outs_dec) fp32 logits of
        model(..., part='decoder', y=histories, memory=memory.repeat(n_hyp, 1, 1))."""
        eng, model = self.eng, self.model
        n = int(histories.shape[0])
        y = histories.contiguous()
        if self._zeros.numel() < n:
            self._zeros = torch.zeros(max(n, 128), dtype=torch.int64, device=y.device)
        tgt_pad = model.create_tgt_padding_mask(y).to(torch.uint8).contiguous()
        seeds = Engine._Seeds(model._next_seed())
        x_dec = eng.decode(y, None, None, self.mem_len.expand(n).contiguous(), n, self.Lm, model.training, seeds, tgt_pad=tgt_pad,
                           Mm=self.Lm, mem_off=self._zeros[:n], cross=self.cross)
        logits = eng.dec_head(x_dec, n * y.shape[1])
        return model._unpad_logits(logits, n, y.shape[1], eng.n_out_dec)

    def step_logits(self, histories):
        """The quantity BeamSearch.py:114 feeds to log_softmax: last position, without the <S> and <PAD> classes."""
        return self(histories)[:, -1, :-2]
