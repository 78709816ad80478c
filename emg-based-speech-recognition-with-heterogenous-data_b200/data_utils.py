"""Host-side staging of a batch for the B200 path.

`combine_fixed_length(tensor_list, length)` keeps the name and the result of the reference's packer (data_utils.py:165-174: the
utterances laid end to end along time, the last chunk completed with the VALUE 42, viewed as (n, length, channels)) because the
training loop calls it by that name (recognition_model.py:77) -- but it is a staging routine here, not a concat: the chunks are
written straight into ONE page-locked buffer, which is what the asynchronous host-to-device copy of the step wants
(`Trainer.to_device`, bench.py's e2e leg).  The reference's host-side inverse (`decollate_tensor`) has no counterpart on this
side: the split by utterance lengths happens on the device (sst_gather_rows_pad, include/sst.h).
"""
import torch

PAD = 42


class ChunkStager:
    """Reusable page-locked staging area for (n, length, channels) chunk tensors: grows to the largest batch seen, is handed out
    as a view, and is therefore only valid until the next `pack` call (one step ahead is enough for Trainer.run's pipeline when
    two stagers alternate)."""

    def __init__(self, pin=None):
        self.pin = torch.cuda.is_available() if pin is None else pin
        self.buf = None

    def _storage(self, numel, dtype):
        if self.buf is None or self.buf.numel() < numel or self.buf.dtype != dtype:
            self.buf = torch.empty(numel, dtype=dtype)
            if self.pin:
                self.buf = self.buf.pin_memory()
        return self.buf[:numel]

    def pack(self, tensor_list, length):
        first = tensor_list[0]
        inner = tuple(first.shape[1:])
        row = 1
        for s in inner:
            row *= s
        rows = sum(int(t.shape[0]) for t in tensor_list)
        n = -(-rows // length)
        flat = self._storage(n * length * row, first.dtype).view(n * length, row)
        at = 0
        for t in tensor_list:
            k = int(t.shape[0])
            flat[at:at + k].copy_(t.reshape(k, row))
            at += k
        if at < n * length:
            flat[at:].fill_(PAD)                       # the tail of the last chunk holds the value 42 (data_utils.py:170)
        return flat.view(n, length, *inner)


def combine_fixed_length(tensor_list, length, stager=None):
    """Same call and result as the reference's data_utils.combine_fixed_length; with `stager` (a ChunkStager) the result is a view
    into its page-locked buffer, otherwise a fresh (pageable) tensor."""
    if stager is not None:
        return stager.pack(tensor_list, length)
    return ChunkStager(pin=False).pack(tensor_list, length)
