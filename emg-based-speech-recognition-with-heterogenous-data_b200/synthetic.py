"""Synthetic inputs in the reference's `EMGDataset.collate_raw` output contract (read_emg.py:463-504):
dict(raw_emg=[(8*frames_i, 8) f32], lengths=[frames_i], phonemes_int=[int64 ids incl. <S>/</S>], phonemes_int_lengths).
Amplitudes follow the real pipeline's 50*tanh(x/20/50) soft clip range (read_emg.py:426-427): N(0,5) clipped to +-50."""
import torch

PAD, SOS, EOS = 42, 41, 40


def make_batch(n_utt=64, frames=1000, tgt_min=80, tgt_max=120, seed=1234, lengths=None):
    g = torch.Generator().manual_seed(seed)
    lengths = list(lengths) if lengths is not None else [frames] * n_utt
    raw, phon = [], []
    for Lf in lengths:
        raw.append((torch.randn(8 * Lf, 8, generator=g) * 5.0).clamp_(-50.0, 50.0))
        tl = int(torch.randint(tgt_min, tgt_max + 1, (1,), generator=g))
        tl = min(tl, max(1, Lf // 2))                       # keep CTC feasible for short utterances
        ids = torch.randint(0, 40, (tl,), generator=g, dtype=torch.int64)
        phon.append(torch.cat([torch.tensor([SOS]), ids, torch.tensor([EOS])]))
    return dict(raw_emg=raw, lengths=lengths, phonemes_int=phon, phonemes_int_lengths=[int(p.numel()) for p in phon])


def lognormal_lengths(n, seed=0, mu=450.0, sigma=0.5, lo=100, hi=1500):
    """Frame lengths shaped like the testset_largedev utterances."""
    g = torch.Generator().manual_seed(seed)
    x = torch.exp(torch.randn(n, generator=g) * sigma + torch.log(torch.tensor(mu)))
    return [int(v) for v in x.clamp(lo, hi).round()]
