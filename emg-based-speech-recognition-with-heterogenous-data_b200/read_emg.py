"""Batching contract either side of the hot path: the reference's length-bucketed batch sampler and
collate function, re-stated without the dataset files, plus the data-parallel extension the reference lacks.

Mirrors `speech_recognition/read_emg.py`:
  * `DynamicBatchSampler` (read_emg.py:144-338): lognormal-warped bucket boundaries (`_get_boundaries_through_warping`,
    :220-238), `bucket_lens = max(1, int(max_batch_length / boundary))` (+ a last bucket of one), examples shuffled with
    `torch.randperm(generator seeded seed + epoch)`, an example goes to bucket `searchsorted(boundaries, length)`, a bucket
    is emitted when it holds `bucket_lens[b]` (or `max_batch_ex`) examples, leftovers are dumped unless `drop_last`, batches
    are then reordered (`random` / `ascending` / `descending`).  Same constructor, same batches for the same seed.
    The reference reads every example's length (and skips texts without letters) from `<idx>_info.json` files; here the
    lengths come from `lengths_list` (or `dataset.lengths_list` / the same json files when the dataset has them) and the
    optional `keep` predicate replaces the text filter.
  * `collate_raw` (read_emg.py:463-504): same dictionary of per-example lists.

Data-parallel extension (`rank`, `world_size`): the gradient all-reduce makes a step as slow as its slowest rank, so the
ranks of one step must carry similar work.  Batches of one bucket hold similar lengths and the same example count, so each
global step deals `world_size` batches OF THE SAME BUCKET to the ranks (a bucket's trailing batches that do not fill a step
are merged with the next bucket's); every rank sees `len(sampler)` batches per epoch and the union over ranks is a
partition of the reference's batch list.  `world_size = 1` is exactly the reference sampler.
"""
import json
import os
import string
from typing import List, Optional

import numpy as np
import torch

__all__ = ["DynamicBatchSampler", "collate_raw", "lognormal_boundaries"]


def _lognorm_ppf(q):
    """scipy.stats.lognorm.ppf(q, 1) = exp(norm.ppf(q)) without scipy (read_emg.py:232)."""
    return np.exp(np.sqrt(2.0) * _erfinv(2.0 * np.asarray(q, dtype=np.float64) - 1.0))


def _erfinv(y):
    try:
        from scipy.special import erfinv
        return erfinv(y)
    except Exception:                                       # pragma: no cover - scipy is present in the image
        return torch.erfinv(torch.as_tensor(y, dtype=torch.float64)).numpy()


def lognormal_boundaries(max_batch_length: int, num_quantiles: int) -> List[float]:
    """read_emg.py:220-238."""
    num_boundaries = num_quantiles + 1
    latent = np.linspace(1 / num_boundaries, num_quantiles / num_boundaries, num_quantiles)
    quantiles = _lognorm_ppf(latent)
    return list(sorted(quantiles * max_batch_length / quantiles[-1]))


class DynamicBatchSampler(torch.utils.data.Sampler):
    def __init__(self, dataset, max_batch_length: int, num_buckets: Optional[int] = None, shuffle: bool = True,
                 batch_ordering: str = "random", max_batch_ex: Optional[int] = None, bucket_boundaries: List[int] = [],
                 seed: int = 42, epoch: int = 0, drop_last: bool = False, verbose: bool = False,
                 lengths_list: Optional[List[int]] = None, keep=None, rank: int = 0, world_size: int = 1):
        self._dataset = dataset
        self.verbose = verbose
        if lengths_list is None:
            lengths_list = getattr(dataset, "lengths_list", None)
        self._texts = None
        if lengths_list is None:                            # the reference's on-disk form (read_emg.py:164-169)
            lengths_list, self._texts = [], []
            for directory_info, file_idx in dataset.example_indices:
                with open(os.path.join(directory_info.directory, f"{file_idx}_info.json")) as f:
                    info = json.load(f)
                lengths_list.append(sum(emg_len for emg_len, _, _ in info["chunks"]))
                self._texts.append(info["text"])
        self.lengths_list = list(lengths_list)
        self._ex_lengths = {str(i): l for i, l in enumerate(self.lengths_list)}
        self._keep = keep

        if len(bucket_boundaries) > 0:
            if not all(x >= 0 for x in bucket_boundaries):
                raise ValueError("All elements in bucket boundaries should be non-negative (>= 0).")
            if not len(set(bucket_boundaries)) == len(bucket_boundaries):
                raise ValueError("Bucket_boundaries should not contain duplicates.")
            np.testing.assert_array_equal(np.array(bucket_boundaries), np.array(sorted(bucket_boundaries)),
                                          err_msg="The arg bucket_boundaries should be an ascending sorted list of non negative values values!")
            self._bucket_boundaries = np.array(sorted(bucket_boundaries))
        else:
            self._bucket_boundaries = np.array(lognormal_boundaries(max_batch_length, num_buckets))

        self._max_batch_length = max_batch_length
        self._shuffle_ex = shuffle
        self._batch_ordering = batch_ordering
        self._seed = seed
        self._drop_last = drop_last
        self._max_batch_ex = np.inf if max_batch_ex is None else max_batch_ex
        self._bucket_lens = [max(1, int(max_batch_length / b)) for b in self._bucket_boundaries] + [1]
        self._epoch = epoch
        if not (0 <= rank < world_size):
            raise ValueError("rank %d outside world of %d" % (rank, world_size))
        self._rank, self._world = rank, world_size
        self._generate_batches()

    # ---- reference API --------------------------------------------------------------------------------------------
    def get_durations(self, batch):
        return [self._ex_lengths[str(idx)] for idx in batch]

    def _kept(self, idx):
        if self._keep is not None:
            return bool(self._keep(idx))
        if self._texts is not None:                          # read_emg.py:292 -- skip utterances without any letter
            return any(ch in string.ascii_letters for ch in self._texts[idx])
        return True

    def _permute_batches(self):
        if self._batch_ordering == "random":
            g = torch.Generator()
            g.manual_seed(self._seed + self._epoch)
            order = torch.randperm(len(self._batches), generator=g).tolist()
            self._batches = [self._batches[i] for i in order]
            self._batch_bucket = [self._batch_bucket[i] for i in order]
        elif self._batch_ordering in ("ascending", "descending"):
            order = sorted(range(len(self._batches)), key=lambda i: max(self._ex_lengths[str(idx)] for idx in self._batches[i]),
                           reverse=self._batch_ordering == "descending")
            self._batches = [self._batches[i] for i in order]
            self._batch_bucket = [self._batch_bucket[i] for i in order]
        else:
            raise NotImplementedError

    def _generate_batches(self):
        n = len(self.lengths_list)
        if self._shuffle_ex:
            g = torch.Generator()
            g.manual_seed(self._seed + self._epoch)
            sampler = torch.randperm(n, generator=g).tolist()
        else:
            sampler = range(n)
        self._batches, self._batch_bucket = [], []
        bucket_batches = [[] for _ in self._bucket_lens]
        for idx in sampler:
            if not self._kept(idx):
                continue
            b = int(np.searchsorted(self._bucket_boundaries, self._ex_lengths[str(idx)]))
            bucket_batches[b].append(idx)
            if len(bucket_batches[b]) >= self._bucket_lens[b] or len(bucket_batches[b]) >= self._max_batch_ex:
                self._batches.append(bucket_batches[b])
                self._batch_bucket.append(b)
                bucket_batches[b] = []
        if not self._drop_last:
            for b, batch in enumerate(bucket_batches):
                if batch:
                    self._batches.append(batch)
                    self._batch_bucket.append(b)
        self._permute_batches()
        self._shard()

    # ---- data-parallel sharding -----------------------------------------------------------------------------------
    def _shard(self):
        """Global steps of `world` batches with equal bucket ids; this rank keeps entry `rank` of every step."""
        self._global_batches = list(self._batches)
        self._steps = []
        if self._world == 1:
            return
        W = self._world
        by_bucket = {}
        for pos, b in enumerate(self._batch_bucket):          # keep the (permuted) order inside a bucket
            by_bucket.setdefault(b, []).append(pos)
        steps, carry = [], []                                 # carry: leftovers of smaller buckets, merged upwards
        first_pos = {}
        for b in sorted(by_bucket):
            pool = carry + by_bucket[b]
            while len(pool) >= W:
                grp, pool = pool[:W], pool[W:]
                first_pos[len(steps)] = min(grp)
                steps.append(grp)
            carry = pool
        # the final partial step is dropped (every rank must run the same number of steps)
        order = sorted(range(len(steps)), key=lambda s: first_pos[s])          # steps follow the reference's batch order
        self._steps = [steps[s] for s in order]
        mine = []
        for grp in self._steps:
            # heaviest batch to rank 0, ...: deterministic and the same on every rank
            grp = sorted(grp, key=lambda p: (-self._frames(self._global_batches[p]), p))
            mine.append(self._global_batches[grp[self._rank]])
        self._batches = mine

    def _frames(self, batch):
        return sum(self._ex_lengths[str(i)] for i in batch)

    def step_imbalance(self):
        """max over ranks / mean over ranks of the frames a step carries, averaged over steps (1.0 = perfectly balanced)."""
        if self._world == 1 or not self._steps:
            return 1.0
        r = []
        for grp in self._steps:
            f = [self._frames(self._global_batches[p]) for p in grp]
            r.append(max(f) / (sum(f) / len(f)))
        return float(np.mean(r))

    def __iter__(self):
        for batch in self._batches:
            yield batch

    def set_epoch(self, epoch):
        self._epoch = epoch
        if self._shuffle_ex:
            self._generate_batches()

    def __len__(self):
        return len(self._batches)


def collate_raw(batch):
    """EMGDataset.collate_raw (read_emg.py:463-504): a dictionary of per-example lists, nothing padded or stacked."""
    audio_features, audio_feature_lengths, parallel_emg = [], [], []
    for ex in batch:
        if ex["silent"]:
            audio_features.append(ex["parallel_voiced_audio_features"])
            audio_feature_lengths.append(ex["parallel_voiced_audio_features"].shape[0])
            parallel_emg.append(ex["parallel_voiced_emg"])
        else:
            audio_features.append(ex["audio_features"])
            audio_feature_lengths.append(ex["audio_features"].shape[0])
            parallel_emg.append(np.zeros(1))
    return {"audio_features": audio_features,
            "audio_feature_lengths": audio_feature_lengths,
            "emg": [ex["emg"] for ex in batch],
            "raw_emg": [ex["raw_emg"] for ex in batch],
            "parallel_voiced_emg": parallel_emg,
            "phonemes": [ex["phonemes"] for ex in batch],
            "phonemes_int": [ex["phonemes_int"] for ex in batch],
            "phonemes_int_lengths": [ex["phonemes_int"].shape[0] for ex in batch],
            "session_ids": [ex["session_ids"] for ex in batch],
            "lengths": [ex["emg"].shape[0] for ex in batch],
            "silent": [ex["silent"] for ex in batch],
            "text": [ex["text"] for ex in batch],
            "text_int": [ex["text_int"] for ex in batch],
            "text_int_lengths": [ex["text_int"].shape[0] for ex in batch]}
