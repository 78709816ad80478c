"""TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference `DynamicBatchSampler` (read_emg.py:144-338) in the build container on a
synthetic corpus (one `<i>_info.json` per example, as the reference reads them) and writes its batches for a few
(seed, epoch, ordering) settings to tests/golden/sampler_batches.json, so that the mirror in sst_b200/read_emg.py can be held
to the same batches on machines without /root/reference."""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402


def corpus_lengths(n=400, seed=7):
    rng = np.random.RandomState(seed)
    return [int(x) for x in np.clip(rng.lognormal(np.log(3600.0), 0.5, n), 800, 12000)]     # raw-EMG-feature frames per utterance


def texts(n):
    return ["utterance %d" % i if i % 37 else "1234 ..." for i in range(n)]                  # a few letter-free texts (filtered out)


def reference_batches(lengths, txts, settings):
    ref_harness.load()
    import read_emg as ref_read_emg                     # the reference module
    tmp = tempfile.mkdtemp()
    for i, (l, t) in enumerate(zip(lengths, txts)):
        with open(os.path.join(tmp, "%d_info.json" % i), "w") as f:
            json.dump({"chunks": [[l // 2, 0, 0], [l - l // 2, 0, 0]], "text": t}, f)

    class Dir:
        directory = tmp

    class DS:
        example_indices = [(Dir, i) for i in range(len(lengths))]

        def __len__(self):
            return len(lengths)
    out = []
    for s in settings:
        smp = ref_read_emg.DynamicBatchSampler(DS(), s["max_batch_length"], s["num_buckets"], shuffle=s["shuffle"],
                                               batch_ordering=s["ordering"], seed=s["seed"], epoch=s["epoch"],
                                               drop_last=s["drop_last"], max_batch_ex=s.get("max_batch_ex"))
        b0 = [list(map(int, b)) for b in smp]
        smp.set_epoch(s["epoch"] + 1)
        b1 = [list(map(int, b)) for b in smp]
        out.append(dict(setting=s, batches=b0, batches_next_epoch=b1, boundaries=[float(x) for x in smp._bucket_boundaries]))
    return out


SETTINGS = [dict(max_batch_length=80000, num_buckets=16, shuffle=True, ordering="random", seed=42, epoch=0, drop_last=False),
            dict(max_batch_length=80000, num_buckets=16, shuffle=True, ordering="ascending", seed=3, epoch=5, drop_last=True),
            dict(max_batch_length=40000, num_buckets=8, shuffle=False, ordering="descending", seed=42, epoch=0, drop_last=False,
                 max_batch_ex=6)]

if __name__ == "__main__":
    lengths = corpus_lengths()
    txts = texts(len(lengths))
    res = dict(lengths=lengths, texts=txts, cases=reference_batches(lengths, txts, SETTINGS))
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "sampler_batches.json")
    with open(path, "w") as f:
        json.dump(res, f)
    print("wrote", path, [len(c["batches"]) for c in res["cases"]])
