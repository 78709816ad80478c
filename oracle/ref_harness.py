"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the UNMODIFIED reference (`/root/reference/speech_recognition`) in the build
container so that (1) `oracle/sst_oracle.py` (the CPU restatement that travels to the
GPU box) can be validated against the real thing and (2) `oracle/make_golden.py` can
emit the fixtures under `tests/golden/`.  `/root/reference` does not exist on the GPU
box, so nothing under `-m gpu`, `smoke()` or `bench.py` may import this module.

Three harness-side shims, zero edits to reference files:
  1. stub the absent third-party imports of data_utils.py:5-13;
  2. define FLAGS.pad (recognition_model.py:38 cannot be imported: jiwer/kenlm/dataset);
  3. replace torch-2.11's nn.TransformerEncoder/Decoder.forward with the plain layer
     loop of the torch 1.12/1.13 the reference pins (environment.yml:154,216).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("SST_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "speech_recognition")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "architecture.py"))


_loaded = {}


def load(argv=()):
    """Returns (architecture, transformer, LabelSmoothingLoss, data_utils, FLAGS)."""
    if _loaded:
        _set_flags(_loaded["FLAGS"], argv)
        return _loaded["mods"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for name in ("librosa", "soundfile", "jiwer", "num2words", "unidecode",
                 "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["num2words"].num2words = lambda *a, **k: ""
    sys.modules["unidecode"].unidecode = lambda s: s
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from absl import flags
    import architecture            # noqa: E402  (reference module)
    import transformer             # noqa: E402
    import LabelSmoothingLoss      # noqa: E402
    import data_utils              # noqa: E402
    FLAGS = flags.FLAGS
    if "pad" not in FLAGS:
        flags.DEFINE_integer("pad", 42, "Padding value (recognition_model.py:38)")
    _loaded["FLAGS"] = FLAGS
    _loaded["mods"] = (architecture, transformer, LabelSmoothingLoss, data_utils, FLAGS)
    _set_flags(FLAGS, argv)
    return _loaded["mods"]


def _set_flags(FLAGS, argv):
    FLAGS.unparse_flags()
    FLAGS(["ref_harness"] + list(argv))


def cfg_to_argv(cfg):
    return ["--model_size=%d" % cfg["d_model"],
            "--feed_forward_layer_size=%d" % cfg["d_ff"],
            "--num_layers_encoder=%d" % cfg["n_enc"],
            "--num_layers_decoder=%d" % cfg["n_dec"],
            "--n_heads_encoder=%d" % cfg["n_heads"],
            "--n_heads_decoder=%d" % cfg.get("n_heads_dec", cfg["n_heads"]),
            "--relative_distance=%d" % cfg["rel_dist"],
            "--dropout_model=%g" % cfg.get("dropout", 0.0),
            "--dropout_pos_emb=%g" % cfg.get("dropout_pos", 0.0)]


def build_model(cfg, state_dict=None):
    """Reference `architecture.Model` on CPU with the torch-1.12 container loops."""
    import torch
    architecture, transformer, _, _, FLAGS = load(cfg_to_argv(cfg))
    model = architecture.Model(112, 44, 43, "cpu")

    def enc_forward(src, mask=None, src_key_padding_mask=None, **_):
        out = src
        for mod in model.transformerEncoder.layers:
            out = mod(out, src_mask=mask, src_key_padding_mask=src_key_padding_mask)
        return out

    def dec_forward(tgt, memory, tgt_mask=None, memory_mask=None,
                    tgt_key_padding_mask=None, memory_key_padding_mask=None, **_):
        out = tgt
        for mod in model.transformerDecoder.layers:
            out = mod(out, memory, tgt_mask=tgt_mask, memory_mask=memory_mask,
                      tgt_key_padding_mask=tgt_key_padding_mask,
                      memory_key_padding_mask=memory_key_padding_mask)
        return out

    model.transformerEncoder.forward = enc_forward
    model.transformerDecoder.forward = dec_forward
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=True)
    return model
