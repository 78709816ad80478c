"""sst_b200 -- B200-native hot path of the Silent Speech Transformer (EMG encoder fwd/bwd + hybrid
CTC/attention loss + data-parallel AdamW step) behind the reference's architecture.Model API.

All math runs in hand-written sm_100a CUDA kernels inside libsst.so (csrc/), reached through the C ABI
declared in include/sst.h; PyTorch only provides device memory, streams, autograd glue and
torch.distributed.  There is no CPU / eager fallback: importing `lib` without a built libsst.so, or
running on a non-sm_100 device, raises."""
__version__ = "0.1"
