"""Driver for the per-family DRAM-traffic capture (not a pytest file): warms the cfg2 training step up, then runs ONE step between
cudaProfilerStart/Stop and writes the issue-ordered list of C-ABI calls (kernel family per call) next to the ncu CSV:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \\
        --csv --log-file gpurun_out/r02_step_metrics.csv python tools/traffic_step.py gpurun_out/r02_step_kinds.json
    python tools/kernel_traffic.py gpurun_out/r02_step_metrics.csv gpurun_out/r02_step_kinds.json      (here, no GPU needed)"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L
from sst_b200 import architecture as A
from sst_b200.synthetic import make_batch
from sst_b200.train import Trainer
import bench
w = bench.WORKLOADS["cfg2"]
dev = torch.device("cuda", 0)
A.configure(model_size=768, feed_forward_layer_size=3072, num_layers_encoder=w["n_enc"], num_layers_decoder=w["n_dec"], n_heads_encoder=8,
            n_heads_decoder=8, relative_distance=100, dropout_model=0.2, dropout_pos_emb=0.2, sst_dtype="bf16")
torch.manual_seed(0)
model = A.Model(112, 44, 43, dev).to(dev)
tr = Trainer(model, alpha_loss=w["alpha"], batch_size_grad=1, seed=0)
d = tr.to_device(tr.prepare(make_batch(w["n_utt"], w["frames"], w["tgt"][0], w["tgt"][1], seed=1234)))
pristine = d["X"].clone()
for _ in range(3):
    d["X"].copy_(pristine)
    tr.step_device(d)
torch.cuda.synchronize()
d["X"].copy_(pristine)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
with L.Profiler() as prof:
    tr.step_device(d)
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
calls = [{"kind": k, "flops": f, "bytes": b, "tag": t} for k, f, b, _, _, t in prof.records]
json.dump({"workload": "cfg2", "calls": calls}, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_step_kinds.json", "w"))
print("recorded %d C-ABI calls" % len(calls))
