"""Diagnostic (not a pytest file): engine vs CPU oracle, stage by stage."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, torch.nn.functional as F
import sst_oracle as O
from helpers import load_golden, golden_inputs, rel_err
from test_engine_gpu import make_engine, run_step
name = sys.argv[1] if len(sys.argv) > 1 else "short_hybrid"
dtype = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
z, meta = load_golden(name)
cfg, sd, batch = golden_inputs(meta)
eng = make_engine(cfg, sd, dtype)
X = O.combine_fixed_length(batch["raw_emg"])
# ---- oracle stages
with torch.no_grad():
    xo = X.transpose(1, 2)
    stages = []
    for i in range(3):
        xo = O.res_block(xo, sd, "conv_blocks.%d" % i, 2, True, {})
        stages.append(xo.transpose(1, 2).contiguous())
    xlin_o = F.linear(xo.transpose(1, 2), sd["w_raw_in.weight"], sd["w_raw_in.bias"])
# ---- engine stages
x_enc, ctx = eng.encode(X.to("cuda"), batch["lengths"], True, 1)
torch.cuda.synchronize()
for i, c in enumerate(ctx.blocks):
    out = c.out.float().cpu()
    lead = c.lead
    T = c.T
    print("block %d out rel err %.3e" % (i, rel_err(out[:, lead:lead + T], stages[i])))
    if i == 0:
        y1 = F.conv1d(X.transpose(1, 2), sd["conv_blocks.0.conv1.weight"], sd["conv_blocks.0.conv1.bias"], stride=2, padding=1).transpose(1, 2)
        print("   conv1 raw rel err %.3e" % rel_err(c.yc1[:, :768].float().cpu().view(y1.shape), y1))
        yr = F.conv1d(X.transpose(1, 2), sd["conv_blocks.0.residual_path.weight"], sd["conv_blocks.0.residual_path.bias"], stride=2).transpose(1, 2)
        print("   res raw rel err %.3e" % rel_err(c.yr.float().cpu().reshape(yr.shape), yr))
        h1 = F.relu(O.batch_norm(y1.transpose(1, 2), sd, "conv_blocks.0.bn1", True)).transpose(1, 2)
        print("   h1 rel err %.3e" % rel_err(c.h1p[:, 1:T + 1].float().cpu(), h1))
        y2 = F.conv1d(h1.transpose(1, 2), sd["conv_blocks.0.conv2.weight"], sd["conv_blocks.0.conv2.bias"], padding=1).transpose(1, 2)
        print("   conv2 raw rel err %.3e" % rel_err(c.yc2.float().cpu().view(y2.shape), y2))
    else:
        inp = stages[i - 1]
        pfx = "conv_blocks.%d" % i
        y1 = F.conv1d(inp.transpose(1, 2), sd[pfx + ".conv1.weight"], sd[pfx + ".conv1.bias"], stride=2, padding=1).transpose(1, 2)
        print("   conv1 raw rel err %.3e" % rel_err(c.yc1.float().cpu().view(y1.shape), y1))
        yr = F.conv1d(inp.transpose(1, 2), sd[pfx + ".residual_path.weight"], sd[pfx + ".residual_path.bias"], stride=2).transpose(1, 2)
        print("   res raw rel err %.3e" % rel_err(c.yr.float().cpu().reshape(yr.shape), yr))
res, grads, stats = O.loss_and_grads(sd, cfg, batch, True, 0)
out_enc, out_dec, loss, loss_dec, loss_enc, G, ctx = run_step(make_engine(cfg, sd, dtype), cfg, batch)
for b, l in enumerate(batch["lengths"]):
    print("out_enc[%d] rel %.3e" % (b, rel_err(out_enc[b, :l], res["out_enc"][b, :l])))
if out_dec is not None:
    print("out_dec rel %.3e" % rel_err(out_dec, res["out_dec"]))
    print("loss_dec %.6f vs %.6f" % (loss_dec, float(res["loss_dec"])))
print("loss_enc %.6f vs %.6f" % (loss_enc, float(res["loss_enc"])))
gmax = max(float(g.abs().max()) for g in grads.values())
for n in grads:
    e = rel_err(G[n], grads[n], floor=1e-4 * gmax)
    flag = " <<<<" if e > (2e-4 if dtype == torch.float32 else 4e-2) else ""
    print("grad %-70s rel %.3e  max %.3e%s" % (n, e, float(grads[n].abs().max()), flag))
