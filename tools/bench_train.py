"""Sampler training step (config 4's transformer: 24 x 512, 8 heads, 512 tokens, 18 x 1024-way heads):
forward + loss + backward + Adam through SamplerTrainer.optimize_parameters, timed with CUDA events.
usage: python tools/bench_train.py [fp32|fp16] [B] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R  # noqa: E402
from text2human_b200 import ops  # noqa: E402
from text2human_b200.transformer_arch import TransformerMultiHead  # noqa: E402
from text2human_b200.transformer_train import SamplerTrainer, targets_from_gt_list  # noqa: E402

CFG = dict(codebook_size=18432, segm_codebook_size=1024, texture_codebook_size=18, bert_n_emb=512,
           bert_n_layers=24, bert_n_head=8, block_size=512, latent_shape=[32, 16], embd_pdrop=0.0,
           resid_pdrop=0.0, attn_pdrop=0.0, num_head=18)
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ops.set_precision(prec)
dev = torch.device("cuda:0")
net = TransformerMultiHead(**CFG)
net.load_state_dict(R.fill_state_dict(R.spec_of(net), 5), strict=True)
net.to(dev)
tr = SamplerTrainer(net)
x_0, gt_list, segm, tex = R.sampler_train_batch(6, B=B, cfg=CFG)
x_0, segm, tex = x_0.to(dev), segm.to(dev), tex.to(dev)
own = targets_from_gt_list([g.to(dev) for g in gt_list])
gen = torch.Generator(device=dev).manual_seed(1)
losses = []
for _ in range(2):
    losses.append(float(tr.optimize_parameters(x_0, own, segm, tex, gen)[0]))
torch.cuda.synchronize()
l0 = ops.COUNTERS["launches"]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss, vb = tr.optimize_parameters(x_0, own, segm, tex, gen)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
n_par = tr.flat_p.numel()
# dense flops: 6 * params-in-GEMMs * tokens + attention (fwd 4*T*C per token per layer, x3 with backward)
gemm_par = 24 * (12 * 512 * 512) + 18432 * 512
tok = B * 512
flops = 6 * gemm_par * tok + 3 * 24 * 4 * 512 * 512 * tok
print(f"[{prec}] train step B={B}: {ms:.2f} ms/step, {tok / ms * 1e3:.0f} tokens/s, "
      f"{(ops.COUNTERS['launches'] - l0) // steps} launches/step, {flops / ms / 1e9:.1f} TFLOP/s algorithmic, "
      f"params {n_par / 1e6:.1f} M, loss {losses[0]:.4f} -> {float(loss):.4f}, "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")

if os.environ.get("T2H_TRAIN_PROFILE"):
    # per-shape time of every tap-GEMM launch of one step (events around each launch; serialises nothing)
    from collections import defaultdict
    ops.profile_tapgemm(True)
    e0.record()
    tr.optimize_parameters(x_0, own, segm, tex, gen)
    e1.record()
    torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for algo, issued, a, b, shape in ops.profile_records():
        r = agg[shape]
        r[0] += 1
        r[1] += a.elapsed_time(b)
        r[2] += issued
    ops.profile_tapgemm(False)
    tot = sum(r[1] for r in agg.values())
    print(f"tapgemm total {tot:.2f} ms of {e0.elapsed_time(e1):.2f} ms step (instrumented)")
    print(f"{'(batch, H, rows, n_out, K)':36s} {'n':>4s} {'ms':>8s} {'TF/s issued':>12s}")
    for shape, (n, ms_, issued) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{str(shape):36s} {n:4d} {ms_:8.3f} {issued / ms_ / 1e9:12.1f}")
