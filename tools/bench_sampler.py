"""Config 4: transformer sampling throughput (B=4, 32x16 tokens, 24x512 transformer, 18 heads)."""
import contextlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R  # noqa: E402
from text2human_b200 import ops  # noqa: E402
from text2human_b200.pipeline import Sampler  # noqa: E402

OPT = dict(codebook_size=18432, segm_codebook_size=1024, texture_codebook_size=18, bert_n_emb=512,
           bert_n_layers=24, bert_n_head=8, block_size=512, latent_shape=[32, 16], embd_pdrop=0.0,
           resid_pdrop=0.0, attn_pdrop=0.0, num_head=18, sample_steps=256)
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ops.set_precision(prec)
if os.environ.get("T2H_SPLITK_INFER"):
    ops.set_split_k(inference=True)
dev = torch.device("cuda:0")
torch.manual_seed(0)
s = Sampler(OPT).to(dev).eval()
B = 4
segm = torch.randint(0, 1024, (B, 512), device=dev)
mask = R.blocky_mask(4, B, 512, 256, 64).to(dev)
tex = ops.mask_to_ids(mask, 32, 16).view(B, 512).long()
x_t = torch.full((B, 512), 18432, dtype=torch.long, device=dev)
for _ in range(3):
    s.sampler_fn.forward_logits(x_t, segm, tex)
torch.cuda.synchronize()
l0 = ops.COUNTERS["launches"]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    s.sampler_fn.forward_logits(x_t, segm, tex)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"[{prec}] forward_logits B=4: {ms:.3f} ms/step, {(ops.COUNTERS['launches'] - l0) // 10} launches/step, "
      f"{4 * 99.86 / ms:.1f} TFLOP/s algorithmic")
gen = torch.Generator(device=dev).manual_seed(2021)
s.sample_fn(segm, mask, sample_steps=4, generator=gen)
torch.cuda.synchronize()
t0 = time.perf_counter()
out, xt = s.sample_fn(segm, mask, sample_steps=steps, generator=gen)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"[{prec}] sample_fn {steps} steps: {dt * 1e3:.1f} ms -> {dt / steps * 1e3:.3f} ms/step; "
      f"256-step sample of 2048 tokens = {2048 / (dt / steps * 256):.0f} tokens/s (extrapolated)")
