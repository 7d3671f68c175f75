"""Real-clock kernel durations of one sampler forward (B=4) from torch.profiler (CUPTI), eager and under
CUDA-graph replay: separates time inside kernels from launch gaps."""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from text2human_b200 import ops  # noqa: E402
from text2human_b200.pipeline import GraphedStep, Sampler  # noqa: E402

OPT = dict(codebook_size=18432, segm_codebook_size=1024, texture_codebook_size=18, bert_n_emb=512,
           bert_n_layers=24, bert_n_head=8, block_size=512, latent_shape=[32, 16], embd_pdrop=0.0,
           resid_pdrop=0.0, attn_pdrop=0.0, num_head=18, sample_steps=256)
ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
dev = torch.device("cuda:0")
torch.manual_seed(0)
s = Sampler(OPT).to(dev).eval()
B = 4
segm = torch.randint(0, 1024, (B, 512), device=dev)
tex = torch.randint(0, 18, (B, 512), device=dev)
x_t = torch.full((B, 512), 18432, dtype=torch.long, device=dev)
fwd = s.sampler_fn.forward_logits
for _ in range(3):
    fwd(x_t, segm, tex)
g = GraphedStep(fwd, (x_t, segm, tex))
for _ in range(3):
    g(x_t, segm, tex)
torch.cuda.synchronize()
for name, fn in (("eager", fwd), ("graph", g)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        e0.record()
        fn(x_t, segm, tex)
        e1.record()
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    t_min, t_max = None, None
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and ev.name and "Memcpy" not in ev.name:
            k = ev.name.replace("void ", "").replace("t2h::", "")[:46]
            agg[k][0] += 1
            agg[k][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    tot = sum(v[1] for v in agg.values())
    print(f"== {name}: step {e0.elapsed_time(e1):.3f} ms (events), sum of kernel durations {tot / 1e3:.3f} ms, "
          f"{sum(v[0] for v in agg.values())} kernels")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"   {k:46s} {n:4d} {t / 1e3:8.3f} ms {t / n:8.2f} us each")
