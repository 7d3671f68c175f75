"""Config 4: two transformer forwards of the sampler (B=4) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R  # noqa: E402
from bench import SAMPLER_OPT  # noqa: E402
from text2human_b200 import ops  # noqa: E402
from text2human_b200.pipeline import Sampler  # noqa: E402

ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
dev = torch.device("cuda:0")
torch.manual_seed(0)
s = Sampler(SAMPLER_OPT).to(dev).eval()
B = 4
segm = torch.randint(0, 1024, (B, 512), device=dev)
mask = R.blocky_mask(4, B, 512, 256, 64).to(dev)
tex = ops.mask_to_ids(mask, 32, 16).view(B, 512).long()
x_t = torch.full((B, 512), 18432, dtype=torch.long, device=dev)
for _ in range(3):
    s.sampler_fn.forward_logits(x_t, segm, tex)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(2):
    s.sampler_fn.forward_logits(x_t, segm, tex)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
