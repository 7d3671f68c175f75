"""One bench step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`.
Usage: ncu ... python tools/profile_step.py [--precision fp32|fp16] [--batch 16] [--workload vq|hier|sampler]"""
import argparse
import contextlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R  # noqa: E402
from bench import VQVAE_TOP  # noqa: E402
from text2human_b200 import ops  # noqa: E402
from text2human_b200.pipeline import VQImageSegmTextureModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()
ops.set_precision(args.precision)
dev = torch.device("cuda:0")
torch.manual_seed(2021)
with contextlib.redirect_stdout(sys.stderr):
    model = VQImageSegmTextureModel(VQVAE_TOP).to(dev).eval()
x = R.image(100, args.batch, 3, 512, 256).to(dev)
m = R.blocky_mask(100, args.batch, 512, 256, 32).to(dev)
for _ in range(args.warmup):
    model.forward_step(x, m)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
model.forward_step(x, m)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step", file=sys.stderr)
