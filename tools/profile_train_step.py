"""One VQGAN training step (config 5 nets, one micro-batch) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  Usage: ncu ... python tools/profile_train_step.py [--precision fp16] [--batch 8]"""
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
from text2human_b200.vqgan_arch import Discriminator  # noqa: E402
from text2human_b200.vqgan_train import VQGANTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp16")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--events", action="store_true", help="no profiler: CUDA-event timing of tap-GEMM / wgrad launches")
args = ap.parse_args()
ops.set_precision(args.precision)
dev = torch.device("cuda:0")
torch.manual_seed(5)
with contextlib.redirect_stdout(sys.stderr):
    model = VQImageSegmTextureModel(VQVAE_TOP).to(dev)
disc = Discriminator(3, 64, n_layers=3).to(dev)
tr = VQGANTrainer(model, disc, micro_batch=args.batch)
data = dict(image=R.image(300, args.batch, 3, 512, 256).to(dev), texture_mask=R.blocky_mask(300, args.batch, 512, 256, 32).to(dev))
gen = torch.Generator(device=dev).manual_seed(17)
for i in range(2):
    tr.optimize_parameters(data, 2 + i, gen)
torch.cuda.synchronize()
if args.events:
    ops.profile_tapgemm(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tr.optimize_parameters(data, 5, gen)
    e1.record()
    torch.cuda.synchronize()
    rec = ops.profile_records()
    tot = e0.elapsed_time(e1)
    by = {}
    for algo, issued, a, b, shape in rec:
        k = "wgrad" if shape[0] == "wgrad" else "tapgemm"
        t = a.elapsed_time(b)
        by.setdefault(k, [0, 0.0, 0.0])
        by[k][0] += 1; by[k][1] += t; by[k][2] += algo
    print(f"step {tot:.2f} ms (instrumented) for {args.batch} images")
    for k, (n, t, fl) in by.items():
        print(f"  {k:8s} {n:5d} launches {t:8.2f} ms  {fl / t / 1e9:8.1f} TFLOP/s algorithmic")
    # the slowest shapes
    rows = sorted(((a.elapsed_time(b), shape, algo) for algo, issued, a, b, shape in rec), key=lambda r: -r[0])[:25]
    for t, shape, algo in rows:
        print(f"   {t:7.3f} ms {algo / t / 1e9:8.1f} TF/s {shape}")
    agg = {}
    for algo, issued, a, b, shape in rec:
        e = agg.setdefault(tuple(shape), [0, 0.0])
        e[0] += 1; e[1] += a.elapsed_time(b)
    print("by shape (count, total ms):")
    for shape, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"   {n:4d} {t:8.3f} ms  {shape}")
    ops.profile_tapgemm(False)
else:
    torch.cuda.cudart().cudaProfilerStart()
    tr.optimize_parameters(data, 5, gen)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done", file=sys.stderr)
