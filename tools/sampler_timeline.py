"""Where one graph-replayed sampler step (config 4, B=4) spends its time: run with T2H_DEBUG=16 so that CTA 0 of every
tap-GEMM / fused-attention launch records globaltimer stamps (t2h_debug_read), replay a few steps, print the per-launch
timeline of the last one: gap since the previous record's end (= the launches in between + launch latency),
dependency wait, operand latency, contraction loop, MMA drain, epilogue."""
import ctypes as C
import os
import sys

os.environ.setdefault("T2H_DEBUG", "16")
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R  # noqa: E402
from bench import SAMPLER_OPT  # noqa: E402
from text2human_b200 import _lib, ops  # noqa: E402
from text2human_b200.pipeline import Sampler  # noqa: E402

ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
dev = torch.device("cuda:0")
torch.manual_seed(0)
s = Sampler(SAMPLER_OPT).to(dev).eval()
B = 4
segm = torch.randint(0, 1024, (B, 512), device=dev)
mask = R.blocky_mask(4, B, 512, 256, 64).to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
s.sample_fn(segm, mask, sample_steps=4, generator=gen)
torch.cuda.synchronize()


def read():
    n = 1 + 2048 * 8
    buf = (C.c_longlong * n)()
    _lib.check(_lib.load().t2h_debug_read(buf, n))
    return buf[0], [list(buf[1 + 8 * i: 9 + 8 * i]) for i in range(2048)]


c0, _ = read()
s.sample_fn(segm, mask, sample_steps=6, generator=gen)
c1, ring = read()
per_step = (c1 - c0) // 6
print(f"{per_step} traced launches per step")
recs = [ring[i % 2048] for i in range(c1 - per_step, c1)]
recs.sort(key=lambda r: r[0])
print(f"{'#':>3s} {'kind':5s} {'work':>5s} {'kch':>4s} | {'gap':>6s} {'wait':>6s} {'load':>6s} {'kloop':>6s} {'drain':>6s} {'epi':>6s} | {'total':>6s}  (us)")
prev_end = None
tot = [0.0] * 7
for i, r in enumerate(recs):
    t0, t1, t2, t3, t4, t5, work, k = r
    kind = "attn" if (k >> 32) else "gemm"
    gap = (t0 - prev_end) / 1e3 if prev_end else 0.0
    vals = [gap, (t1 - t0) / 1e3, (t2 - t1) / 1e3, (t3 - t2) / 1e3, (t4 - t3) / 1e3, (t5 - t4) / 1e3, (t5 - t0) / 1e3]
    for j, v in enumerate(vals):
        tot[j] += v
    if i < 30 or i >= len(recs) - 4:
        print(f"{i:3d} {kind:5s} {work:5d} {k & 0xffffffff:4d} | " + " ".join(f"{v:6.2f}" for v in vals[:6]) + f" | {vals[6]:6.2f}")
    prev_end = t5
n = len(recs)
print("mean".ljust(20) + " | " + " ".join(f"{v / n:6.2f}" for v in tot[:6]) + f" | {tot[6] / n:6.2f}")
print(f"step span {(recs[-1][5] - recs[0][0]) / 1e3:.1f} us over {n} traced launches (+ the untraced reduce / LayerNorm / "
      f"embedding / sampling kernels in the gaps)")
