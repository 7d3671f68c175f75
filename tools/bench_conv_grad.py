"""Time the three launches of one training-step conv (forward, data gradient, weight gradient) with CUDA events,
L2-cold via rotating buffers.  Under `ncu --set full -k regex:tapgemm` this is the capture target for the
weight-gradient kernel's roofline numbers."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text2human_b200 import conv_grad as G  # noqa: E402
from text2human_b200 import ops  # noqa: E402

dev = "cuda"
iters = int(os.environ.get("ITERS", "5"))


def bench(kind, N, H, W, Ci, Co, terms):
    K = G.KSIZE[kind]
    Ho, Wo = G.out_hw(kind, H, W)
    xs = [ops.f32_to_planes(torch.randn(N, H, W, Ci, device=dev), ops.CVT_PLAIN, terms) for _ in range(2)]
    if G.STRIDE[kind] == 2:
        xs = [ops.planes_s2d(a) for a in xs]
    dys = [ops.f32_to_planes(torch.randn(N, Ho, Wo, Co, device=dev), ops.CVT_PLAIN, terms) for _ in range(2)]
    m = G.oihw_to_master(torch.randn(Co, Ci, K, K, device=dev) / (Ci * K * K) ** 0.5)
    wn, wt = G.weight_planes(m, terms)
    gw = torch.zeros_like(m)
    fl = 2.0 * N * Ho * Wo * Co * Ci * K * K
    out = {}
    for name, fn in (("fwd", lambda i: G.forward(kind, xs[i % 2], wn, None, n=N, in_hw=(H, W))),
                     ("dgrad", lambda i: G.dgrad(kind, dys[i % 2], wt, n=N, in_hw=(H, W))),
                     ("wgrad", lambda i: G.wgrad(kind, dys[i % 2], xs[i % 2], gw, n=N))):
        for i in range(2):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[name] = (ms, fl / ms / 1e9)
    return out


CASES = os.environ.get("CASES")          # e.g. CASES=0 -> only the first shape (ncu capture target)
TERMS = [int(t) for t in os.environ.get("TERMS", "1,2").split(",")]
for ci_, (kind, N, H, W, Ci, Co) in enumerate([("k3", 8, 512, 256, 128, 128), ("k3", 8, 256, 128, 128, 128), ("k3", 8, 128, 64, 256, 256),
                                ("k3", 8, 32, 16, 512, 512), ("down", 8, 512, 256, 128, 128), ("k4s2", 8, 256, 128, 64, 128)]):
    if CASES is not None and str(ci_) not in CASES.split(","):
        continue
    for terms in TERMS:
        r = bench(kind, N, H, W, Ci, Co, terms)
        print(f"{kind} {N}x{H}x{W} {Ci}->{Co} terms={terms}: " +
              "  ".join(f"{k} {v[0] * 1e3:8.1f} us {v[1]:7.1f} TF/s" for k, v in r.items()), flush=True)
