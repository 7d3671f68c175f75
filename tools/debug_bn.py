"""stress: BatchNorm(+LeakyReLU) backward kernel on the tiny discriminator's shapes, many repetitions with allocator
churn; any run-to-run difference beyond rounding is reported"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from text2human_b200 import ops
import torch.nn.functional as F
cuda = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
for shape in ((2, 7, 3, 128), (2, 8, 4, 64), (2, 16, 8, 32), (2, 32, 16, 16)):
    N, H, W, C = shape
    x = torch.randn(shape, generator=g).to(cuda)
    dy = (torch.randn(shape, generator=g) * 300).to(cuda)
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(cuda)
    beta = (0.1 * torch.randn(C, generator=g)).to(cuda)
    xr = x.permute(0, 3, 1, 2).double().requires_grad_(True)
    u = F.batch_norm(xr, None, None, gamma.double(), beta.double(), True, 0.1, 1e-5)
    F.leaky_relu(u, 0.2).backward(dy.permute(0, 3, 1, 2).double())
    want = xr.grad.permute(0, 2, 3, 1)
    outs = {}
    junk = []
    for it in range(300):
        st = ops.norm_stats(x, C, n=1)
        dx, dxp = ops.norm_bwd(x, st, gamma, beta, dy, act="lrelu", groups=C, eps=1e-5, want_planes=True, n=1, terms=2)
        err = float((dx.double() - want).abs().max() / want.abs().max())
        key = round(err, 7)
        outs[key] = outs.get(key, 0) + 1
        if it % 3 == 0:
            junk.append(torch.empty(1024 * (it % 7 + 1), device=cuda))
        if it % 11 == 0:
            junk = []
    print(shape, "distinct max-rel-errors:", sorted(outs.items())[:6])
