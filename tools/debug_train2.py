"""debug: locate the intermittent error in d g_loss / d xr: stage-by-stage comparison with the CPU oracle"""
import os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R
from test_gpu_vqgan_train import build, recorded_draws
from test_vqgan_train_oracle import _state
from oracle import vqgan_train_ref as TR, vqgan_ref as V
from text2human_b200 import ops
from text2human_b200.vqgan_train import VQGANTrainer

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

cuda = torch.device("cuda:0")
ops.set_precision("fp32")
cfg, sd, sdd, cb = _state()
B, H, W = cfg["batch"], 64, 32
x, mask = R.image(107, B, 3, H, W), R.blocky_mask(108, B, H, W, 8)
m, disc, _ = build(cuda)
tr = VQGANTrainer(m, disc)
tr.aug_draw_fn = recorded_draws(109)
S = 8192.0
tr.loss_scale = S; tr.disc_scale = 2.0 ** 14
tr.gen.prepare(2); tr.dsc.prepare(2)
xc, mc = x.to(cuda), mask.to(cuda)
xrec_o, sqerr, zn = tr.gen_forward(xc, mc)
acc = torch.zeros(8, dtype=torch.float64, device=cuda)
r, t = tr._aug_draws(B, H, W, cuda)
xr_o = ops.diffaug_fwd(xrec_o, r, t)
lf_o = tr.dnet.fwd(xr_o)
d_lf = ops.hinge_loss(lf_o, acc[1:2], 0.0, gscale=-S / lf_o.numel())
d_xr = torch.empty_like(xrec_o)
tr.dnet.debug_trace = []
tr.dnet.bwd(d_lf, want_params=False, want_input=True, dx_out=d_xr)
torch.cuda.synchronize()
trace = [(n, g.detach().cpu().double()) for n, g in tr.dnet.debug_trace]
# oracle on the SAME xr (CPU, fp64) with intermediate gradients
xr64 = xr_o.detach().cpu().double().requires_grad_(True)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sdd.items()}
inter = []
def disc_ref(sdx, x, n_layers=3):
    i = 0
    pre = F.conv2d(x, sdx[f"main.{i}.weight"], sdx[f"main.{i}.bias"], stride=2, padding=1); pre.retain_grad(); inter.append(("pre0", pre))
    h = F.leaky_relu(pre, 0.2); h.retain_grad(); inter.append(("h0", h))
    i = 2
    for n in range(1, n_layers + 1):
        stride = 2 if n < n_layers else 1
        pre = F.conv2d(h, sdx[f"main.{i}.weight"], None, stride=stride, padding=1); pre.retain_grad(); inter.append((f"pre{n}", pre))
        bn = f"main.{i + 1}."
        h = F.leaky_relu(F.batch_norm(pre, None, None, sdx[bn + "weight"], sdx[bn + "bias"], True, 0.1, 1e-5), 0.2)
        h.retain_grad(); inter.append((f"h{n}", h))
        i += 3
    return F.conv2d(h, sdx[f"main.{i}.weight"], sdx[f"main.{i}.bias"], stride=1, padding=1)
lf = disc_ref(sd64, xr64)
(-lf.mean()).backward()
want = dict((n, t_.grad.permute(0, 2, 3, 1)) for n, t_ in inter)
# our trace order: d_h_last (=h3), dpre (pre3), d_h (h2), dpre (pre2), d_h (h1), dpre (pre1), d_h (h0), dpre0 (pre0)
names = ["h3", "pre3", "h2", "pre2", "h1", "pre1", "h0", "pre0"]
for (n, g), wn in zip(trace, names):
    w_ = want[wn] * S
    ref = torch.stack((w_.sum(), w_.abs().sum(), (w_ ** 2).sum()))
    print(f"{wn:5s} ({n}) checksum rel diffs sum {float((g[0]-ref[0]).abs()/ref[1]):.2e} abs {float((g[1]-ref[1]).abs()/ref[1]):.2e} sq {float((g[2]-ref[2]).abs()/ref[2]):.2e}")
e = (d_xr.cpu().double() / S - xr64.grad).abs()
bad = e > 1e-3 * xr64.grad.abs().max()
print(f"d_xr rel {float(e.max() / xr64.grad.abs().max()):.3e} bad {int(bad.sum())}/{bad.numel()}", bad.nonzero()[:10].tolist())
