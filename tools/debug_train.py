"""debug: compare intermediate tensors of the VQGAN training step with the CPU oracle (same recorded draws)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R
from test_gpu_vqgan_train import build, recorded_draws
from test_vqgan_train_oracle import _state
from oracle import vqgan_train_ref as TR, vqgan_ref as V
from text2human_b200 import ops
from text2human_b200.vqgan_train import VQGANTrainer
import torch.nn.functional as F

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

cuda = torch.device("cuda:0")
ops.set_precision("fp32")
cfg, sd, sdd, cb = _state()
sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
sdd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sdd.items()}
books = [cb[k].clone().requires_grad_(True) for k in range(18)]
B, H, W = cfg["batch"], 64, 32
x, mask = R.image(107, B, 3, H, W), R.blocky_mask(108, B, H, W, 8)
# oracle pieces
torch.manual_seed(109)
h = V.conv(sd, "quant_conv", V.encoder(sd, x, "encoder."), padding=0)
quant, cbl = TR.quantize_texture_train(books, h, mask)
xrec = V.decoder(sd, V.conv(sd, "post_quant_conv", quant, padding=0), "decoder.")
xrec.retain_grad()
nll = torch.mean(torch.abs(x - xrec))
xr = TR.diff_augment(xrec); xr.retain_grad()
lf = TR.discriminator(sdd, xr, n_layers=3); lf.retain_grad()
g_loss = -torch.mean(lf)
last = sd["decoder.conv_out.weight"]
rg = torch.autograd.grad(nll, last, retain_graph=True)[0]
gg = torch.autograd.grad(g_loss, last, retain_graph=True)[0]
dxr = torch.autograd.grad(g_loss, xr, retain_graph=True)[0]
dxrec_g = torch.autograd.grad(g_loss, xrec, retain_graph=True)[0]
print("oracle |rg| %.6e |gg| %.6e d_weight %.6f" % (rg.norm(), gg.norm(), rg.norm() / (gg.norm() + 1e-4)))
# ours, step by step
m, disc, _ = build(cuda)
tr = VQGANTrainer(m, disc)
tr.aug_draw_fn = recorded_draws(109)
tr.loss_scale = S = 8192.0; tr.disc_scale = 2.0 ** 14
tr.gen.prepare(2); tr.dsc.prepare(2)
xc, mc = x.to(cuda), mask.to(cuda)
xrec_o, sqerr, zn = tr.gen_forward(xc, mc)
print("xrec rel", rel(xrec_o, xrec.detach()))
acc = torch.zeros(8, dtype=torch.float64, device=cuda)
g_nll = ops.l1_loss(xc, xrec_o, acc[0:1], gscale=S / xrec_o.numel())
r, t = tr._aug_draws(B, H, W, cuda)
xr_o = ops.diffaug_fwd(xrec_o, r, t)
print("xr rel", rel(xr_o, xr.detach()))
lf_o = tr.dnet.fwd(xr_o)
print("logits rel", rel(lf_o.view(-1), lf.detach().reshape(-1)))
d_lf = ops.hinge_loss(lf_o, acc[1:2], 0.0, gscale=-S / lf_o.numel())
d_xr = torch.empty_like(xrec_o)
tr.dnet.bwd(d_lf, want_params=False, want_input=True, dx_out=d_xr)
print("d g/d xr rel", rel(d_xr / S, dxr))
g_g = ops.diffaug_bwd(d_xr, r, t)
print("d g/d xrec rel", rel(g_g / S, dxrec_g))
co = tr.gen.convs[m.decoder.conv_out]
rgo, ggo = torch.zeros_like(co.gw), torch.zeros_like(co.gw)
tr.dec_out.wgrad_only(g_nll, rgo); tr.dec_out.wgrad_only(g_g, ggo)
from text2human_b200 import conv_grad as G
print("rg rel", rel(G.master_as_oihw(rgo, 3, 32, 3) / S, rg), "gg rel", rel(G.master_as_oihw(ggo, 3, 32, 3) / S, gg))
print("ours |rg| %.6e |gg| %.6e" % (float(rgo.norm()) / S, float(ggo.norm()) / S))

print("[debug] |xrec| %.8e |xr| %.8e |logits| %.8e |d_lf| %.8e |d_xr| %.8e |g_g| %.6e r %s t %s" % (
    float(xrec_o.norm()), float(xr_o.norm()), float(lf_o.norm()), float(d_lf.norm()), float(d_xr.norm()) / S,
    float(g_g.norm()) / S, r.tolist(), t.tolist()))
import os
os.environ["T2H_TRAIN_DEBUG"] = "1"
m2, disc2, _ = build(cuda)
tr2 = VQGANTrainer(m2, disc2, disc_start_step=0, disc_weight_max=1.0)
tr2.aug_draw_fn = recorded_draws(109)
tr2.training_step(dict(image=x, texture_mask=mask), 5)
