"""debug: Discriminator forward + input-gradient through DiscNet vs torch autograd (fp64), repeated, to catch
run-to-run differences.  usage: python tools/debug_disc.py [sync]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R
from oracle import vqgan_train_ref as TR
from text2human_b200 import ops
from text2human_b200.vqgan_arch import Discriminator
from text2human_b200 import vqgan_autograd as VA, vqgan_train as VT

sync = len(sys.argv) > 1 and sys.argv[1] == "sync"
if sync:
    _orig = ops._count
    def _c(n=1):
        torch.cuda.synchronize()
        _orig(n)
    ops._count = _c
cuda = torch.device("cuda:0")
ops.set_precision("fp32")
disc = Discriminator(3, 16, n_layers=3)
sd = R.fill_state_dict(R.spec_of(disc), 105)
disc.load_state_dict(sd, strict=True)
disc = disc.to(cuda)
x = (R.image(5, 2, 3, 64, 32) * 0.8)
x64 = x.double().requires_grad_(True)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
lf = TR.discriminator(sd64, x64, n_layers=3)
(-lf.mean()).backward()
want = x64.grad
host = VA._Host(want_grads=False)
net = VT.DiscNet(host, disc)
xc = x.to(cuda)
for it in range(4):
    logits = net.fwd(xc)
    if sync: torch.cuda.synchronize()
    dl = torch.full_like(logits, -341.0 / 1.0)
    dx = torch.empty_like(xc)
    net.bwd(dl, want_params=False, want_input=True, dx_out=dx)
    torch.cuda.synchronize()
    got = dx.double().cpu() / (341.0 * logits.numel())
    err = float((got - want).abs().max() / want.abs().max())
    lerr = float((logits.double().cpu().view(-1) - lf.detach().reshape(-1)).abs().max() / lf.abs().max())
    print(f"iter {it} sync={sync} PDL={os.environ.get('T2H_PDL','1')}: logits rel {lerr:.2e}  dx rel {err:.3e}  |dx| {float(dx.norm()):.8e}")
