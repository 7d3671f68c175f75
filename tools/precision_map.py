"""Per-layer precision map: for every conv / 1x1 of the vqvae_top encoder+decoder, the end-to-end
pixel error (max-norm relative to the fp32 result, teacher-forced decoder input for decoder layers, latent error for
encoder layers) when ONLY that layer runs with single-product operands (both operands rounded to fp16, fp32
accumulate -- what one tensor-core product computes), on the oracle restatement, CPU.  Then the largest set of layers
(greedy by FLOPs per squared error) whose combined error stays under the budget.
usage: python tools/precision_map.py [H W]"""
import contextlib, io, os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R
from oracle import vqgan_ref as V
from bench import VQVAE_TOP
from text2human_b200.pipeline import VQImageSegmTextureModel

torch.set_num_threads(8)
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 128)
torch.manual_seed(2021)
with contextlib.redirect_stdout(io.StringIO()):
    m = VQImageSegmTextureModel(VQVAE_TOP).eval()
sd = {k: v.detach() for k, v in m.state_dict().items()}
cb = R.codebooks(7, 18, 1024, 256, "trained")
x = R.image(2021, 1, 3, H, W)
mask = R.blocky_mask(2021, 1, H, W, 32)
ROUND = set()
_orig_conv = V.conv
FLOPS = {}

def conv(sdx, name, xx, stride=1, padding=1):
    w, b = sdx[name + ".weight"], sdx[name + ".bias"]
    if name in ROUND:
        xx, w = xx.half().float(), w.half().float()
    y = F.conv2d(xx, w, b, stride=stride, padding=padding)
    FLOPS[name] = 2.0 * y.numel() * w.shape[1] * w.shape[2] * w.shape[3]
    return y
V.conv = conv

def run():
    with torch.no_grad():
        return V.vq_forward_step(sd, cb, x, mask)
base = run()
q_in = base["quant"]

def dec_only():
    with torch.no_grad():
        return V.decoder(sd, V.conv(sd, "post_quant_conv", q_in, padding=0), "decoder.")
names = [k[:-7] for k in sd if k.endswith(".weight") and sd[k].dim() == 4]
rows = []
dec_base = dec_only()
for n in names:
    ROUND.clear(); ROUND.add(n)
    if n.startswith("encoder") or n == "quant_conv":
        with torch.no_grad():
            z = V.conv(sd, "quant_conv", V.encoder(sd, x, "encoder."), padding=0)
        err = float((z - base["z"]).abs().max() / base["z"].abs().max())
        kind = "z"
    else:
        d = dec_only()
        err = float((d - dec_base).abs().max() / dec_base.abs().max())
        kind = "px"
    rows.append((n, kind, err, FLOPS.get(n, 0.0)))
ROUND.clear()
tot = sum(r[3] for r in rows)
print(f"{'layer':52s} kind   err(single)   GFLOP   share")
for n, kind, err, fl in rows:
    print(f"{n:52s} {kind:3s} {err:12.3e} {fl / 1e9:8.2f} {100 * fl / tot:6.2f}%")
# all single
ROUND.update(names)
allr = run()
print("ALL single-product: z err %.3e px err %.3e" % (float((allr['z'] - base['z']).abs().max() / base['z'].abs().max()),
                                                      float((dec_only() - dec_base).abs().max() / dec_base.abs().max())))
for budget, kind in ((7e-4, "px"), (7e-4, "z")):
    cand = sorted([r for r in rows if r[1] == kind], key=lambda r: -(r[3] / max(r[2], 1e-12) ** 2))
    chosen, acc = [], 0.0
    for n, k_, err, fl in cand:
        if (acc + err ** 2) ** 0.5 <= budget:
            chosen.append(n); acc += err ** 2
    ROUND.clear(); ROUND.update(chosen)
    if kind == "px":
        e = float((dec_only() - dec_base).abs().max() / dec_base.abs().max())
    else:
        with torch.no_grad():
            z = V.conv(sd, "quant_conv", V.encoder(sd, x, "encoder."), padding=0)
        e = float((z - base["z"]).abs().max() / base["z"].abs().max())
    fl = sum(r[3] for r in rows if r[0] in chosen)
    print(f"[{kind}] quadrature-greedy set under {budget:g}: {len(chosen)} layers, {100 * fl / tot:.1f}% of all conv FLOPs, "
          f"measured combined err {e:.3e}")
    print("   ", chosen)
