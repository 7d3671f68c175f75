"""Generate tests/golden/vqgan_train.npz by running the REAL reference GAN training step
(/root/reference/models/vqgan_model.py VQImageSegmTextureModel.training_step / forward_step / encode / decode,
unbound, on a stand-in ``self`` that carries the attributes they read) around the REAL reference Encoder / Decoder /
VectorQuantizerTexture / Discriminator and the REAL models/losses/vqgan_loss.py, followed by ``loss.backward()`` and
``d_loss.backward()``.  `lpips` is absent (and would download VGG weights): a stand-in module whose LPIPS returns
zeros is registered, as BASELINE config 5 prescribes ("LPIPS stubbed").

The fixture stores losses, the adaptive weight and, per parameter tensor, the gradient's norm and its first 8
entries (full gradients of even the reduced nets would be tens of MB).
Run in the build container only:  python oracle/make_golden_vqgan_train.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R  # noqa: E402

REF = "/root/reference/models"
OUT = os.path.join(ROOT, "tests", "golden")


def _load(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def build():
    for name in ("models", "models.archs", "models.losses"):
        sys.modules.setdefault(name, types.ModuleType(name))
    lp = types.ModuleType("lpips")

    class LPIPS(torch.nn.Module):
        def __init__(self, net="vgg"):
            super().__init__()

        def forward(self, a, b):
            return torch.zeros(a.size(0), 1, 1, 1)
    lp.LPIPS = LPIPS
    sys.modules["lpips"] = lp
    vq = _load("models.archs.vqgan_arch", os.path.join(REF, "archs", "vqgan_arch.py"))
    _load("models.losses.vqgan_loss", os.path.join(REF, "losses", "vqgan_loss.py"))
    _load("models.losses.segmentation_loss", os.path.join(REF, "losses", "segmentation_loss.py"))
    vm = _load("models.vqgan_model", os.path.join(REF, "vqgan_model.py"))
    return vq, vm


def make_fake(vq, vm, cfg):
    Model = vm.VQImageSegmTextureModel
    fake = types.SimpleNamespace()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        fake.encoder = vq.Encoder(**cfg["enc"])
        fake.decoder = vq.Decoder(**cfg["dec"])
    fake.quantize = vq.VectorQuantizerTexture(cfg["n_embed"], cfg["embed_dim"], beta=0.25)
    fake.quant_conv = torch.nn.Conv2d(cfg["enc"]["z_channels"], cfg["embed_dim"], 1)
    fake.post_quant_conv = torch.nn.Conv2d(cfg["embed_dim"], cfg["enc"]["z_channels"], 1)
    fake.disc = vq.Discriminator(3, cfg["ndf"], n_layers=cfg["disc_layers"])
    import lpips
    fake.perceptual = lpips.LPIPS(net="vgg")
    fake.perceptual_weight, fake.disc_start_step, fake.disc_weight_max = 1.0, cfg["disc_start_step"], 1.0
    fake.diff_aug, fake.policy, fake.device, fake.log_dict = True, "color,translation", "cpu", {}
    for name in ("feed_data", "encode", "decode", "forward_step", "training_step"):
        setattr(fake, name, types.MethodType(getattr(Model, name), fake))
    return fake


def main():
    torch.set_num_threads(8)
    vq, vm = build()
    cfg = R.TINY_VQGAN_TRAIN
    fake = make_fake(vq, vm, cfg)
    for name, seed in (("encoder", 101), ("decoder", 102), ("quant_conv", 103), ("post_quant_conv", 104), ("disc", 105)):
        mod = getattr(fake, name)
        mod.load_state_dict(R.fill_state_dict(R.spec_of(mod), seed), strict=True)
    cb = R.codebooks(106, 18, cfg["n_embed"], cfg["embed_dim"], "trained")
    for k, e in enumerate(fake.quantize.embedding_list):
        e.weight.data.copy_(cb[k])
    for m in (fake.encoder, fake.decoder, fake.quantize, fake.quant_conv, fake.post_quant_conv, fake.disc):
        m.train()
    B, H, W = cfg["batch"], cfg["enc"]["resolution"], cfg["enc"]["resolution"] // 2
    data = dict(image=R.image(107, B, 3, H, W), texture_mask=R.blocky_mask(108, B, H, W, 8))
    torch.manual_seed(R.VQGAN_TRAIN_AUG_SEED)
    loss, d_loss = fake.training_step(data, cfg["step"])
    gen_params = {}
    for name in ("encoder", "decoder", "quant_conv", "post_quant_conv"):
        for k, p in getattr(fake, name).named_parameters():
            gen_params[f"{name}.{k}"] = p
    for k, e in enumerate(fake.quantize.embedding_list):
        gen_params[f"quantize.embedding_list.{k}.weight"] = e.weight
    loss.backward()
    out = {"loss": loss.detach().numpy(), "d_loss": d_loss.detach().numpy()}
    for k in ("nll_loss", "g_loss", "codebook_loss"):
        out[k] = np.float32(fake.log_dict[k])
    out["d_weight"] = fake.log_dict["d_weight"].numpy()
    for k, p in gen_params.items():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out["gnorm/" + k] = g.norm().numpy()
        out["ghead/" + k] = g.reshape(-1)[:8].numpy().copy()
    fake.disc.zero_grad()          # disc_optimizer.zero_grad() before d_loss.backward() (vqgan_model.py:341-343)
    d_loss.backward()
    for k, p in fake.disc.named_parameters():
        out["dnorm/" + k] = p.grad.norm().numpy()
        out["dhead/" + k] = p.grad.reshape(-1)[:8].numpy().copy()
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "vqgan_train.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "loss", float(loss), "d_loss", float(d_loss), "d_weight",
          float(out["d_weight"]), "codebook", float(out["codebook_loss"]), len(gen_params), "generator tensors")


if __name__ == "__main__":
    main()
