"""CPU: the GAN-training-step restatement (oracle/vqgan_train_ref.training_step) against the fixture made from the
real reference `VQImageSegmTextureModel.training_step` + `loss.backward()` + `d_loss.backward()`
(oracle/make_golden_vqgan_train.py).  The CUDA path for this step is not built yet; this
pins the oracle it will be measured against."""
import contextlib
import io
import os

import numpy as np
import torch

import golden_recipes as R
from oracle import vqgan_train_ref as TR

GOLD = os.path.join(os.path.dirname(__file__), "golden", "vqgan_train.npz")


def _state():
    """the same seeded weights the generator script loaded into the reference modules, keyed as in state_dict()"""
    from text2human_b200.vqgan_arch import Decoder, Encoder
    cfg = R.TINY_VQGAN_TRAIN
    with contextlib.redirect_stdout(io.StringIO()):
        enc, dec = Encoder(**cfg["enc"]), Decoder(**cfg["dec"])
    sd = {}
    for prefix, mod, seed in (("encoder.", enc, 101), ("decoder.", dec, 102)):
        for k, v in R.fill_state_dict(R.spec_of(mod), seed).items():
            sd[prefix + k] = v
    qc = torch.nn.Conv2d(cfg["enc"]["z_channels"], cfg["embed_dim"], 1)
    pq = torch.nn.Conv2d(cfg["embed_dim"], cfg["enc"]["z_channels"], 1)
    for prefix, mod, seed in (("quant_conv.", qc, 103), ("post_quant_conv.", pq, 104)):
        for k, v in R.fill_state_dict(R.spec_of(mod), seed).items():
            sd[prefix + k] = v
    # the discriminator's keys (nn.Sequential 'main.N'), built from its published layer list
    ndf, spec = cfg["ndf"], []
    spec += [("main.0.weight", (ndf, 3, 4, 4)), ("main.0.bias", (ndf,))]
    i, prev = 2, 1
    for n in range(1, cfg["disc_layers"] + 1):
        mult = min(2 ** n, 8)
        spec += [(f"main.{i}.weight", (ndf * mult, ndf * prev, 4, 4)), (f"main.{i + 1}.weight", (ndf * mult,)),
                 (f"main.{i + 1}.bias", (ndf * mult,)), (f"main.{i + 1}.running_mean", (ndf * mult,)),
                 (f"main.{i + 1}.running_var", (ndf * mult,)), (f"main.{i + 1}.num_batches_tracked", ())]
        prev, i = mult, i + 3
    spec += [(f"main.{i}.weight", (1, ndf * prev, 4, 4)), (f"main.{i}.bias", (1,))]
    sdd = R.fill_state_dict(spec, 105)
    cb = R.codebooks(106, 18, cfg["n_embed"], cfg["embed_dim"], "trained")
    return cfg, sd, sdd, cb


def test_gan_training_step_restatement_matches_reference_fixture():
    gold = np.load(GOLD)
    cfg, sd, sdd, cb = _state()
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    sdd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
           for k, v in sdd.items()}
    books = [cb[k].clone().requires_grad_(True) for k in range(18)]
    B, H, W = cfg["batch"], cfg["enc"]["resolution"], cfg["enc"]["resolution"] // 2
    x, mask = R.image(107, B, 3, H, W), R.blocky_mask(108, B, H, W, 8)
    torch.manual_seed(R.VQGAN_TRAIN_AUG_SEED)          # the reference's DiffAugment draws, replayed in the same order
    r = TR.training_step(sd, books, sdd, x, mask, cfg["step"], disc_start_step=cfg["disc_start_step"],
                         disc_layers=cfg["disc_layers"])
    for k in ("loss", "d_loss", "nll_loss", "g_loss", "d_weight", "codebook_loss"):
        assert abs(float(r[k]) - float(gold[k])) <= 2e-5 * max(1.0, abs(float(gold[k]))), k
    r["loss"].backward(retain_graph=True)
    n_checked = 0
    for key in gold.files:
        if not key.startswith("gnorm/"):
            continue
        name = key[len("gnorm/"):]
        if name.startswith("quantize.embedding_list."):
            g = books[int(name.split(".")[2])].grad
        else:
            g = sd[name].grad
        g = g if g is not None else torch.zeros(1)
        want = float(gold[key])
        assert abs(float(g.norm()) - want) <= 2e-4 * max(want, 1e-6) + 1e-7, name
        head = torch.from_numpy(gold["ghead/" + name])
        assert (g.reshape(-1)[:8] - head).abs().max() <= 2e-4 * max(want, 1e-6) + 1e-7, name
        n_checked += 1
    assert n_checked > 200
    for p in sdd.values():
        if p.requires_grad:
            p.grad = None                     # disc_optimizer.zero_grad() before d_loss.backward()
    r["d_loss"].backward()
    for key in gold.files:
        if key.startswith("dnorm/"):
            name = key[len("dnorm/"):]
            g, want = sdd[name].grad, float(gold[key])
            assert abs(float(g.norm()) - want) <= 2e-4 * max(want, 1e-6) + 1e-7, name
            assert (g.reshape(-1)[:8] - torch.from_numpy(gold["dhead/" + name])).abs().max() <= 2e-4 * want + 1e-7
