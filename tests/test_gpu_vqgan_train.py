"""GPU parity of the VQGAN GAN training step (SURVEY a18 / BASELINE config 5) against the fixture recorded from the
REAL reference `VQImageSegmTextureModel.training_step` + `loss.backward()` + `d_loss.backward()`
(oracle/make_golden_vqgan_train.py; LPIPS stubbed): all six losses / weights, the gradient of every generator tensor
and of every discriminator tensor, in the parity precision mode; then micro-batching, Adam, checkpoint/resume."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

import golden_recipes as R

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "vqgan_train.npz")


def build(cuda, cfg=None):
    from text2human_b200.pipeline import VQImageSegmTextureModel
    from text2human_b200.vqgan_arch import Discriminator
    cfg = cfg or R.TINY_VQGAN_TRAIN
    e = cfg["enc"]
    opt = dict(embed_dim=cfg["embed_dim"], n_embed=cfg["n_embed"], double_z=False, z_channels=e["z_channels"],
               resolution=e["resolution"], in_channels=3, out_ch=3, ch=e["ch"], ch_mult=e["ch_mult"],
               num_res_blocks=e["num_res_blocks"], attn_resolutions=e["attn_resolutions"], dropout=0.0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = VQImageSegmTextureModel(opt)
    disc = Discriminator(3, cfg["ndf"], n_layers=cfg["disc_layers"])
    for name, seed in (("encoder", 101), ("decoder", 102), ("quant_conv", 103), ("post_quant_conv", 104)):
        mod = getattr(m, name)
        mod.load_state_dict(R.fill_state_dict(R.spec_of(mod), seed), strict=True)
    disc.load_state_dict(R.fill_state_dict(R.spec_of(disc), 105), strict=True)
    cb = R.codebooks(106, 18, cfg["n_embed"], cfg["embed_dim"], "trained")
    for k, emb in enumerate(m.quantize.embedding_list):
        emb.weight.data.copy_(cb[k])
    return m.to(cuda), disc.to(cuda), cfg


def recorded_draws(seed):
    """the reference's DiffAugment draws for the fixture: global CPU RNG seeded as the generating script did"""
    torch.manual_seed(seed)

    def fn(B, H, W, device):
        r = torch.cat([torch.rand(B, 1, 1, 1).view(B, 1) for _ in range(3)], 1)
        sx, sy = int(H * 0.125 + 0.5), int(W * 0.125 + 0.5)
        tx = torch.randint(-sx, sx + 1, size=[B, 1, 1]).view(B, 1)
        ty = torch.randint(-sy, sy + 1, size=[B, 1, 1]).view(B, 1)
        return r.contiguous().to(device), torch.cat([tx, ty], 1).int().contiguous().to(device)
    return fn


def _check_grads(gold, prefix, named, scale, tol):
    worst, n = 0.0, 0
    # gradients that are zero in exact arithmetic (e.g. a conv bias in front of a per-channel GroupNorm) are pure
    # rounding noise in the reference too (norms ~1e-8): they are held to an absolute floor instead
    floor = 1e-6 * max(float(gold[k]) for k in gold.files if k.startswith(prefix + "norm/"))
    for key in gold.files:
        if not key.startswith(prefix + "norm/"):
            continue
        name = key[len(prefix) + 5:]
        g = named[name].grad.detach().float() / scale
        want = float(gold[key])
        head = torch.from_numpy(gold[prefix + "head/" + name]).to(g.device)
        err = max(abs(float(g.norm()) - want), float((g.reshape(-1)[:8] - head).abs().max()))
        rel = err / max(want, 1e-9)
        assert rel <= tol or err <= floor, (name, rel, want, floor)
        worst, n = max(worst, rel if want > floor / tol else 0.0), n + 1
    return worst, n


def test_gan_training_step_matches_reference_fixture(cuda):
    mode, tol = "fp32", 1e-3
    from text2human_b200 import ops
    from text2human_b200.vqgan_train import VQGANTrainer
    gold = np.load(GOLD)
    ops.set_precision(mode)
    try:
        m, disc, cfg = build(cuda)
        tr = VQGANTrainer(m, disc, disc_start_step=cfg["disc_start_step"], disc_weight_max=1.0)
        tr.aug_draw_fn = recorded_draws(R.VQGAN_TRAIN_AUG_SEED)
        B, H, W = cfg["batch"], cfg["enc"]["resolution"], cfg["enc"]["resolution"] // 2
        data = dict(image=R.image(107, B, 3, H, W), texture_mask=R.blocky_mask(108, B, H, W, 8))
        tr.training_step(data, cfg["step"])
        torch.cuda.synchronize()
        got = tr.losses()
        ltol = 2e-4
        for k in ("nll_loss", "g_loss", "codebook_loss", "d_weight", "loss", "d_loss"):
            assert abs(got[k] - float(gold[k])) <= ltol * max(1.0, abs(float(gold[k]))), (k, got[k], float(gold[k]))
        named = {}
        for name in ("encoder", "decoder", "quant_conv", "post_quant_conv"):
            for k, p in getattr(m, name).named_parameters():
                named[f"{name}.{k}"] = p
        for k, e in enumerate(m.quantize.embedding_list):
            named[f"quantize.embedding_list.{k}.weight"] = e.weight
        worst_g, n_g = _check_grads(gold, "g", named, tr.loss_scale, tol)
        worst_d, n_d = _check_grads(gold, "d", dict(disc.named_parameters()), tr.disc_scale, tol)
        print(f"[vqgan train {mode}] {n_g} generator / {n_d} discriminator gradient tensors, worst rel err "
              f"{worst_g:.2e} / {worst_d:.2e}; losses {got}")
        assert n_g > 200 and n_d >= 10
    finally:
        ops.set_precision("fp32")


def test_single_product_mode_tracks_the_parity_mode(cuda):
    """fp16 operands / one tensor-core product per contraction (the training precision bench.py uses for config 5's
    "bf16 autocast"): losses within 2e-2 and the whole generator / discriminator gradient within a few percent (cosine)
    of the 3-product parity mode on the same batch and draws"""
    from text2human_b200 import ops
    from text2human_b200.vqgan_train import VQGANTrainer
    res = {}
    try:
        for mode in ("fp32", "fp16"):
            ops.set_precision(mode)
            m, disc, cfg = build(cuda)
            tr = VQGANTrainer(m, disc)
            tr.aug_draw_fn = recorded_draws(R.VQGAN_TRAIN_AUG_SEED)
            B, H, W = cfg["batch"], 64, 32
            data = dict(image=R.image(107, B, 3, H, W), texture_mask=R.blocky_mask(108, B, H, W, 8))
            tr.training_step(data, cfg["step"])
            res[mode] = (tr.losses(), tr.gen.flat_g.clone() / tr.loss_scale, tr.dsc.flat_g.clone() / tr.disc_scale)
    finally:
        ops.set_precision("fp32")
    (l32, g32, d32), (l16, g16, d16) = res["fp32"], res["fp16"]
    for k in ("nll_loss", "g_loss", "codebook_loss", "loss", "d_loss"):
        assert abs(l16[k] - l32[k]) <= 2e-2 * max(1.0, abs(l32[k])), (k, l16[k], l32[k])
    cg = float(torch.nn.functional.cosine_similarity(g16.double(), g32.double(), dim=0))
    cd = float(torch.nn.functional.cosine_similarity(d16.double(), d32.double(), dim=0))
    print(f"[vqgan train fp16 vs fp32] gradient cosine generator {cg:.5f} discriminator {cd:.5f}")
    assert cg > 0.98 and cd > 0.98


def test_micro_batches_adam_and_resume(cuda, tmp_path):
    """two micro-batches of 1 average to (nearly) the gradients of... themselves summed: checked against two separate
    single-image steps; one Adam step equals torch.optim.Adam on the same gradients; save -> load -> identical step"""
    from text2human_b200 import ops
    from text2human_b200.vqgan_train import VQGANTrainer
    ops.set_precision("fp32")
    m, disc, cfg = build(cuda)
    B, H, W = 2, 64, 32
    data = dict(image=R.image(7, B, 3, H, W), texture_mask=R.blocky_mask(8, B, H, W, 8))
    tr = VQGANTrainer(m, disc, micro_batch=1)
    tr.aug_draw_fn = recorded_draws(5)
    tr.training_step(data, 3)
    g_both = tr.gen.flat_g.clone()
    gsum = torch.zeros_like(g_both)
    tr.micro_batch = None
    tr.aug_draw_fn = recorded_draws(5)      # the same draw sequence, consumed image by image
    for i in range(B):
        tr.training_step({k: v[i:i + 1] for k, v in data.items()}, 3)
        gsum += tr.gen.flat_g
    rel = float((g_both - gsum).abs().max() / gsum.abs().max())
    assert rel < 1e-4, rel
    # Adam: against torch.optim.Adam fed the same (unscaled) gradients
    tr.micro_batch = 1
    tr.aug_draw_fn = recorded_draws(5)
    tr.training_step(data, 3)
    p0 = tr.gen.flat_p.clone()
    g0 = tr.gen.flat_g.clone() / (tr.n_micro * tr.loss_scale)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=tr.lr)
    ref_p.grad = g0
    opt.step()
    tr.adam_step()
    assert float((tr.gen.flat_p - ref_p.detach()).abs().max()) < 2e-7
    # checkpoint / resume with optimiser state: the next step is bit-identical parameter-wise
    path = os.path.join(tmp_path, "ck.pth")
    tr.save(path)
    tr.aug_draw_fn = recorded_draws(6)
    tr.optimize_parameters(data, 4)
    want = tr.gen.flat_p.clone()
    m2, disc2, _ = build(cuda)
    tr2 = VQGANTrainer(m2, disc2, micro_batch=1)
    tr2.load(path)
    ck = torch.load(path, map_location=cuda)
    for sp2, key in ((tr2.gen, "gen"), (tr2.dsc, "disc")):          # optimiser state restored exactly
        assert torch.equal(sp2.flat_m, ck["optimizer"][key]["flat_m"]) and torch.equal(sp2.flat_v, ck["optimizer"][key]["flat_v"])
        assert sp2.step_count == 1
    assert torch.equal(m2.decoder.conv_out.weight, ck["decoder"]["conv_out.weight"])
    tr2.aug_draw_fn = recorded_draws(6)
    tr2.optimize_parameters(data, 4)
    # the resumed step equals the uninterrupted one up to what Adam does with rounding-level gradient noise (a
    # gradient that is zero in exact arithmetic gets an update of either sign, bounded by lr)
    dev_ = (tr2.gen.flat_p - want).abs()
    assert float(dev_.max()) <= 2.5 * tr.lr and float((dev_ > 1e-6).float().mean()) < 0.01
    assert tr2.gen.step_count == tr.gen.step_count == 2
    # the mirrors' state_dict still has the reference's keys / shapes after flattening
    sd = m.decoder.state_dict()
    assert sd["conv_out.weight"].shape == (3, cfg["dec"]["ch"], 3, 3)


def test_training_step_with_and_without_the_fused_norm_backward_sums(cuda):
    """Channel counts at which the data-gradient convs run on the swapped kernel (128 / 256, as BASELINE config 5's
    levels): the whole training step with pass 1 of the norm backward in the conv epilogues (ops.nb_context, the
    default) against the same step with the stand-alone reduce launches -- every generator / discriminator gradient
    and the losses; the two differ only in fp32 summation order."""
    from text2human_b200 import ops
    from text2human_b200.vqgan_train import VQGANTrainer
    cfg = dict(enc=dict(ch=128, num_res_blocks=1, attn_resolutions=[16], in_channels=3, resolution=64, z_channels=64,
                        ch_mult=[1, 2, 2], double_z=False, dropout=0.0),
               n_embed=64, embed_dim=64, ndf=16, disc_layers=3, disc_start_step=0, step=5, batch=2)
    ops.set_precision("fp32")
    grads, losses, launches = [], [], []
    for on in (True, False):
        old = ops.FUSE_NB["on"]
        ops.FUSE_NB["on"] = on
        try:
            m, disc, _ = build(cuda, cfg)
            tr = VQGANTrainer(m, disc, disc_start_step=0, disc_weight_max=1.0)
            tr.aug_draw_fn = recorded_draws(R.VQGAN_TRAIN_AUG_SEED)
            data = dict(image=R.image(107, 2, 3, 64, 32), texture_mask=R.blocky_mask(108, 2, 64, 32, 8))
            l0 = ops.COUNTERS["launches"]
            tr.training_step(data, 5)
            torch.cuda.synchronize()
            launches.append(ops.COUNTERS["launches"] - l0)
            losses.append(tr.losses())
            grads.append((tr.gen.flat_g.clone(), tr.dsc.flat_g.clone()))
        finally:
            ops.FUSE_NB["on"] = old
    assert launches[0] < launches[1]          # reduce launches (and their memsets) are gone where the fusion applies
    for k in losses[0]:
        assert abs(losses[0][k] - losses[1][k]) <= 1e-5 * max(1.0, abs(losses[1][k])), k
    for a, b in zip(grads[0], grads[1]):
        e = float((a - b).abs().max() / b.abs().max())
        assert e < 2e-5, e
    print(f"[nb] training step launches {launches[0]} fused vs {launches[1]}; generator / discriminator gradient "
          f"difference {float((grads[0][0] - grads[1][0]).abs().max() / grads[1][0].abs().max()):.1e} / "
          f"{float((grads[0][1] - grads[1][1]).abs().max() / grads[1][1].abs().max()):.1e}")
