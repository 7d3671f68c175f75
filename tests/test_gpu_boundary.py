"""The drop-in boundary SURVEY.md §8b states, exercised end to end: the reference's OWN wrapper files
(models/vqgan_model.py, models/sample_model.py -- staged unmodified into oracle/_ref by oracle/make_ref.py) are
imported with the B200 mirrors installed as `models.archs.vqgan_arch` / `transformer_arch` / `unet_arch` /
`fcn_arch`, constructed by their own `__init__` on 'cuda', and their unmodified methods are run:

  * VQImageSegmTextureModel.forward_step (vqgan_model.py:548)            vs the oracle, pixels / indices
  * VQImageSegmTextureModel.training_step + optimize_parameters (:444-488, :329-344: loss.backward(), two
    torch.optim.Adam, calculate_adaptive_weight's autograd.grad, DiffAugment, hinge_d_loss) on the mirrors as autograd
    nodes                                                                   vs the native VQGANTrainer and the oracle
  * BaseSampleModel.sample_fn (sample_model.py:256)                        vs the same code on the reference archs
"""
import contextlib
import io
import types

import pytest
import torch

import golden_recipes as R

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    got, ref = got.double(), ref.double()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def _loader():
    from oracle import ref_loader as RL
    if not RL.available():
        pytest.skip("reference sources not staged (run oracle/make_ref.py in the build container)")
    return RL


def _tiny_opt():
    cfg = R.TINY_VQGAN_TRAIN
    e = cfg["enc"]
    return dict(embed_dim=cfg["embed_dim"], n_embed=cfg["n_embed"], double_z=False, z_channels=e["z_channels"],
                resolution=e["resolution"], in_channels=3, out_ch=3, ch=e["ch"], ch_mult=e["ch_mult"],
                num_res_blocks=e["num_res_blocks"], attn_resolutions=e["attn_resolutions"], dropout=0.0,
                n_channels=3, ndf=cfg["ndf"], disc_layers=cfg["disc_layers"], perceptual_weight=1.0,
                disc_start_step=cfg["disc_start_step"], disc_weight_max=1.0, diff_aug=True, lr=1e-4)


def _fill(w):
    """the fixture's seeded weights, loaded through the reference wrapper's own attributes"""
    for name, seed in (("encoder", 101), ("decoder", 102), ("quant_conv", 103), ("post_quant_conv", 104), ("disc", 105)):
        mod = getattr(w, name)
        mod.load_state_dict({k: v.cuda() for k, v in R.fill_state_dict(R.spec_of(mod), seed).items()}, strict=True)
    cb = R.codebooks(106, 18, R.TINY_VQGAN_TRAIN["n_embed"], R.TINY_VQGAN_TRAIN["embed_dim"], "trained")
    with torch.no_grad():
        for k, emb in enumerate(w.quantize.embedding_list):
            emb.weight.copy_(cb[k])


def test_reference_vqgan_wrapper_runs_unmodified_on_the_mirrors(cuda):
    from text2human_b200 import ops, vqgan_arch
    ops.set_precision("fp32")
    RL = _loader()
    opt = _tiny_opt()
    B, H, W = 2, 64, 32
    data = dict(image=R.image(107, B, 3, H, W), texture_mask=R.blocky_mask(108, B, H, W, 8))

    # ---- the reference wrapper, built by its own __init__, on the mirrors
    ns = RL.install("mirror", wrappers=("vqgan_model",))
    assert ns.vqgan_model.Encoder is vqgan_arch.Encoder and ns.vqgan_model.Discriminator is vqgan_arch.Discriminator
    with contextlib.redirect_stdout(io.StringIO()):
        wm = ns.vqgan_model.VQImageSegmTextureModel(opt)
    _fill(wm)
    # ---- the same wrapper code on the reference's own arch classes (stock PyTorch): the oracle
    ns_ref = RL.install("reference", wrappers=("vqgan_model",))
    with contextlib.redirect_stdout(io.StringIO()):
        wr = ns_ref.vqgan_model.VQImageSegmTextureModel(opt)
    _fill(wr)

    # inference: forward_step under no_grad, eval mode (vqgan_model.py:493-506)
    for w in (wm, wr):
        for n in ("encoder", "decoder", "quantize", "quant_conv", "post_quant_conv"):
            getattr(w, n).eval()
    x, mask = wm.feed_data(data)
    with torch.no_grad():
        dec_m, diff_m = wm.forward_step(x, mask)
        dec_r, diff_r = wr.forward_step(x, mask)
        _, _, (_, cont_m, _) = wm.encode(x, mask)
        _, _, (_, cont_r, _) = wr.encode(x, mask)
    assert torch.equal(cont_m, cont_r)
    assert _rel(dec_m, dec_r) < 1e-3 and abs(float(diff_m) - float(diff_r)) <= 1e-3 * abs(float(diff_r))

    # training: the reference's optimize_parameters, unmodified, with the same RNG stream for DiffAugment.
    # The step's gradient is discontinuous at the discriminator's LeakyReLU / hinge kinks: pick a DiffAugment seed
    # under which the reference run itself has no pre-activation within 1e-4 of a kink (hooks on the reference disc).
    step = R.TINY_VQGAN_TRAIN["step"]
    for n in ("encoder", "decoder", "quantize", "quant_conv", "post_quant_conv", "disc"):
        getattr(wr, n).train()
    margins = []
    hooks = [mod.register_forward_pre_hook(lambda m_, inp: margins.append(float(inp[0].detach().abs().min())))
             for mod in wr.disc.main if isinstance(mod, torch.nn.LeakyReLU)]
    best = (0.0, None)
    for cand in range(31, 71):
        margins.clear()
        torch.manual_seed(cand)
        loss, d_loss = wr.training_step(data, step)
        best = max(best, (min(margins), cand))
    for h_ in hooks:
        h_.remove()
    margin, seed = best
    print(f"[boundary] DiffAugment seed {seed}: min distance of a discriminator pre-activation to its kink {margin:.2e}")
    assert margin > 3e-5
    before = {k: v.detach().clone() for k, v in wm.decoder.state_dict().items()}
    torch.manual_seed(seed)
    wm.optimize_parameters(data, step)
    torch.manual_seed(seed)
    wr.optimize_parameters(data, step)
    for k in ("nll_loss", "g_loss", "codebook_loss"):
        assert abs(wm.log_dict[k] - wr.log_dict[k]) <= 2e-4 * max(1.0, abs(wr.log_dict[k])), k
    assert abs(float(wm.log_dict["d_weight"]) - float(wr.log_dict["d_weight"])) <= 2e-3 * float(wr.log_dict["d_weight"])
    assert abs(float(wm.log_dict["d_loss"]) - float(wr.log_dict["d_loss"])) <= 2e-4
    worst, gmax = 0.0, 0.0
    pairs = []
    for name in ("encoder", "decoder", "quantize", "quant_conv", "post_quant_conv"):
        pm, pr = dict(getattr(wm, name).named_parameters()), dict(getattr(wr, name).named_parameters())
        assert pm.keys() == pr.keys()
        for k in pm:
            gm, gr = pm[k].grad, pr[k].grad
            if gr is None or float(gr.abs().max()) == 0.0:
                continue
            assert gm is not None, (name, k)
            pairs.append((f"{name}.{k}", gm, gr))
            gmax = max(gmax, float(gr.abs().max()))
    for key, gm, gr in pairs:
        # gradients that vanish in exact arithmetic (a conv bias in front of a per-channel GroupNorm) are rounding
        # noise in the reference as well: held to an absolute floor
        if float(gr.abs().max()) < 1e-6 * gmax:
            assert float((gm - gr).abs().max()) < 1e-5 * gmax, key
            continue
        e = _rel(gm, gr)
        assert e < 1e-3, (key, e)
        worst = max(worst, e)
    print(f"[boundary] reference training_step on the mirrors: worst generator gradient rel err {worst:.2e}")
    assert worst < 1e-3
    dworst = 0.0
    for (k, pm_), (_, pr_) in zip(wm.disc.named_parameters(), wr.disc.named_parameters()):
        dworst = max(dworst, _rel(pm_.grad, pr_.grad))
    print(f"[boundary] discriminator gradient rel err {dworst:.2e}")
    assert dworst < 1e-3
    # both Adam steps happened on the mirrors' own parameters
    after = wm.decoder.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before)
    assert _rel(wm.decoder.conv_out.weight, wr.decoder.conv_out.weight) < 1e-3


def test_reference_sample_fn_runs_unmodified_on_the_mirrors(cuda):
    """BaseSampleModel.sample_fn (sample_model.py:256-328), unbound onto a stand-in that carries the attributes it
    reads (its __init__ only loads checkpoints from disk), with the mirror transformer vs the reference transformer:
    same seed -> the same reveal schedule exactly, and the same tokens except where two candidates tie within
    float rounding of the logits."""
    from text2human_b200 import ops
    ops.set_precision("fp32")
    RL = _loader()
    cfg = R.SAMPLE_TRANSFORMER
    sd = R.fill_state_dict(R.spec_of(__import__("text2human_b200.transformer_arch", fromlist=["x"]).TransformerMultiHead(**cfg)), 81)
    segm, mask = R.sample_inputs(82, R.SAMPLE_BATCH)
    outs = []
    for archs in ("mirror", "reference"):
        ns = RL.install(archs, wrappers=("sample_model",))
        net = ns.transformer_arch.TransformerMultiHead(**cfg)
        net.load_state_dict(sd, strict=True)
        net = net.cuda()
        fake = types.SimpleNamespace(batch_size=R.SAMPLE_BATCH, shape=(32, 16), device=torch.device("cuda"),
                                     mask_id=cfg["codebook_size"], texture_mask=mask.cuda(), segm_tokens=segm.cuda(),
                                     sampler_fn=net)
        torch.manual_seed(83)
        with torch.no_grad():
            outs.append(torch.stack(ns.sample_model.BaseSampleModel.sample_fn(fake, temp=1.0,
                                                                                sample_steps=R.SAMPLE_STEPS)))
    got, want = outs
    assert torch.equal(got >= 0, want >= 0)
    agree = float((got == want).float().mean())
    print(f"[boundary] reference sample_fn on the mirror transformer: {100 * agree:.3f} % of tokens identical")
    assert agree > 0.995
