"""GPU parity of the module mirrors (through libt2h) against (a) the golden outputs of the real
reference modules and (b) the torch fp32 oracle on the same device, in both precision modes."""
import os

import numpy as np
import pytest
import torch

import golden_recipes as R

pytestmark = pytest.mark.gpu

# tolerance north_star states: decoded pixels / logits within 1e-3 relative (max-norm) of fp32.
TOL_EXACT = 1e-3     # "fp32" mode (3-product fp16 split): expected ~1e-5
TOL_FAST = 3e-2      # "fp16" mode (TF32-like single product): documented, not the parity mode


def _rel(got, ref):
    got = torch.as_tensor(got).double().cpu()
    ref = torch.as_tensor(ref).double().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def _load(mod, seed, cuda):
    sd = R.fill_state_dict(R.spec_of(mod), seed)
    mod.load_state_dict(sd, strict=True)
    return mod.to(cuda).eval(), {k: v.to(cuda) for k, v in sd.items()}


@pytest.fixture(params=["fp32", "fp16"])
def mode(request):
    from text2human_b200 import ops
    ops.set_precision(request.param)
    yield request.param
    ops.set_precision("fp32")


def _tol(mode):
    return TOL_EXACT if mode == "fp32" else TOL_FAST


def test_blocks_match_reference_golden(cuda, golden_dir, mode):
    from text2human_b200 import vqgan_arch as A
    g = np.load(os.path.join(golden_dir, "vqgan_modules.npz"))
    xb = R.latent(49, (2, 64, 16, 8)).to(cuda)
    rb, _ = _load(A.ResnetBlock(in_channels=64, out_channels=128, temb_channels=0, dropout=0.0), 48, cuda)
    assert _rel(rb(xb), g["resblock"]) < _tol(mode)
    ab, _ = _load(A.AttnBlock(64), 50, cuda)
    assert _rel(ab(xb), g["attnblock"]) < _tol(mode)
    up, _ = _load(A.Upsample(64, True), 51, cuda)
    assert _rel(up(xb), g["upsample"]) < _tol(mode)
    dn, _ = _load(A.Downsample(64, True), 52, cuda)
    assert _rel(dn(xb), g["downsample"]) < _tol(mode)


def test_encoder_decoder_match_reference_golden(cuda, golden_dir, mode):
    from text2human_b200 import vqgan_arch as A
    g = np.load(os.path.join(golden_dir, "vqgan_modules.npz"))
    enc, _ = _load(A.Encoder(**R.TINY_ENC), 41, cuda)
    assert _rel(enc(R.image(42, 2, 3, 32, 16).to(cuda)), g["enc_z"]) < _tol(mode)
    dec, _ = _load(A.Decoder(**R.TINY_DEC), 43, cuda)
    z = R.latent(44, (2, 32, 4, 2)).to(cuda)
    bot_h = R.latent(45, (2, 64, 8, 4), name="bot_h").to(cuda)
    keep = bot_h.clone()
    assert _rel(dec(z), g["dec_plain"]) < _tol(mode)
    assert _rel(dec(z, bot_h=bot_h), g["dec_both"]) < _tol(mode)
    assert torch.equal(bot_h, keep), "caller's bot_h must not be modified"
    assert tuple(dec.last_z_shape) == (2, 32, 4, 2)
    assert _rel(dec.get_feature_top(z), g["dec_feature_top"]) < _tol(mode)
    dres, _ = _load(A.DecoderRes(**R.TINY_DECRES), 46, cuda)
    assert _rel(dres(R.latent(47, (2, 32, 8, 4)).to(cuda)), g["decres"]) < _tol(mode)


def test_quantizer_modules_match_reference_golden(cuda, golden_dir):
    from text2human_b200 import vqgan_arch as A
    g = np.load(os.path.join(golden_dir, "quantizers.npz"))
    for kind in ("default", "trained"):
        zs = 1.0 if kind == "trained" else 0.02
        q = A.VectorQuantizerTexture(128, 256, beta=0.25)
        for k, e in enumerate(q.embedding_list):
            e.weight.data.copy_(R.codebooks(11, 18, 128, 256, kind)[k])
        q = q.to(cuda)
        z = R.latent(12, (2, 256, 32, 16), zs).to(cuda)
        for mname, mask in (("blocky", R.blocky_mask(13, 2, 512, 256, 64, extra_ids=(20,))),
                            ("iid", R.iid_mask(14, 2, 512, 256))):
            zq, loss, (perp, cont, lst) = q(z, mask.to(cuda))
            tag = f"top_{kind}_{mname}"
            assert perp is None and cont.dtype == torch.int64 and cont.shape == (2, 32, 16)
            assert np.array_equal(cont.cpu().numpy(), g[tag + "_cont"])
            assert len(lst) == 18 and np.array_equal(torch.stack(lst).cpu().numpy(), g[tag + "_list"])
            assert abs(loss.item() - float(g[tag + "_loss"])) <= 1e-5 * float(g[tag + "_loss"])
            assert abs(zq.double().abs().sum().item() - float(g[tag + "_zq_abs"])) <= 1e-6 * float(g[tag + "_zq_abs"])
        ent = q.get_codebook_entry(lst, mask.to(cuda), (2, 32, 16, 256))
        assert abs(ent.double().abs().sum().item() - float(g[f"top_{kind}_entry_abs"])) \
            <= 1e-6 * float(g[f"top_{kind}_entry_abs"])
        qb = A.VectorQuantizerSpatialTextureAware(64, 32, beta=0.25, spatial_size=2)
        for k, e in enumerate(qb.embedding_list):
            e.weight.data.copy_(R.codebooks(21, 18, 64, 128, kind)[k])
        qb = qb.to(cuda)
        zb = R.latent(22, (2, 32, 32, 16), zs).to(cuda)
        maskb = R.blocky_mask(23, 2, 256, 128, 32).to(cuda)
        zqb, lossb, (_, contb, lstb) = qb(zb, maskb)
        assert contb.dim() == 1 and np.array_equal(contb.cpu().numpy(), g[f"bot_{kind}_cont"])
        assert np.array_equal(torch.stack(lstb).cpu().numpy(), g[f"bot_{kind}_list"])
        assert np.allclose(zqb.cpu().numpy(), g[f"bot_{kind}_zq"], rtol=0, atol=1e-7)
        entb = qb.get_codebook_entry(lstb, maskb, (2, 16, 8, 32))
        assert np.array_equal(entb.cpu().numpy(), g[f"bot_{kind}_entry"])
        qs = A.VectorQuantizer(128, 32, beta=0.25, sane_index_shape=True)
        qs.embedding.weight.data.copy_(R.codebooks(31, 1, 128, 32, kind)[0])
        qs = qs.to(cuda)
        _, _, (_, _, idxs) = qs(R.latent(32, (2, 32, 32, 16), zs).to(cuda))
        assert idxs.shape == (2, 32, 16) and np.array_equal(idxs.cpu().numpy(), g[f"plain_{kind}_idx"])


def test_transformer_matches_reference_golden(cuda, golden_dir, mode):
    from text2human_b200 import transformer_arch as T
    g = np.load(os.path.join(golden_dir, "transformer.npz"))
    tf, _ = _load(T.TransformerMultiHead(**R.TINY_TRANSFORMER), 61, cuda)
    gen = R._gen(62, "tokens")
    idx = torch.randint(0, 18 * 16 + 1, (2, 32), generator=gen).to(cuda)
    segm = torch.randint(0, 32, (2, 32), generator=gen).to(cuda)
    tex = torch.randint(0, 18, (2, 32), generator=gen).to(cuda)
    out = tf(idx, segm, tex)
    assert len(out) == 18 and out[0].shape == (2, 32, 16)
    assert _rel(torch.stack(out), g["logits"]) < _tol(mode)


def test_vq_pipeline_config1_matches_oracle(cuda):
    """BASELINE config 1: vqvae_top nets, 1x3x256x128, encode -> quantize -> decode, vs the torch fp32
    oracle on the same device.  Indices are compared through the oracle's own z (teacher-forced) and
    end to end with a trained-like codebook."""
    from oracle import vqgan_ref
    from text2human_b200 import ops
    from text2human_b200.pipeline import VQImageSegmTextureModel
    ops.set_precision("fp32")
    opt = dict(embed_dim=256, n_embed=1024, double_z=False, z_channels=256, resolution=512, in_channels=3,
               out_ch=3, ch=128, ch_mult=[1, 1, 2, 2, 4], num_res_blocks=2, attn_resolutions=[32], dropout=0.0)
    torch.manual_seed(2021)
    m = VQImageSegmTextureModel(opt)
    cb = R.codebooks(7, 18, 1024, 256, "trained")
    for k, e in enumerate(m.quantize.embedding_list):
        e.weight.data.copy_(cb[k])
    m = m.to(cuda).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    x = R.image(2021, 1, 3, 256, 128).to(cuda)
    mask = R.blocky_mask(2021, 1, 256, 128, 32).to(cuda)
    with torch.no_grad():
        want = vqgan_ref.vq_forward_step(sd, cb.to(cuda), x, mask)
    dec, loss, info = m.forward_step(x, mask, return_info=True)
    z = info["z_nhwc"].permute(0, 3, 1, 2)
    assert _rel(z, want["z"]) < TOL_EXACT
    from test_gpu_baseline_configs import check_flips   # every differing index must be a near-tie the z error explains
    ok = check_flips("config1", z, want["z"], cb.to(cuda), info["idx_cont"], want["idx_cont"], 1024)
    assert bool(ok.any())
    if bool(ok.all()):   # single image: the end-to-end comparison needs a flip-free image (teacher-forced check below
        assert _rel(dec, want["dec"]) < TOL_EXACT    # covers the decoder unconditionally); flip count is printed
    # teacher-forced decode: identical quantized input -> pixels within 1e-3
    dec_tf = m.decode(want["quant"])
    with torch.no_grad():
        ref_dec = vqgan_ref.decoder(sd, vqgan_ref.conv(sd, "post_quant_conv", want["quant"], padding=0), "decoder.")
    assert _rel(dec_tf, ref_dec) < TOL_EXACT
    assert abs(loss.item() - want["loss"].item()) <= 1e-3 * abs(want["loss"].item())


HIER_OPT = dict(embed_dim=256, n_embed=1024, codebook_spatial_size=2, bot_n_embed=512, bot_double_z=False,
                bot_z_channels=256, bot_resolution=512, bot_in_channels=3, bot_out_ch=3, bot_ch=128,
                bot_ch_mult=[1, 1, 2, 4], bot_num_res_blocks=2, bot_attn_resolutions=[64], bot_dropout=0.0,
                top_double_z=False, top_z_channels=256, top_resolution=512, top_in_channels=3, top_out_ch=3,
                top_ch=128, top_ch_mult=[1, 1, 2, 2, 4], top_num_res_blocks=2, top_attn_resolutions=[32],
                top_dropout=0.0)


def test_hierarchy_forward_step_matches_oracle(cuda):
    """BASELINE config 3 nets (vqvae_top + vqvae_bottom) on a 128x64 crop: top/bottom latents, both index
    maps and the residual decode against the torch fp32 oracle (hierarchy_vqgan_model.py:215-239)."""
    from oracle import vqgan_ref
    from text2human_b200 import ops
    from text2human_b200.pipeline import HierarchyVQSpatialTextureAwareModel
    ops.set_precision("fp32")
    torch.manual_seed(7)
    m = HierarchyVQSpatialTextureAwareModel(HIER_OPT)
    cbt = R.codebooks(8, 18, 1024, 256, "trained")
    cbb = R.codebooks(9, 18, 512, 1024, "trained")
    for k in range(18):
        m.top_quantize.embedding_list[k].weight.data.copy_(cbt[k])
        m.bot_quantize.embedding_list[k].weight.data.copy_(cbb[k])
    m = m.to(cuda).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    x = R.image(3, 2, 3, 128, 64).to(cuda)
    mask = R.blocky_mask(3, 2, 128, 64, 16).to(cuda)
    with torch.no_grad():
        want = vqgan_ref.hierarchy_forward_step(sd, cbt.to(cuda), cbb.to(cuda), x, mask)
    dec, loss, info = m.forward_step(x, mask, return_info=True)
    from test_gpu_baseline_configs import check_flips
    from text2human_b200.vqgan_arch import conv1x1_nhwc
    zt = conv1x1_nhwc(m.top_encoder.forward_nhwc(x), m.top_quant_conv).permute(0, 3, 1, 2)
    zb = conv1x1_nhwc(m.bot_encoder.forward_nhwc(x), m.bot_quant_conv).permute(0, 3, 1, 2)
    assert _rel(zt, want["z_top"]) < TOL_EXACT and _rel(zb, want["z_bot"]) < TOL_EXACT
    ok_t = check_flips("hier-crop top", zt, want["z_top"], cbt.to(cuda), info["top_idx"], want["top_idx"], 1024)
    ok_b = check_flips("hier-crop bottom", zb, want["z_bot"], cbb.to(cuda), info["bot_idx"], want["bot_idx"], 512, ps=2)
    img_ok = ok_t.view(2, -1).all(1) & ok_b.view(2, -1).all(1)
    if bool(img_ok.any()):
        assert _rel(dec[img_ok], want["dec"][img_ok]) < TOL_EXACT
    if bool(img_ok.all()):
        assert abs(loss.item() - want["loss"].item()) <= 1e-3 * abs(want["loss"].item())
    # module-level API (NCHW in/out) of the pieces the reference wrapper calls, teacher-forced
    with torch.no_grad():
        zt = vqgan_ref.conv(sd, "top_quant_conv", vqgan_ref.encoder(sd, x, "top_encoder."), padding=0)
        zb = vqgan_ref.conv(sd, "bot_quant_conv", vqgan_ref.encoder(sd, x, "bot_encoder."), padding=0)
        qt, _, _, _ = vqgan_ref.quantize_texture(cbt.to(cuda), zt, mask)
        qb, _, _, _ = vqgan_ref.quantize_texture(cbb.to(cuda), zb, mask, ps=2, cont_stride=512)
        qt = vqgan_ref.conv(sd, "top_post_quant_conv", qt, padding=0)
        res = vqgan_ref.decoder_res(sd, vqgan_ref.conv(sd, "bot_post_quant_conv", qb, padding=0), "bot_decoder_res.")
        ref_dec = vqgan_ref.decoder(sd, qt, "decoder.", bot_h=res)
    assert _rel(m.bot_encoder(x), vqgan_ref.encoder(sd, x, "bot_encoder.")) < TOL_EXACT
    assert _rel(m.bot_decoder_res(vqgan_ref.conv(sd, "bot_post_quant_conv", qb, padding=0)), res) < TOL_EXACT
    assert _rel(m.decode(qt, res), ref_dec) < TOL_EXACT


SAMPLER_OPT = dict(codebook_size=18432, segm_codebook_size=1024, texture_codebook_size=18, bert_n_emb=512,
                   bert_n_layers=24, bert_n_head=8, block_size=512, latent_shape=[32, 16], embd_pdrop=0.0,
                   resid_pdrop=0.0, attn_pdrop=0.0, num_head=18, sample_steps=256)


def test_sampler_transformer_logits_and_sampling_loop(cuda, mode):
    """BASELINE config 4 transformer (24 x 512, 18 heads): logits vs the torch fp32 oracle on a masked /
    unmasked token mix (teacher-forced parity), then the sampling loop's structural contract."""
    from oracle import transformer_ref
    from text2human_b200.pipeline import Sampler
    torch.manual_seed(11)
    s = Sampler(SAMPLER_OPT)
    with torch.no_grad():  # the reference leaves pos_emb at zero; make it matter for the test
        s.sampler_fn.pos_emb.normal_(0, 0.02)
    s = s.to(cuda).eval()
    sd = {k: v.detach() for k, v in s.sampler_fn.state_dict().items()}
    B, T = 2, 512
    g = torch.Generator().manual_seed(5)
    tex = torch.randint(0, 18, (B, T), generator=g)
    idx = torch.where(torch.rand(B, T, generator=g) < 0.5, torch.full((B, T), 18432),
                      torch.randint(0, 1024, (B, T), generator=g) + 1024 * tex).to(cuda)
    segm = torch.randint(0, 1024, (B, T), generator=g).to(cuda)
    tex = tex.to(cuda)
    with torch.no_grad():
        want = torch.stack(transformer_ref.transformer_logits(sd, idx, segm, tex, n_head=8), 2)
    got = s.sampler_fn.forward_logits(idx, segm, tex)
    assert got.shape == (B, T, 18, 1024)
    assert _rel(got, want) < _tol(mode)
    lst = s.sampler_fn(idx, segm, tex)   # deterministic by default: a second forward is bit-identical
    assert len(lst) == 18 and lst[3].shape == (B, T, 1024) and torch.equal(lst[3], got[:, :, 3])
    from text2human_b200 import ops
    old = ops.set_split_k(inference=True)   # opt-in small-batch path: k-slices reduce-added into the stream
    try:
        got_sk = s.sampler_fn.forward_logits(idx, segm, tex)
    finally:
        ops.set_split_k(**old)
    # (in fp16 mode an fp32-ulp change of the stream can flip an fp16 rounding downstream)
    assert _rel(got_sk, want) < _tol(mode) and _rel(got_sk, got) < (1e-4 if mode == "fp32" else _tol(mode))
    # own-head path of the sampling loop: positions grouped by texture, one batched GEMM against each group's
    # own head -- the same numbers as the all-heads logits gathered at the own head, bit for bit
    for tex_case in (tex, torch.full_like(tex, 5), torch.where(tex < 9, tex, torch.full_like(tex, 20))):
        tc = tex_case.clamp(0, 17)
        dest, rows = s.sampler_fn.group_by_texture(tc, 18)
        assert rows % 128 == 0 and int(dest.max()) < 18 * rows and dest.unique().numel() == B * T
        hf = torch.zeros((ops.get_terms(), 18 * rows, 512), dtype=torch.float16, device=cuda)
        own = s.sampler_fn.forward_own_logits(idx, segm, tc, dest, hf)
        full = s.sampler_fn.forward_logits(idx, segm, tc)
        want_own = full.gather(2, tc.view(B, T, 1, 1).expand(B, T, 1, 1024)).view(B * T, 1024)
        assert torch.equal(own, want_own)
    if mode == "fp32":
        mask = R.blocky_mask(4, B, 512, 256, 64).to(cuda)
        gen = torch.Generator(device=cuda).manual_seed(2021)
        out, x_t = s.sample_fn(segm, mask, sample_steps=6, generator=gen)
        texm = torch.nn.functional.interpolate(mask, (32, 16), mode="nearest").view(B, -1).long()
        assert (x_t != 18432).all(), "every position must be unmasked after the last step"
        assert ((x_t // 1024) == texm).all(), "tokens must come from the position's own texture codebook"
        for k in range(18):
            assert ((out[k] >= 0) == (texm == k)).all() and out[k].max() < 1024


SEGM_OPT = dict(segm_double_z=False, segm_z_channels=32, segm_resolution=512, segm_in_channels=24, segm_out_ch=24,
                segm_ch=64, segm_ch_mult=[1, 1, 2, 2, 4], segm_num_res_blocks=1, segm_attn_resolutions=[16],
                segm_dropout=0.0, segm_num_segm_classes=24, segm_n_embed=1024, segm_embed_dim=32)


def test_segm_tokenizer_matches_oracle(cuda):
    """get_quantized_segm (sample_model.py:330-340): one-hot(24) -> segm Encoder (ch 64, 2 channels per
    GroupNorm group) -> 1x1 -> VectorQuantizer(1024, 32) on a 128x64 parsing map."""
    from oracle import vqgan_ref
    from text2human_b200 import ops
    from text2human_b200.pipeline import SegmTokenizer
    ops.set_precision("fp32")
    torch.manual_seed(21)
    m = SegmTokenizer(SEGM_OPT)
    m.segm_quantizer.embedding.weight.data.copy_(R.codebooks(5, 1, 1024, 32, "trained")[0])
    m = m.to(cuda).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    segm = R.blocky_mask(6, 2, 128, 64, 8, n_ids=24).to(cuda)
    tokens = m.get_quantized_segm(segm)
    assert tokens.shape == (2, 8, 4) and tokens.dtype == torch.int64
    with torch.no_grad():
        oh = torch.nn.functional.one_hot(segm.squeeze(1).long(), 24).permute(0, 3, 1, 2).float()
        z = vqgan_ref.conv(sd, "segm_quant_conv", vqgan_ref.encoder(sd, oh, "segm_encoder."), padding=0)
        _, _, want = vqgan_ref.quantize_plain(sd["segm_quantizer.embedding.weight"], z)
    a = ops.onehot_to_planes(segm, 24)
    assert torch.equal(a.float().sum(0).permute(0, 3, 1, 2), oh)
    zz = m.segm_encoder.forward_planes(a)
    assert _rel(zz.permute(0, 3, 1, 2), vqgan_ref.encoder(sd, oh, "segm_encoder.")) < TOL_EXACT
    assert (tokens == want).float().mean().item() >= 0.95


def test_decode_from_indices_matches_oracle(cuda):
    """tokens -> image (the decode half of sample_and_refine, sample_model.py:225-243), batched"""
    from oracle import vqgan_ref
    from text2human_b200 import ops
    from text2human_b200.pipeline import HierarchyVQSpatialTextureAwareModel
    ops.set_precision("fp32")
    torch.manual_seed(9)
    m = HierarchyVQSpatialTextureAwareModel(HIER_OPT).to(cuda).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    cbt = torch.stack([e.weight.detach() for e in m.top_quantize.embedding_list])
    cbb = torch.stack([e.weight.detach() for e in m.bot_quantize.embedding_list])
    B = 2
    mask = R.blocky_mask(12, B, 512, 256, 64).to(cuda)
    tex = torch.nn.functional.interpolate(mask, (32, 16), mode="nearest")[:, 0].long()
    g = torch.Generator().manual_seed(1)
    top = torch.randint(0, 1024, (B, 32, 16), generator=g).to(cuda)
    bot = torch.randint(0, 512, (B, 32, 16), generator=g).to(cuda)
    top_list = [torch.where(tex == k, top, torch.full_like(top, -1)) for k in range(18)]
    bot_list = [torch.where(tex == k, bot, torch.full_like(bot, -1)) for k in range(18)]
    dec = m.decode_from_indices(top_list, bot_list, mask)
    with torch.no_grad():
        qt = vqgan_ref.codebook_entry_texture(cbt, top_list, mask, (B, 32, 16, 256))
        qt = vqgan_ref.conv(sd, "top_post_quant_conv", qt, padding=0)
        qb = vqgan_ref.codebook_entry_texture(cbb, bot_list, mask, (B, 32, 16, 256), ps=2)
        res = vqgan_ref.decoder_res(sd, vqgan_ref.conv(sd, "bot_post_quant_conv", qb, padding=0), "bot_decoder_res.")
        want = vqgan_ref.decoder(sd, qt, "decoder.", bot_h=res)
    assert dec.shape == (B, 3, 512, 256)
    assert _rel(dec, want) < TOL_EXACT
