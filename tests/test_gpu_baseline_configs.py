"""Parity at BASELINE.json's own configurations and sizes (configs 2, 3, 4), CUDA path vs the torch fp32 oracle on the
same device.  No conditional assertions: the encoder latent must be within 1e-3; every end-to-end index that differs
from the oracle's must be a near-tie that the latent error explains (the oracle's own distance margin between the
two candidates is below the bound the latent difference implies); pixels are compared teacher-forced for the whole
batch AND end-to-end on every image whose indices all agree.  Flip counts are printed.

Reference sequences: models/vqgan_model.py:532-551 (config 2), models/hierarchy_vqgan_model.py:215-239 (config 3),
models/archs/transformer_arch.py:249-273 (config 4)."""
import contextlib
import io

import pytest
import torch

import golden_recipes as R

pytestmark = pytest.mark.gpu

TOL = 1e-3

VQVAE_TOP = dict(embed_dim=256, n_embed=1024, double_z=False, z_channels=256, resolution=512, in_channels=3,
                 out_ch=3, ch=128, ch_mult=[1, 1, 2, 2, 4], num_res_blocks=2, attn_resolutions=[32], dropout=0.0)
HIER_OPT = dict(embed_dim=256, n_embed=1024, codebook_spatial_size=2, bot_n_embed=512, bot_double_z=False,
                bot_z_channels=256, bot_resolution=512, bot_in_channels=3, bot_out_ch=3, bot_ch=128,
                bot_ch_mult=[1, 1, 2, 4], bot_num_res_blocks=2, bot_attn_resolutions=[64], bot_dropout=0.0,
                top_double_z=False, top_z_channels=256, top_resolution=512, top_in_channels=3, top_out_ch=3,
                top_ch=128, top_ch_mult=[1, 1, 2, 2, 4], top_num_res_blocks=2, top_attn_resolutions=[32],
                top_dropout=0.0)
SAMPLER_OPT = dict(codebook_size=18432, segm_codebook_size=1024, texture_codebook_size=18, bert_n_emb=512,
                   bert_n_layers=24, bert_n_head=8, block_size=512, latent_shape=[32, 16], embd_pdrop=0.0,
                   resid_pdrop=0.0, attn_pdrop=0.0, num_head=18, sample_steps=256)


def _rel(got, ref):
    got, ref = got.double(), ref.double()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def _rows(z_nchw, ps):
    """[B,C,H,W] -> [B*hp*wp, C*ps*ps] in the reference's row order (F.unfold order for ps=2)"""
    b, c, h, w = z_nchw.shape
    if ps == 1:
        return z_nchw.permute(0, 2, 3, 1).reshape(-1, c)
    return torch.nn.functional.unfold(z_nchw, (ps, ps), stride=ps).permute(0, 2, 1).reshape(-1, c * ps * ps)


def check_flips(tag, z_got, z_ref, codebooks, cont_got, cont_ref, n_e_stride, ps=1):
    """every index that differs from the oracle's must be a near-tie explained by the latent difference.
    Returns the per-row agreement mask."""
    cg, cr = cont_got.reshape(-1), cont_ref.reshape(-1)
    assert bool(((cg < 0) == (cr < 0)).all()), f"{tag}: selected-row sets differ"
    diff = (cg != cr).nonzero().flatten()
    n = cg.numel()
    print(f"[{tag}] {diff.numel()} of {n} end-to-end indices differ from the oracle's "
          f"({100.0 * diff.numel() / n:.3f} %)")
    assert diff.numel() <= 0.02 * n, f"{tag}: {diff.numel()} of {n} indices differ"
    if diff.numel():
        rg, rr = _rows(z_got, ps).double()[diff], _rows(z_ref, ps).double()[diff]
        flat = codebooks.reshape(-1, codebooks.shape[-1]).double()
        book = cr[diff] // n_e_stride
        assert torch.equal(book, cg[diff] // n_e_stride), f"{tag}: a flip crossed codebooks"
        eg = flat[book * codebooks.shape[1] + cg[diff] % n_e_stride]
        er = flat[book * codebooks.shape[1] + cr[diff] % n_e_stride]
        # the oracle's own margin between the two candidates, evaluated on ITS latent in fp64
        margin = ((rr - eg) ** 2).sum(1) - ((rr - er) ** 2).sum(1)
        # what a latent difference dz can move that margin by: 2*|dz . (e_ref - e_got)|, plus the fp32 rounding of the
        # oracle's own distance expression (|z|^2 + |e|^2 - 2 z.e evaluated in fp32: a few ulp of its terms)
        dz = (rg - rr)
        bound = 2.0 * (dz * (er - eg)).sum(1).abs() + 2.0 * dz.norm(dim=1) * (er - eg).norm(dim=1) * 0.05
        ulp = 8 * 1.2e-7 * ((rr ** 2).sum(1) + (er ** 2).sum(1))
        bad = margin.abs() > bound + ulp
        assert not bool(bad.any()), (f"{tag}: {int(bad.sum())} flipped indices are NOT near-ties "
                                     f"(margin {margin[bad][:4].tolist()}, bound {bound[bad][:4].tolist()})")
    return cg == cr


def _top_model(cuda, seed, codebook_kind):
    from text2human_b200.pipeline import VQImageSegmTextureModel
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = VQImageSegmTextureModel(VQVAE_TOP)
    cb = R.codebooks(7, 18, 1024, 256, codebook_kind)
    for k, e in enumerate(m.quantize.embedding_list):
        e.weight.data.copy_(cb[k])
    m = m.to(cuda).eval()
    return m, {k: v.detach() for k, v in m.state_dict().items()}, cb.to(cuda)


@pytest.mark.parametrize("batch,mask_kind,codebook_kind", [(16, "blocky", "trained"), (4, "iid", "default")])
def test_config2_vqvae_top_full_size(cuda, batch, mask_kind, codebook_kind):
    """BASELINE config 2: vqvae_top nets, 512x256, the benchmarked batch (16) with blocky masks and a trained-like
    codebook; i.i.d. masks (all 18 codebooks in every image) with the reference's default near-tied codebook init."""
    from oracle import vqgan_ref
    from text2human_b200 import ops
    ops.set_precision("fp32")
    m, sd, cb = _top_model(cuda, 2021, codebook_kind)
    x = R.image(2021, batch, 3, 512, 256).to(cuda)
    mask = (R.blocky_mask(2021, batch, 512, 256, 32) if mask_kind == "blocky" else R.iid_mask(2021, batch, 512, 256))
    mask = mask.to(cuda)
    with torch.no_grad():
        want = vqgan_ref.vq_forward_step(sd, cb, x, mask)
    dec, loss, info = m.forward_step(x, mask, return_info=True)
    z = info["z_nhwc"].permute(0, 3, 1, 2)
    zerr = _rel(z, want["z"])
    print(f"[config2 B={batch} {mask_kind}/{codebook_kind}] encoder latent rel err {zerr:.2e}")
    assert zerr < TOL
    agree = check_flips(f"config2 {mask_kind}/{codebook_kind}", z, want["z"], cb, info["idx_cont"], want["idx_cont"],
                        1024)
    # teacher-forced decode of the oracle's quantized latent: whole batch
    with torch.no_grad():
        ref_dec = vqgan_ref.decoder(sd, vqgan_ref.conv(sd, "post_quant_conv", want["quant"], padding=0), "decoder.")
    perr = _rel(m.decode(want["quant"]), ref_dec)
    print(f"[config2 B={batch}] teacher-forced pixel rel err {perr:.2e}")
    assert perr < TOL
    # end to end on the images whose indices all agree (at least one must)
    img_ok = agree.view(batch, -1).all(1)
    assert int(img_ok.sum()) >= 1
    e2e = _rel(dec[img_ok], want["dec"][img_ok])
    print(f"[config2 B={batch}] end-to-end pixel rel err {e2e:.2e} on {int(img_ok.sum())}/{batch} flip-free images")
    assert e2e < TOL
    if bool(img_ok.all()):
        assert abs(loss.item() - want["loss"].item()) <= 1e-3 * abs(want["loss"].item())
    # quantizer bit-exactness on the kernel's own latent, at full size: against the fixed-order C oracle
    from oracle import vq_oracle
    ids = vq_oracle.nearest_ids(mask.cpu().numpy(), 32, 16)
    wq = vq_oracle.search(info["z_nhwc"].cpu().numpy(), cb.cpu().numpy(), ids, cont_stride=1024)
    assert (info["idx_cont"].cpu().numpy() == wq["idx_cont"]).all(), "indices not bit-exact on the kernel's own latent"


def test_config2_mixed_precision_map(cuda):
    """ops.set_precision("mixed") (tools/precision_map.py: the six 128-channel convs of the decoder's full-resolution
    level single-product, 30 % of the conv FLOPs) at BASELINE config 2's full size: the encoder is untouched, so the
    indices are those of the parity mode bit for bit; decoded pixels stay within the 1e-3 the north star allows."""
    from oracle import vqgan_ref
    from text2human_b200 import ops
    m, sd, cb = _top_model(cuda, 2021, "trained")
    batch = 16
    x = R.image(2021, batch, 3, 512, 256).to(cuda)
    mask = R.blocky_mask(2021, batch, 512, 256, 32).to(cuda)
    with torch.no_grad():
        want = vqgan_ref.vq_forward_step(sd, cb, x, mask)
    ops.set_precision("fp32")
    _, _, info32 = m.forward_step(x, mask, return_info=True)
    ops.set_precision("mixed")
    try:
        dec, loss, info = m.forward_step(x, mask, return_info=True)
    finally:
        ops.set_precision("fp32")
    assert torch.equal(info["idx_cont"], info32["idx_cont"]) and torch.equal(info["z_nhwc"], info32["z_nhwc"])
    assert torch.equal(info["idx_cont"], want["idx_cont"])
    e2e = _rel(dec, want["dec"])
    print(f"[config2 mixed] end-to-end pixel rel err {e2e:.2e} (budget 1e-3)")
    assert e2e < TOL


def test_config3_hierarchy_full_size(cuda):
    """BASELINE config 3: vqvae_top + vqvae_bottom nets at 512x256 (B=2 of the benchmarked 8: the path is
    per-image), forward_step = top_encode + bot_encode + decode with the bottom residual."""
    from oracle import vqgan_ref
    from text2human_b200 import ops
    from text2human_b200.pipeline import HierarchyVQSpatialTextureAwareModel
    ops.set_precision("fp32")
    torch.manual_seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        m = HierarchyVQSpatialTextureAwareModel(HIER_OPT)
    cbt, cbb = R.codebooks(8, 18, 1024, 256, "trained"), R.codebooks(9, 18, 512, 1024, "trained")
    for k in range(18):
        m.top_quantize.embedding_list[k].weight.data.copy_(cbt[k])
        m.bot_quantize.embedding_list[k].weight.data.copy_(cbb[k])
    m = m.to(cuda).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    cbt, cbb = cbt.to(cuda), cbb.to(cuda)
    B = 2
    x = R.image(3, B, 3, 512, 256).to(cuda)
    mask = R.blocky_mask(3, B, 512, 256, 32).to(cuda)
    with torch.no_grad():
        want = vqgan_ref.hierarchy_forward_step(sd, cbt, cbb, x, mask)
    dec, loss, info = m.forward_step(x, mask, return_info=True)
    from text2human_b200.vqgan_arch import conv1x1_nhwc
    zt = conv1x1_nhwc(m.top_encoder.forward_nhwc(x), m.top_quant_conv).permute(0, 3, 1, 2)
    zb = conv1x1_nhwc(m.bot_encoder.forward_nhwc(x), m.bot_quant_conv).permute(0, 3, 1, 2)
    et, eb = _rel(zt, want["z_top"]), _rel(zb, want["z_bot"])
    print(f"[config3] latent rel err top {et:.2e} bottom {eb:.2e}")
    assert et < TOL and eb < TOL
    ok_t = check_flips("config3 top", zt, want["z_top"], cbt, info["top_idx"], want["top_idx"], 1024)
    ok_b = check_flips("config3 bottom", zb, want["z_bot"], cbb, info["bot_idx"], want["bot_idx"], 512, ps=2)
    # teacher-forced: the oracle's quantized latents through our decode chain
    with torch.no_grad():
        qt, _, _, _ = vqgan_ref.quantize_texture(cbt, want["z_top"], mask)
        qb, _, _, _ = vqgan_ref.quantize_texture(cbb, want["z_bot"], mask, ps=2, cont_stride=512)
        qt = vqgan_ref.conv(sd, "top_post_quant_conv", qt, padding=0)
        res = vqgan_ref.decoder_res(sd, vqgan_ref.conv(sd, "bot_post_quant_conv", qb, padding=0), "bot_decoder_res.")
        ref_dec = vqgan_ref.decoder(sd, qt, "decoder.", bot_h=res)
    assert _rel(m.bot_decoder_res(vqgan_ref.conv(sd, "bot_post_quant_conv", qb, padding=0)), res) < TOL
    perr = _rel(m.decode(qt, res), ref_dec)
    print(f"[config3] teacher-forced pixel rel err {perr:.2e}")
    assert perr < TOL
    img_ok = ok_t.view(B, -1).all(1) & ok_b.view(B, -1).all(1)
    print(f"[config3] {int(img_ok.sum())}/{B} images flip-free end to end")
    if bool(img_ok.any()):
        assert _rel(dec[img_ok], want["dec"][img_ok]) < TOL
    if bool(img_ok.all()):
        assert abs(loss.item() - want["loss"].item()) <= 1e-3 * abs(want["loss"].item())


def test_config4_sampler_logits_batch4(cuda):
    """BASELINE config 4 transformer at its batch (4 x 512 tokens): all-heads logits vs the oracle, and the own-head
    path the sampling loop uses, at a mid-trajectory token mix"""
    from oracle import transformer_ref
    from text2human_b200 import ops
    from text2human_b200.pipeline import Sampler
    ops.set_precision("fp32")
    torch.manual_seed(11)
    s = Sampler(SAMPLER_OPT)
    with torch.no_grad():
        s.sampler_fn.pos_emb.normal_(0, 0.02)
    s = s.to(cuda).eval()
    sd = {k: v.detach() for k, v in s.sampler_fn.state_dict().items()}
    B, T = 4, 512
    g = torch.Generator().manual_seed(5)
    mask = R.blocky_mask(9, B, 512, 256, 64)
    tex = torch.nn.functional.interpolate(mask, (32, 16), mode="nearest").view(B, T).long()
    idx = torch.where(torch.rand(B, T, generator=g) < 0.5, torch.full((B, T), 18432),
                      torch.randint(0, 1024, (B, T), generator=g) + 1024 * tex).to(cuda)
    segm = torch.randint(0, 1024, (B, T), generator=g).to(cuda)
    tex = tex.to(cuda)
    with torch.no_grad():
        want = torch.stack(transformer_ref.transformer_logits(sd, idx, segm, tex, n_head=8), 2)
    got = s.sampler_fn.forward_logits(idx, segm, tex)
    err = _rel(got, want)
    print(f"[config4 B=4] logits rel err {err:.2e}")
    assert got.shape == (B, T, 18, 1024) and err < TOL
    dest, rows = s.sampler_fn.group_by_texture(tex, 18)
    hf = torch.zeros((ops.get_terms(), 18 * rows, 512), dtype=torch.float16, device=cuda)
    own = s.sampler_fn.forward_own_logits(idx, segm, tex, dest, hf)
    want_own = got.gather(2, tex.view(B, T, 1, 1).expand(B, T, 1, 1024)).view(B * T, 1024)
    assert torch.equal(own, want_own)


def test_config4_sample_fn_against_pinned_oracle(cuda):
    """BASELINE config 4 sampling loop (B=4, real 24x512 transformer, 18 heads) against oracle/transformer_ref.sample_fn
    (pinned token-for-token to the real BaseSampleModel.sample_fn by tests/test_sample_oracle.py):
      * identical reveal schedule under shared reveal uniforms (the set revealed at every step),
      * per-step own-head logits at the traced states equal the oracle's logits within 1e-3,
      * every token comes from its position's own texture codebook; all positions end unmasked,
      * CUDA-graph replay == eager launches, and a weight update invalidates the captured graph."""
    from oracle import transformer_ref
    from text2human_b200 import ops
    from text2human_b200.pipeline import Sampler
    ops.set_precision("fp32")
    torch.manual_seed(11)
    s = Sampler(SAMPLER_OPT)
    with torch.no_grad():
        s.sampler_fn.pos_emb.normal_(0, 0.02)
    s = s.to(cuda).eval()
    sd = {k: v.detach() for k, v in s.sampler_fn.state_dict().items()}
    B, T, steps = 4, 512, 8
    g = torch.Generator().manual_seed(6)
    segm = torch.randint(0, 1024, (B, T), generator=g).to(cuda)
    mask = R.blocky_mask(9, B, 512, 256, 64).to(cuda)
    u = torch.rand((steps, B, T), generator=g).to(cuda)
    trace = []
    out, x_t = s.sample_fn(segm, mask, sample_steps=steps, reveal_u=u, seed=77, trace=trace)
    # oracle run with the same reveal uniforms (its own categorical draws): the reveal sets must coincide
    otrace = []
    with torch.no_grad():
        transformer_ref.sample_fn(lambda x, sg, tx: transformer_ref.transformer_logits(sd, x, sg, tx, n_head=8), segm,
                                  mask, [32, 16], 18432, steps, trace=otrace, reveal_u=u)
    assert len(trace) == len(otrace) == steps
    for i, ((xb, ch), (oxb, och)) in enumerate(zip(trace, otrace)):
        assert torch.equal(ch, och), f"step {i}: revealed set differs from the reference schedule"
    tex = torch.nn.functional.interpolate(mask, (32, 16), mode="nearest").view(B, T).long()
    assert bool((x_t != 18432).all()) and bool(((x_t // 1024) == tex).all())
    for k in range(18):
        assert bool(((out[k] >= 0) == (tex == k)).all()) and int(out[k].max()) < 1024
    # teacher-forced: the logits our loop sampled from at its own traced states vs the oracle's logits there
    for i in (0, steps // 2, steps - 1):
        xb = trace[i][0]
        with torch.no_grad():
            want = torch.stack(transformer_ref.transformer_logits(sd, xb, segm, tex, n_head=8), 2)
        got = s.sampler_fn.forward_logits(xb, segm, tex)
        assert _rel(got, want) < TOL, i
    # graph replay vs eager: same random inputs -> same tokens
    out2, x_t2 = s.sample_fn(segm, mask, sample_steps=steps, reveal_u=u, seed=77, use_graph=False)
    assert torch.equal(x_t, x_t2)
    # a weight change must not replay stale packed weights
    with torch.no_grad():
        for p in s.sampler_fn.parameters():
            p.mul_(1.01)
    _, xa = s.sample_fn(segm, mask, sample_steps=steps, reveal_u=u, seed=77)
    _, xb2 = s.sample_fn(segm, mask, sample_steps=steps, reveal_u=u, seed=77, use_graph=False)
    assert torch.equal(xa, xb2)
    assert len(s._graphs) <= s.MAX_GRAPHS
