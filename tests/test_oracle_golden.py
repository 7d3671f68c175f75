"""Pins the oracle (oracle/) against outputs of the REAL reference modules
(tests/golden/*.npz, produced by oracle/make_golden.py from /root/reference).
CPU only.  Index parity is exact; floating-point parity is ~fp32 round-off since
both sides are fp32 PyTorch on the same recipe."""
import os

import numpy as np
import pytest
import torch

import golden_recipes as R
from oracle import transformer_ref, vq_oracle, vqgan_ref

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


@pytest.fixture(scope="module")
def gq(golden_dir):
    return np.load(os.path.join(golden_dir, "quantizers.npz"))


@pytest.fixture(scope="module")
def gm(golden_dir):
    return np.load(os.path.join(golden_dir, "vqgan_modules.npz"))


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().numpy()


@pytest.mark.parametrize("kind", ["default", "trained"])
@pytest.mark.parametrize("mname", ["blocky", "iid"])
def test_top_quantizer_indices_match_reference(gq, kind, mname):
    cb = R.codebooks(11, 18, 128, 256, kind).numpy()
    z = R.latent(12, (2, 256, 32, 16), 1.0 if kind == "trained" else 0.02)
    mask = R.blocky_mask(13, 2, 512, 256, 64, extra_ids=(20,)) if mname == "blocky" else R.iid_mask(14, 2, 512, 256)
    ids = vq_oracle.nearest_ids(mask.numpy(), 32, 16)
    r = vq_oracle.search(_nhwc(z), cb, ids, ps=1, cont_stride=1024)
    tag = f"top_{kind}_{mname}"
    assert np.array_equal(r["idx_cont"], gq[tag + "_cont"].astype(np.int64))
    assert np.array_equal(r["idx_list"], gq[tag + "_list"].astype(np.int64))
    numel = z.numel()
    loss = 1.25 * r["sqerr"] / numel
    assert abs(loss - float(gq[tag + "_loss"])) <= 2e-6 * abs(float(gq[tag + "_loss"])) + 1e-12
    assert abs(np.abs(r["zq_nhwc"].astype(np.float64)).sum() - float(gq[tag + "_zq_abs"])) \
        <= 1e-6 * float(gq[tag + "_zq_abs"]) + 1e-9
    if mname == "blocky":
        assert (r["idx_cont"] == -1).any(), "recipe must exercise ids that select no codebook"


@pytest.mark.parametrize("kind", ["default", "trained"])
def test_bottom_quantizer_matches_reference(gq, kind):
    cb = R.codebooks(21, 18, 64, 128, kind).numpy()
    z = R.latent(22, (2, 32, 32, 16), 1.0 if kind == "trained" else 0.02)
    mask = R.blocky_mask(23, 2, 256, 128, 32)
    ids = vq_oracle.nearest_ids(mask.numpy(), 16, 8)
    r = vq_oracle.search(_nhwc(z), cb, ids, ps=2, cont_stride=64)
    # the spatial quantizer returns the continual indices flat (reference :460, no reshape)
    assert gq[f"bot_{kind}_cont"].ndim == 1
    assert np.array_equal(r["idx_cont"].reshape(-1), gq[f"bot_{kind}_cont"].astype(np.int64))
    assert np.array_equal(r["idx_list"], gq[f"bot_{kind}_list"].astype(np.int64))
    zq_nchw = np.transpose(r["zq_nhwc"], (0, 3, 1, 2))
    np.testing.assert_allclose(zq_nchw, gq[f"bot_{kind}_zq"], rtol=0, atol=1e-7)
    # get_codebook_entry (gather + 2x2 fold)
    ent = vq_oracle.gather(cb, r["idx"], ids, 2, 32, 16, 32, ps=2)
    np.testing.assert_array_equal(np.transpose(ent, (0, 3, 1, 2)), gq[f"bot_{kind}_entry"])


@pytest.mark.parametrize("kind", ["default", "trained"])
def test_plain_quantizer_matches_reference(gq, kind):
    cb = R.codebooks(31, 1, 128, 32, kind).numpy()
    z = R.latent(32, (2, 32, 32, 16), 1.0 if kind == "trained" else 0.02)
    r = vq_oracle.search(_nhwc(z), cb, None)
    assert np.array_equal(r["idx"], gq[f"plain_{kind}_idx"].astype(np.int64))


def test_torch_order_quantizer_agrees_with_fixed_order_oracle(gq):
    """the torch-expression restatement (reference summation order on this CPU) and the fixed-order C
    oracle pick the same codes on the recipe"""
    cb = R.codebooks(11, 18, 128, 256, "trained")
    z = R.latent(12, (2, 256, 32, 16), 1.0)
    mask = R.iid_mask(14, 2, 512, 256)
    _, _, cont, _ = vqgan_ref.quantize_texture(cb, z, mask)
    assert np.array_equal(cont.numpy(), gq["top_trained_iid_cont"].astype(np.int64))


def _sd(cfg_cls, cfg, seed):
    m = cfg_cls(**cfg) if isinstance(cfg, dict) else cfg_cls(*cfg)
    return R.fill_state_dict(R.spec_of(m), seed)


def _close(a, b, tol=2e-5):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    assert err <= tol, f"max rel err {err:.3e} > {tol}"


def test_float_modules_match_reference(gm):
    from text2human_b200 import vqgan_arch as A  # CPU construction only: state_dict spec
    with torch.no_grad():
        sd = _sd(A.Encoder, R.TINY_ENC, 41)
        _close(vqgan_ref.encoder(sd, R.image(42, 2, 3, 32, 16)), gm["enc_z"])
        sd = _sd(A.Decoder, R.TINY_DEC, 43)
        z = R.latent(44, (2, 32, 4, 2))
        bot_h = R.latent(45, (2, 64, 8, 4), name="bot_h")
        _close(vqgan_ref.decoder(sd, z), gm["dec_plain"])
        _close(vqgan_ref.decoder(sd, z, bot_h=bot_h), gm["dec_both"])
        _close(vqgan_ref.decoder(sd, z, stop_after_level=4), gm["dec_feature_top"])
        sd = _sd(A.DecoderRes, R.TINY_DECRES, 46)
        _close(vqgan_ref.decoder_res(sd, R.latent(47, (2, 32, 8, 4))), gm["decres"])
        xb = R.latent(49, (2, 64, 16, 8))
        sd = R.fill_state_dict(R.spec_of(A.ResnetBlock(in_channels=64, out_channels=128, temb_channels=0,
                                                       dropout=0.0)), 48)
        _close(vqgan_ref.resnet_block(sd, "", xb), gm["resblock"])
        sd = R.fill_state_dict(R.spec_of(A.AttnBlock(64)), 50)
        _close(vqgan_ref.attn_block(sd, "", xb), gm["attnblock"])
        sd = R.fill_state_dict(R.spec_of(A.Upsample(64, True)), 51)
        _close(vqgan_ref.conv(sd, "conv", torch.nn.functional.interpolate(xb, scale_factor=2.0, mode="nearest")),
               gm["upsample"])
        sd = R.fill_state_dict(R.spec_of(A.Downsample(64, True)), 52)
        _close(vqgan_ref.conv(sd, "conv", torch.nn.functional.pad(xb, (0, 1, 0, 1)), stride=2, padding=0),
               gm["downsample"])


def test_transformer_matches_reference(golden_dir):
    from text2human_b200 import transformer_arch as T
    g = np.load(os.path.join(golden_dir, "transformer.npz"))
    sd = R.fill_state_dict(R.spec_of(T.TransformerMultiHead(**R.TINY_TRANSFORMER)), 61)
    gen = R._gen(62, "tokens")
    idx = torch.randint(0, 18 * 16 + 1, (2, 32), generator=gen)
    segm = torch.randint(0, 32, (2, 32), generator=gen)
    tex = torch.randint(0, 18, (2, 32), generator=gen)
    with torch.no_grad():
        lg = torch.stack(transformer_ref.transformer_logits(sd, idx, segm, tex, n_head=4))
    _close(lg, g["logits"])
