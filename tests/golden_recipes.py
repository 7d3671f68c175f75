"""Deterministic recipes for weights and inputs shared by oracle/make_golden.py
(which runs the real reference on them) and the tests (which re-create the same
tensors without the reference).  All randomness is torch CPU generators seeded
from (seed, crc32(name)), so values do not depend on module construction order.
"""
import zlib

import torch


def _gen(seed, name):
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2**63 - 1))
    return g


def fill_state_dict(spec, seed):
    """spec: iterable of (key, shape).  conv/linear/embedding weights ~ N(0, 1/fan_in) (embeddings N(0,1)*0.5),
    biases ~ 0.1*N(0,1), norm weights 1 + 0.1*N(0,1), pos_emb/start_tok 0.02*N(0,1); BatchNorm buffers:
    running_mean 0.1*N(0,1), running_var U(0.5,1.5), num_batches_tracked 0."""
    sd = {}
    for key, shape in spec:
        shape = tuple(shape)
        g = _gen(seed, key)
        r = torch.randn(shape, generator=g)
        leaf = key.split(".")[-1]
        parent = key.split(".")[-2] if "." in key else ""
        if leaf == "num_batches_tracked":
            sd[key] = torch.zeros(shape, dtype=torch.long)
            continue
        if leaf == "running_var":  # BatchNorm running variance: positive, around 1
            sd[key] = (0.5 + torch.rand(shape, generator=g)).float()
            continue
        if key in ("pos_emb", "start_tok"):
            v = 0.02 * r
        elif "emb" in parent or "embedding" in parent:
            v = 0.5 * r
        elif leaf == "weight" and len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            v = r / fan_in ** 0.5
        elif leaf == "weight":  # norm scale
            v = 1.0 + 0.1 * r
        else:
            v = 0.1 * r
        sd[key] = v.float()
    return sd


def spec_of(module):
    return [(k, tuple(v.shape)) for k, v in module.state_dict().items()]


def image(seed, b, c, h, w):
    return torch.rand((b, c, h, w), generator=_gen(seed, "image")) * 2 - 1


def latent(seed, shape, scale=1.0, name="latent"):
    return torch.randn(shape, generator=_gen(seed, name)) * scale


def blocky_mask(seed, b, h, w, tile, n_ids=18, extra_ids=()):
    """float id map [b,1,h,w], constant on tile x tile blocks; ids uniform in [0,n_ids) plus optional
    out-of-range ids (which select no codebook)."""
    g = _gen(seed, "mask")
    pool = list(range(n_ids)) + list(extra_ids)
    pick = torch.randint(0, len(pool), (b, (h + tile - 1) // tile, (w + tile - 1) // tile), generator=g)
    ids = torch.tensor(pool, dtype=torch.float32)[pick]
    m = ids.repeat_interleave(tile, 1).repeat_interleave(tile, 2)[:, :h, :w]
    return m.unsqueeze(1).contiguous()


def iid_mask(seed, b, h, w, n_ids=18):
    g = _gen(seed, "iid_mask")
    return torch.randint(0, n_ids, (b, 1, h, w), generator=g).float()


def codebooks(seed, n_books, n_e, d, kind):
    """kind 'default': the reference init uniform(-1/n_e, 1/n_e) (vqgan_arch.py:169);
    kind 'trained': rows ~ N(0, 1) (spread like encoder outputs, so distances are not near-ties)."""
    g = _gen(seed, "codebooks_" + kind)
    if kind == "default":
        return (torch.rand((n_books, n_e, d), generator=g) * 2 - 1) / n_e
    return torch.randn((n_books, n_e, d), generator=g)


# ---------------------------------------------------------------- model configs
TINY_ENC = dict(ch=64, num_res_blocks=1, attn_resolutions=[8], in_channels=3, resolution=32, z_channels=32,
                ch_mult=[1, 2, 2], double_z=False, dropout=0.0)          # x [B,3,32,16] -> z [B,32,8,4]
TINY_DEC = dict(in_channels=3, resolution=64, z_channels=32, ch=32, out_ch=3, num_res_blocks=1,
                attn_resolutions=[4], ch_mult=[1, 1, 1, 2, 2], dropout=0.0, resamp_with_conv=True,
                give_pre_end=False)                                       # z [B,32,4,2] -> [B,3,64,32]
TINY_DECRES = dict(in_channels=3, resolution=32, z_channels=32, ch=32, num_res_blocks=1, ch_mult=[1, 2, 2],
                   dropout=0.0, give_pre_end=False)                       # z [B,32,8,4] -> [B,64,8,4]
TINY_TRANSFORMER = dict(codebook_size=18 * 16, segm_codebook_size=32, texture_codebook_size=18, bert_n_emb=64,
                        bert_n_layers=2, bert_n_head=4, block_size=32, latent_shape=[8, 4], embd_pdrop=0.0,
                        resid_pdrop=0.0, attn_pdrop=0.0, num_head=18)


def sampler_train_batch(seed, B=2, cfg=None):
    """deterministic training batch for the index-prediction transformer: continual tokens x_0, the 18
    per-texture ground-truth lists (-1 outside the texture), segm and texture tokens"""
    cfg = cfg or TINY_TRANSFORMER
    T = cfg["block_size"]
    ncls = cfg["codebook_size"] // cfg["num_head"]
    g = _gen(seed, "train_batch")
    tex = torch.randint(0, cfg["num_head"], (B, T), generator=g)
    own = torch.randint(0, ncls, (B, T), generator=g)
    segm = torch.randint(0, cfg["segm_codebook_size"], (B, T), generator=g)
    x_0 = own + ncls * tex
    gt_list = [torch.where(tex == k, own, torch.full_like(own, -1)) for k in range(cfg["num_head"])]
    return x_0, gt_list, segm, tex


# reduced index-prediction nets for the fixture (the real ones: UNet(in_channels=256) with base 64, and
# MultiHeadFCNHead(in_channels=64, channels=64, num_classes=512), configs/sample_from_parsing.yml:49-58)
TINY_UNET = dict(in_channels=32, base_channels=8)
TINY_FCN = dict(in_channels=8, channels=8, in_index=4, num_convs=1, concat_input=False, dropout_ratio=0.1,
                num_classes=16, align_corners=False, num_head=18)
REAL_UNET = dict(in_channels=256)
REAL_FCN = dict(in_channels=64, channels=64, in_index=4, num_convs=1, concat_input=False, dropout_ratio=0.1,
                num_classes=512, align_corners=False, num_head=18)


# reduced VQGAN + discriminator for the GAN-training-step fixture (real: configs/vqvae_top.yml, ndf 64, 3 layers)
TINY_VQGAN_TRAIN = dict(
    enc=dict(ch=32, num_res_blocks=1, attn_resolutions=[8], in_channels=3, resolution=64, z_channels=32,
             ch_mult=[1, 2, 2, 4], double_z=False, dropout=0.0),                    # x [B,3,64,32] -> z [B,32,8,4]
    dec=dict(in_channels=3, resolution=64, z_channels=32, ch=32, out_ch=3, num_res_blocks=1, attn_resolutions=[8],
             ch_mult=[1, 2, 2, 4], dropout=0.0, resamp_with_conv=True, give_pre_end=False),
    n_embed=64, embed_dim=32, ndf=16, disc_layers=3, disc_start_step=0, step=5, batch=2)
# Global-RNG seed under which the reference draws DiffAugment's brightness / saturation / contrast / translation for
# the fixture.  The step's gradient is discontinuous at the LeakyReLU / hinge kinks of the discriminator; the seed was
# chosen (scan of 109..399 with the restatement) so that no pre-activation of the three discriminator passes is within
# 7e-5 of a kink -- an fp32-equivalent implementation (activation error ~1e-5) then cannot land on the other side of
# one, which the first fixture (seed 109: a pre-activation at +1.5e-5) made a coin flip.
VQGAN_TRAIN_AUG_SEED = 153


# reduced index-prediction transformer for the sample_fn fixture: the reference loop hard-codes the 32x16 token grid
# and the 1024-per-texture index stride (sample_model.py:270,313), so those are kept; width / depth are reduced
SAMPLE_TRANSFORMER = dict(codebook_size=18432, segm_codebook_size=32, texture_codebook_size=18, bert_n_emb=64,
                          bert_n_layers=2, bert_n_head=4, block_size=512, latent_shape=[32, 16], embd_pdrop=0.0,
                          resid_pdrop=0.0, attn_pdrop=0.0, num_head=18)
SAMPLE_BATCH, SAMPLE_STEPS = 2, 12


def sample_inputs(seed, B):
    g = _gen(seed, "sample_inputs")
    segm_tokens = torch.randint(0, 32, (B, 512), generator=g)
    return segm_tokens, blocky_mask(seed, B, 512, 256, 64)


def dataset_items(seed, n, h, w):
    """synthetic dataset items: uint8 HWC images, parsing maps with class ids 0..23 on 4x4 blocks, fused attributes
    (upper, lower, outer) in 0..17 (17 = 'NA': that clothes group keeps the common codebook)"""
    import numpy as np
    g = _gen(seed, "dataset_items")
    imgs = torch.randint(0, 256, (n, h, w, 3), generator=g).numpy().astype(np.uint8)
    seg = torch.randint(0, 24, (n, h // 4, w // 4), generator=g).repeat_interleave(4, 1).repeat_interleave(4, 2)
    attrs = torch.randint(0, 18, (n, 3), generator=g).numpy().astype(np.int64)
    return imgs, seg.numpy().astype(np.int64), attrs
