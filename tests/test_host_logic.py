"""CPU checks of the host-side logic: fp16 plane splitting, weight packing, packed-weight cache
invalidation, state_dict (checkpoint ABI) parity with the reference key names."""
import numpy as np
import torch

import golden_recipes as R
from text2human_b200 import ops
from text2human_b200 import transformer_arch as T
from text2human_b200 import vqgan_arch as A


def test_split_planes_reconstructs_fp32_to_22_bits():
    x = torch.randn(1000) * 3
    p = ops.split_planes(x, 2)
    assert p.shape == (2, 1000) and p.dtype == torch.float16
    # hi carries 11 bits, lo the next 11 (until lo hits the fp16 subnormal step 2^-24)
    err = (p.float().sum(0) - x).abs()
    assert (err <= x.abs() * 2 ** -21 + 2 ** -24).all()
    assert torch.equal(ops.split_planes(x, 1)[0], x.half())


def test_pack_conv_weight_layout_and_padding():
    w = torch.randn(5, 3, 3, 3)
    p = ops.pack_conv_weight(w, 2)
    assert p.shape == (2, 9, 5, 8)
    rec = p.float().sum(0)
    for kh in range(3):
        for kw in range(3):
            assert torch.allclose(rec[kh * 3 + kw, :, :3], w[:, :, kh, kw], atol=1e-6)
    assert (rec[..., 3:] == 0).all()
    l = ops.pack_linear_weight(torch.randn(7, 16, 1, 1), 1)
    assert l.shape == (1, 1, 7, 16)


def test_packed_weight_cache_tracks_in_place_updates():
    conv = torch.nn.Conv2d(8, 8, 3)
    a = A._conv_w(conv)
    assert A._conv_w(conv) is a
    with torch.no_grad():
        conv.weight.add_(1.0)
    b = A._conv_w(conv)
    assert b is not a
    assert not torch.equal(a, b)


def test_state_dict_keys_follow_reference_checkpoint_abi():
    enc = A.Encoder(ch=128, num_res_blocks=2, attn_resolutions=[32], in_channels=3, resolution=512,
                    z_channels=256, ch_mult=[1, 1, 2, 2, 4], double_z=False)
    keys = list(enc.state_dict())
    assert len(keys) == 144
    assert sum(p.numel() for p in enc.parameters()) == 29298176
    for k in ("conv_in.weight", "down.0.block.0.norm1.weight", "down.4.attn.1.proj_out.bias",
              "down.3.downsample.conv.weight", "mid.attn_1.q.weight", "norm_out.bias", "conv_out.weight",
              "down.2.block.0.nin_shortcut.weight"):
        assert k in keys, k
    dec = A.Decoder(in_channels=3, resolution=512, z_channels=256, ch=128, out_ch=3, num_res_blocks=2,
                    attn_resolutions=[32], ch_mult=[1, 1, 2, 2, 4])
    assert sum(p.numel() for p in dec.parameters()) == 42450307 or abs(
        sum(p.numel() for p in dec.parameters()) / 1e6 - 42.45) < 0.01
    assert "up.4.attn.2.k.weight" in dec.state_dict() and "up.1.upsample.conv.bias" in dec.state_dict()
    q = A.VectorQuantizerTexture(1024, 256, 0.25)
    assert list(q.state_dict())[0] == "embedding_list.0.weight" and len(q.state_dict()) == 18
    qb = A.VectorQuantizerSpatialTextureAware(512, 256, 0.25, spatial_size=2)
    assert qb.state_dict()["embedding_list.17.weight"].shape == (512, 1024)
    tf = T.TransformerMultiHead(codebook_size=18432, segm_codebook_size=1024, texture_codebook_size=18,
                                bert_n_emb=512, bert_n_layers=2, bert_n_head=8, block_size=512,
                                latent_shape=[32, 16], embd_pdrop=0, resid_pdrop=0, attn_pdrop=0, num_head=18)
    sd = tf.state_dict()
    assert sd["tok_emb.weight"].shape == (18433, 512) and sd["pos_emb"].shape == (1, 512, 512)
    assert sd["head_list.17.weight"].shape == (1024, 512)
    for k in ("blocks.0.attn.key.weight", "blocks.1.mlp.2.bias", "ln_f.weight", "start_tok",
              "segm_emb.weight", "texture_emb.weight"):
        assert k in sd
    assert (sd["pos_emb"] == 0).all() and (sd["start_tok"] == 0).all()  # reference never applies _init_weights


def test_default_codebook_init_matches_reference_range():
    q = A.VectorQuantizerTexture(1024, 256, 0.25)
    w = q.embedding_list[3].weight
    assert w.abs().max() <= 1.0 / 1024 and w.abs().max() > 0.9 / 1024


def test_recipe_is_deterministic():
    a = R.fill_state_dict([("x.weight", (4, 3, 3, 3)), ("x.bias", (4,))], 5)
    b = R.fill_state_dict([("x.bias", (4,)), ("x.weight", (4, 3, 3, 3))], 5)
    assert all(torch.equal(a[k], b[k]) for k in a)
    m = R.blocky_mask(1, 2, 64, 32, 16, extra_ids=(20,))
    assert m.shape == (2, 1, 64, 32) and set(np.unique(m.numpy())).issubset(set(range(18)) | {20})


def test_group_by_texture_places_every_position_once_inside_its_head_group():
    """host logic of the sampler's own-head GEMM: rows grouped by texture id, groups padded to 128-row tiles"""
    import torch
    from text2human_b200.transformer_arch import TransformerMultiHead
    g = torch.Generator().manual_seed(3)
    for tex in (torch.randint(0, 18, (4, 512), generator=g), torch.full((2, 512), 7),
                torch.randint(0, 3, (1, 40), generator=g)):
        dest, rows = TransformerMultiHead.group_by_texture(tex, 18)
        flat = tex.reshape(-1)
        assert rows % 128 == 0 and rows >= int(torch.bincount(flat, minlength=18).max())
        assert dest.unique().numel() == flat.numel()                      # no two positions share a row
        assert torch.equal(dest // rows, flat)                            # each inside its own head's group
        for k in range(18):                                               # groups are filled from their start,
            r = (dest[flat == k] % rows).sort().values                    # in position order (stable)
            assert torch.equal(r, torch.arange(r.numel()))
            assert torch.equal(dest[flat == k], dest[flat == k].sort().values)


def test_split_k_switches_and_choice():
    from text2human_b200 import ops
    assert ops.SPLIT_K == {"wgrad": True, "inference": False, "small_batch": True}
    old = ops.set_split_k(inference=True, wgrad=False)
    assert ops.SPLIT_K == {"wgrad": False, "inference": True, "small_batch": True}
    assert old == {"wgrad": True, "inference": False, "small_batch": True}
    ops.set_split_k(**old)
    assert ops.SPLIT_K == old
    # the deterministic (stored-partials) form: slices the kernel produces for a k_partials request
    assert ops.split_slices(512, 4) == 4 and ops.split_slices(2048, 4) == 4 and ops.split_slices(64, 4) == 1
    assert ops.split_slices(8 * 64, 3) == 3 and ops.split_slices(7 * 64, 4) == 4 and ops.split_slices(5 * 64, 4) == 3
    assert ops.wgrad_k_split(512, 512, 8192) == 18        # 8 output tiles: up to 148 // 8 slices (the kernel
                                                          # rounds to 16 slices of 8 chunks)
    assert ops.wgrad_k_split(18432, 512, 8192) == 0       # 288 tiles already fill the GPU
    assert ops.wgrad_k_split(2048, 512, 2048) == 4        # fc2 at B=4
    assert ops.wgrad_k_split(64, 64, 64) == 0             # a single K chunk cannot be split


def test_conv_data_gradient_is_a_forward_conv_with_flipped_transposed_weights():
    """host logic for the conv backward plan: dX of a stride-1 'same' conv equals the forward conv of
    dY with ops.conv_weight_for_dgrad(W), for 3x3 and 1x1 kernels"""
    import torch
    import torch.nn.functional as F
    from text2human_b200 import ops
    g = torch.Generator().manual_seed(1)
    for k, cin, cout in ((3, 5, 7), (1, 4, 6)):
        x = torch.randn(2, cin, 9, 6, generator=g, requires_grad=True)
        w = torch.randn(cout, cin, k, k, generator=g)
        dy = torch.randn(2, cout, 9, 6, generator=g)
        F.conv2d(x, w, padding=k // 2).backward(dy)
        wd = ops.conv_weight_for_dgrad(w)
        assert wd.shape == (cin, cout, k, k)
        got = F.conv2d(dy, wd, padding=k // 2)
        assert (got - x.grad).abs().max() <= 1e-4 * x.grad.abs().max()
        assert ops.pack_conv_weight(wd, 2).shape == (2, k * k, cin, (cout + 7) // 8 * 8)


def test_fusion_gates_choose_the_kernels_shape_constraints():
    """which launches the host logic hands to the fused kernels: t2h_attn_fwd needs head_dim 64 and 128..512 tokens
    in multiples of 128 (the score rows of a query block fill tensor memory); the norm-backward sums ride only on
    data-gradient convs the swapped-operand kernel runs (channels % 128, more than one image row)"""
    import torch
    from text2human_b200 import ops
    assert ops.can_fuse_attn(512, 64) and ops.can_fuse_attn(128, 64) and ops.can_fuse_attn(384, 64)
    assert not ops.can_fuse_attn(640, 64) and not ops.can_fuse_attn(64, 64) and not ops.can_fuse_attn(192, 64)
    assert not ops.can_fuse_attn(512, 128) and not ops.can_fuse_attn(512, 32)
    old = ops.set_fused_attn(False)
    try:
        assert not ops.can_fuse_attn(512, 64)
    finally:
        ops.set_fused_attn(old)
    gamma, beta = torch.ones(128), torch.zeros(128)
    st = torch.zeros(2, 32, 2, dtype=torch.float64)
    ctx = ops.nb_context(torch.zeros(2, 8, 4, 128), st, gamma, beta, act="swish", groups=32, eps=1e-6)
    assert ctx is not None and ctx["sums"].shape == (2 * 128 * 2,) and ctx["sums"].dtype == torch.float64
    assert float(ctx["sums"].abs().sum()) == 0.0            # the epilogue accumulates into it
    assert ops.nb_context(torch.zeros(2, 8, 4, 64), st, gamma[:64], beta[:64], act="swish", groups=32, eps=1e-6) is None
    assert ops.nb_context(torch.zeros(2, 1, 4, 128), st, gamma, beta, act="swish", groups=32, eps=1e-6) is None
    old_nb = ops.FUSE_NB["on"]
    ops.FUSE_NB["on"] = False
    try:
        assert ops.nb_context(torch.zeros(2, 8, 4, 128), st, gamma, beta, act="swish", groups=32, eps=1e-6) is None
    finally:
        ops.FUSE_NB["on"] = old_nb


def test_branch_free_gelu_constants_reproduce_erf_gelu():
    """the tap-GEMM epilogue's GELU (csrc/t2h_ptx.cuh:gelu_erf, Abramowitz & Stegun 7.1.26 evaluated through erfc):
    the constants as written in the kernel source, restated in float32 numpy, against torch's exact-erf nn.GELU
    (transformer_arch.py:85) over [-8, 8] -- the absolute error must stay at the fp32 formula's own level"""
    import os
    import re
    import numpy as np
    import torch
    src = open(os.path.join(os.path.dirname(__file__), "..", "text2human_b200", "csrc", "t2h_ptx.cuh")).read()
    body = src[src.index("__device__ __forceinline__ float gelu_erf(float x)"):]
    body = body[:body.index("\n}\n")]
    nums = [float(v) for v in re.findall(r"(-?\d+\.\d+)f", body)]
    # sqrt(1/2), p, 1, a5, a4, a3, a2, a1, log2(e), 0.5, 2
    inv_sqrt2, p, one, a5, a4, a3, a2, a1, log2e = nums[:9]
    assert abs(inv_sqrt2 - 2 ** -0.5) < 1e-9 and one == 1.0 and abs(log2e - 1.4426950408889634) < 1e-9
    f = np.float32
    x = np.linspace(-8, 8, 400001).astype(np.float32)
    z = np.abs(x) * f(inv_sqrt2)
    t = f(1) / (f(p) * z + f(1))
    poly = ((((f(a5) * t + f(a4)) * t + f(a3)) * t + f(a2)) * t + f(a1))
    e = np.exp2((-(z * z)) * f(log2e)).astype(np.float32)
    c = poly * t * e
    got = (f(0.5) * x * np.where(x >= 0, f(2) - c, c)).astype(np.float32)
    ref = torch.nn.functional.gelu(torch.from_numpy(x).double()).numpy()
    assert np.abs(got - ref).max() < 6e-7
    assert np.abs(got[np.abs(x) < 1] - ref[np.abs(x) < 1]).max() < 2e-7
