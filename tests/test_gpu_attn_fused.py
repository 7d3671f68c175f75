"""t2h_attn_fwd (csrc/attn_fused.cuh): the transformer's multi-head attention as one kernel, against fp64 PyTorch of
the reference's formula (transformer_arch.py:41-67: q k^T / sqrt(hs), softmax over keys, att v, heads re-assembled)
and against the three-launch path it replaces (t2h_tapgemm q k^T, t2h_softmax_rows, t2h_tapgemm p v)."""
import math

import pytest
import torch

import golden_recipes as R

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    got, ref = got.double(), ref.double()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def _reference(qkv_exact, B, Tn, nh, scale):
    """fp64 attention of the values the planes actually hold"""
    Cc = qkv_exact.shape[1] // 3
    q, k, v = (qkv_exact[:, i * Cc:(i + 1) * Cc].view(B, Tn, nh, 64).transpose(1, 2) for i in range(3))
    att = torch.softmax((q @ k.transpose(-2, -1)) * scale, dim=-1)
    return (att @ v).transpose(1, 2).reshape(B * Tn, Cc)


@pytest.mark.parametrize("terms,B,Tn,nh,spread", [(2, 4, 512, 8, 1.0), (2, 2, 512, 8, 4.0), (2, 3, 256, 4, 2.0),
                                                  (2, 1, 128, 2, 1.0), (2, 2, 384, 3, 3.0), (1, 2, 512, 8, 1.0)])
def test_fused_attention_matches_fp64_and_the_three_launch_path(cuda, terms, B, Tn, nh, spread):
    from text2human_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(1000 * Tn + nh)
    Cc = nh * 64
    # `spread` widens the logits (q k^T / 8 std ~16 at 4.0): rows dominated by one key, tiny tails
    qkv = torch.randn(B * Tn, 3 * Cc, device=cuda, generator=g)
    qkv[:, :2 * Cc] *= spread
    planes = ops.split_planes(qkv, terms)
    exact = planes.double().sum(0)
    scale = 1.0 / math.sqrt(64)
    ref = _reference(exact, B, Tn, nh, scale)

    y = ops.attn_fused(planes, B, Tn, nh, scale)
    assert y.shape == (terms, B * Tn, Cc)
    got = y.double().sum(0)
    # the dropped lo*lo products leave ~2^-22 * sum|q_i k_i| in a score: the exponent error grows with the spread
    tol = (2e-5 if terms == 2 else 2e-3) * max(1.0, spread)
    e_fused = _rel(got, ref)

    s = ops.mha_scores(planes[:, :, :Cc], B, Tn, nh, k=planes[:, :, Cc:2 * Cc])
    p = ops.softmax_rows(s, scale=scale, terms=terms)
    y3 = ops.mha_pv(p, planes[:, :, 2 * Cc:], B, Tn, nh, v_tok=True)
    e_three = _rel(y3.double().sum(0), ref)
    print(f"[attn] terms={terms} B={B} T={Tn} heads={nh} spread={spread}: fused {e_fused:.2e}  three-launch {e_three:.2e}")
    assert e_fused < tol
    assert _rel(got, y3.double().sum(0)) < 2 * tol
    # every row is a convex combination of the value rows of its sequence and head
    v = exact[:, 2 * Cc:].view(B, Tn, Cc)
    assert (got.view(B, Tn, Cc) <= v.max(1, keepdim=True).values + 1e-4).all()
    assert (got.view(B, Tn, Cc) >= v.min(1, keepdim=True).values - 1e-4).all()


def test_fused_attention_rejects_unsupported_shapes(cuda):
    from text2human_b200 import ops, _lib
    planes = torch.zeros(2, 192, 3 * 64, dtype=torch.float16, device=cuda)
    assert not ops.can_fuse_attn(192, 64) and not ops.can_fuse_attn(512, 32) and ops.can_fuse_attn(512, 64)
    with pytest.raises(_lib.T2HError):
        ops.attn_fused(planes, 1, 192, 1, 0.125)


def test_transformer_logits_equal_with_and_without_fused_attention(cuda):
    """BASELINE config 4's transformer (24 layers, 8 heads x 64, 512 tokens): logits with the fused attention kernel
    vs the three-launch path, both in parity mode."""
    from text2human_b200 import ops
    from text2human_b200.transformer_arch import TransformerMultiHead
    ops.set_precision("fp32")
    cfg = dict(codebook_size=18432, segm_codebook_size=1024, texture_codebook_size=18, bert_n_emb=512,
               bert_n_layers=4, bert_n_head=8, block_size=512, latent_shape=[32, 16], embd_pdrop=0.0,
               resid_pdrop=0.0, attn_pdrop=0.0, num_head=18)
    net = TransformerMultiHead(**cfg)
    net.load_state_dict(R.fill_state_dict(R.spec_of(net), 91), strict=True)
    net = net.to(cuda).eval()
    g = torch.Generator().manual_seed(92)
    idx = torch.randint(0, 18433, (2, 512), generator=g).to(cuda)
    segm = torch.randint(0, 1024, (2, 512), generator=g).to(cuda)
    tex = torch.randint(0, 18, (2, 512), generator=g).to(cuda)
    outs = []
    for on in (True, False):
        old = ops.set_fused_attn(on)
        try:
            l0 = ops.COUNTERS["launches"]
            with torch.no_grad():
                outs.append(torch.stack(net(idx, segm, tex)))
            n = ops.COUNTERS["launches"] - l0
        finally:
            ops.set_fused_attn(old)
        print(f"[attn] fused={on}: {n} libt2h launches per forward")
    e = _rel(outs[0], outs[1])
    print(f"[attn] logits fused vs three-launch: {e:.2e}")
    assert e < 2e-5
