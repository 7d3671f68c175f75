"""Deterministic split-K of the small-batch transformer: t2h_tapgemm with k_partials stores every
k-slice to its own slab, t2h_splitk_reduce_ln sums them in slab order, adds bias + residual and applies the next
LayerNorm (reference: Block.forward, transformer_arch.py:91-99 -- x = x + proj(...); ln2(x))."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    got, ref = got.double(), ref.double()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("M,K,N,ks", [(2048, 512, 512, 4), (2048, 2048, 512, 4), (384, 320, 1024, 3), (256, 64, 512, 4)])
def test_partial_gemm_reduce_layernorm_matches_fp64(cuda, M, K, N, ks):
    from text2human_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(M + K + N)
    a = torch.randn(M, K, device=cuda, generator=g)
    w = torch.randn(N, K, device=cuda, generator=g) / K ** 0.5
    bias, gamma, beta = (torch.randn(N, device=cuda, generator=g) for _ in range(3))
    res = torch.randn(M, N, device=cuda, generator=g)
    ap = ops.split_planes(a, 2)
    wp = ops.pack_linear_weight(w, 2)
    part = ops.linear_partials(ap, wp, ks)
    assert part.shape == (ops.split_slices(K, ks), M, N)
    x, ln = ops.splitk_reduce_ln(part, bias, res, gamma, beta, 1e-5, terms=2)
    xr = res.double() + bias.double() + ap.double().sum(0) @ wp.double().sum(0).reshape(N, K).t()
    assert _rel(x, xr) < 2e-6
    lr = F.layer_norm(xr, (N,), gamma.double(), beta.double(), 1e-5)
    assert _rel(ln.double().sum(0), lr) < 5e-6
    # bit-reproducible: slabs are summed in index order, nothing is reduce-added
    part2 = ops.linear_partials(ap, wp, ks)
    x2, ln2 = ops.splitk_reduce_ln(part2, bias, res, gamma, beta, 1e-5, terms=2)
    assert torch.equal(x, x2) and torch.equal(ln, ln2)
    # the sum itself, in fp32 and in the kernel's order
    xs = res + bias
    for s in range(part.shape[0]):
        xs = xs + part[s]
    assert torch.equal(x, xs)


def test_reduce_layernorm_scalar_path_and_row_scatter(cuda):
    """C % 4 != 0 takes the scalar kernel; row_map scatters the normalised rows (grouped-by-texture head GEMM)"""
    from text2human_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(7)
    for Cc in (510, 512):
        M = 96
        part = torch.randn(3, M, Cc, device=cuda, generator=g)
        res = torch.randn(M, Cc, device=cuda, generator=g)
        bias, gamma, beta = (torch.randn(Cc, device=cuda, generator=g) for _ in range(3))
        x, ln = ops.splitk_reduce_ln(part, bias, res, gamma, beta, 1e-5, terms=2)
        xr = ((res + bias) + part[0] + part[1]) + part[2]
        assert torch.equal(x, xr)
        assert _rel(ln.double().sum(0), F.layer_norm(xr.double(), (Cc,), gamma.double(), beta.double(), 1e-5)) < 5e-6
        perm = torch.randperm(M + 32, device=cuda, generator=g)[:M].contiguous()
        out = torch.zeros(2, M + 32, Cc, dtype=torch.float16, device=cuda)
        ops.splitk_reduce_ln(part, bias, res, gamma, beta, 1e-5, ln_out=out, row_map=perm)
        assert torch.equal(out[:, perm], ln)
        # in place on the residual stream (x_out may alias residual: the transformer passes the same tensor?)
        x3, _ = ops.splitk_reduce_ln(part, None, None, gamma, beta, 1e-5, terms=1)
        assert torch.equal(x3, (part[0] + part[1]) + part[2])
