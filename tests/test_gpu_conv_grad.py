"""GPU parity of the training-step kernels (through libt2h) against torch autograd of the same op:
conv forward / data gradient / weight gradient for every conv kind of the VQGAN + Discriminator, the norm
(GroupNorm / BatchNorm + activation) backward, and the small loss / augmentation kernels.
Reference: what autograd derives for the layers of models/archs/vqgan_arch.py and models/losses/vqgan_loss.py."""
import pytest
import torch
import torch.nn.functional as F

import golden_recipes as R

pytestmark = pytest.mark.gpu

TOL = {2: 3e-5, 1: 4e-3}      # 3-product (fp32-equivalent) / single-product operands; reference = fp64 autograd


def _rel(got, ref):
    got, ref = got.double(), ref.double()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def _torch_conv(kind, x, w, b):
    if kind == "k3":
        return F.conv2d(x, w, b, padding=1)
    if kind == "k1":
        return F.conv2d(x, w, b)
    if kind == "down":
        return F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=2)
    if kind == "k4s2":
        return F.conv2d(x, w, b, stride=2, padding=1)
    return F.conv2d(x, w, b, stride=1, padding=1)


CASES = [  # kind, N, Cin, Cout, H, W
    ("k3", 2, 64, 128, 16, 8), ("k3", 1, 128, 128, 32, 16), ("k3", 2, 32, 64, 4, 2), ("k3", 2, 3, 64, 16, 8),
    ("k3", 2, 64, 3, 16, 8), ("k3", 1, 256, 512, 8, 4), ("k1", 2, 64, 128, 8, 4), ("down", 2, 64, 64, 16, 8),
    ("down", 1, 128, 128, 64, 32), ("k4s2", 2, 3, 16, 32, 16), ("k4s2", 2, 64, 128, 16, 8),
    ("k4s1", 2, 64, 128, 9, 5), ("k4s1", 2, 128, 1, 8, 4), ("k3", 2, 128, 128, 64, 32),
]


@pytest.mark.parametrize("terms", [2, 1])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(str(v) for v in c))
def test_conv_forward_dgrad_wgrad_match_autograd(cuda, case, terms):
    from text2human_b200 import conv_grad as G
    from text2human_b200 import ops
    kind, N, Ci, Co, H, W = case
    K = G.KSIZE[kind]
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(repr(case).encode()) % 1000)
    x = torch.randn(N, Ci, H, W, generator=g).to(cuda).requires_grad_(True)
    w = (torch.randn(Co, Ci, K, K, generator=g) / (Ci * K * K) ** 0.5).to(cuda).requires_grad_(True)
    b = (0.1 * torch.randn(Co, generator=g)).to(cuda).requires_grad_(True)
    x64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    y = _torch_conv(kind, x64, w64, b64)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(3)).to(cuda)
    y.backward(dy.double())
    xg, wg, y = x64.grad, w64.grad, y.detach()
    Ho, Wo = G.out_hw(kind, H, W)
    assert y.shape[2:] == (Ho, Wo)
    # ---- operands as the training step holds them
    m = G.oihw_to_master(w)
    wn, wt = G.weight_planes(m, terms)
    xp = ops.nchw_to_planes(x.detach(), terms=terms)                      # [T,N,H,W,Ci_p]
    a = ops.planes_s2d(xp) if G.STRIDE[kind] == 2 else xp
    bias = torch.zeros(G.pad8(Co), device=cuda)
    bias[:Co] = b.detach()
    out = G.forward(kind, a, wn, bias, n=N, in_hw=(H, W))                  # fp32 NHWC [N,Ho,Wo,Co_p]
    assert out.shape == (N, Ho, Wo, G.pad8(Co))
    assert _rel(out[..., :Co].permute(0, 3, 1, 2), y.detach()) < TOL[terms]
    if G.pad8(Co) != Co:
        assert float(out[..., Co:].abs().max()) == 0.0
    dyp = ops.nchw_to_planes(dy, terms=terms)                              # [T,N,Ho,Wo,Co_p]
    # ---- weight gradient
    gw = torch.zeros_like(m)
    G.wgrad(kind, dyp, a, gw, n=N)
    want_gw = G.oihw_to_master(wg)
    assert _rel(gw, want_gw) < TOL[terms], "wgrad"
    G.wgrad(kind, dyp, a, gw, n=N, alpha=0.5)                              # accumulates
    assert _rel(gw, 1.5 * want_gw) < TOL[terms], "wgrad accumulate"
    # ---- data gradient
    dx = G.dgrad(kind, dyp, wt, n=N, in_hw=(H, W))
    assert dx.shape == (N, H, W, G.pad8(Ci))
    assert _rel(dx[..., :Ci].permute(0, 3, 1, 2), xg) < TOL[terms], "dgrad"
    if Ci == 3:   # image gradient written straight into NCHW
        dx_img = torch.zeros((N, 3, H, W), device=cuda)
        G.dgrad(kind, dyp, wt, n=N, in_hw=(H, W), cin=3, out=dx_img, d_strides=(3 * H * W, W, 1, H * W))
        assert _rel(dx_img, xg) < TOL[terms], "dgrad nchw"


@pytest.mark.parametrize("act,groups_mode", [("swish", "gn"), (None, "gn"), ("lrelu", "bn")])
def test_norm_backward_matches_autograd(cuda, act, groups_mode):
    from text2human_b200 import ops
    N, H, W, Cc = 3, 16, 8, 64
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, Cc, H, W, generator=g).to(cuda).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(Cc, generator=g)).to(cuda).requires_grad_(True)
    beta = (0.1 * torch.randn(Cc, generator=g)).to(cuda).requires_grad_(True)
    dy = torch.randn(N, Cc, H, W, generator=g).to(cuda)
    add = torch.randn(N, H, W, Cc, generator=g).to(cuda)
    if groups_mode == "gn":
        u = F.group_norm(x, 32, gamma, beta, eps=1e-6)
        groups, eps, n = 32, 1e-6, None
    else:
        u = F.batch_norm(x, None, None, gamma, beta, True, 0.1, 1e-5)
        groups, eps, n = Cc, 1e-5, 1
    yy = u * torch.sigmoid(u) if act == "swish" else (F.leaky_relu(u, 0.2) if act == "lrelu" else u)
    yy.backward(dy)
    xh = x.detach().permute(0, 2, 3, 1).contiguous()
    stats = ops.norm_stats(xh, groups, n=n)
    fwd = ops.norm_apply(xh, stats, gamma.detach(), beta.detach(), act=act, groups=groups, eps=eps, n=n, terms=2)
    assert _rel(fwd.float().sum(0).permute(0, 3, 1, 2), yy.detach()) < 1e-5
    dgam, dbet = torch.zeros(Cc, device=cuda), torch.zeros(Cc, device=cuda)
    dx, dxp = ops.norm_bwd(xh, stats, gamma.detach(), beta.detach(), dy.permute(0, 2, 3, 1).contiguous(), act=act,
                           groups=groups, eps=eps, dgamma=dgam, dbeta=dbet, add=add, want_planes=True, n=n, terms=2)
    assert _rel(dx - add, x.grad.permute(0, 2, 3, 1)) < 1e-4
    assert _rel(dxp.float().sum(0), dx) < 1e-6
    assert _rel(dgam, gamma.grad) < 1e-4 and _rel(dbet, beta.grad) < 1e-4
    if groups_mode == "bn":
        rm, rv = torch.zeros(Cc, device=cuda), torch.ones(Cc, device=cuda)
        ops.bn_update_running(stats, rm, rv, N * H * W, 0.1)
        rm2, rv2 = torch.zeros(Cc, device=cuda), torch.ones(Cc, device=cuda)
        F.batch_norm(x.detach(), rm2, rv2, None, None, True, 0.1, 1e-5)
        assert _rel(rm, rm2) < 1e-5 and _rel(rv, rv2) < 1e-5


def test_small_training_kernels(cuda):
    from oracle import vqgan_train_ref as TR
    from text2human_b200 import ops
    g = torch.Generator().manual_seed(9)
    # nearest-x2 adjoint
    src = torch.randn(2, 16, 8, 4, generator=g).to(cuda).requires_grad_(True)        # NCHW
    upg = torch.randn(2, 16, 16, 8, generator=g).to(cuda)
    F.interpolate(src, scale_factor=2.0, mode="nearest").backward(upg)
    y = ops.sumpool2(upg.permute(0, 2, 3, 1).contiguous())                            # NHWC [2,8,4,16]
    assert _rel(y.permute(0, 3, 1, 2), src.grad) < 1e-6
    # planes s2d
    a = ops.split_planes(torch.randn(2, 8, 4, 16, generator=g).to(cuda), 2)
    s = ops.planes_s2d(a)
    for ph in range(4):
        assert torch.equal(s[:, ph], a[:, :, ph >> 1::2, ph & 1::2])
    # L1 / hinge
    xa, xb = torch.randn(2, 3, 16, 8, generator=g).to(cuda), torch.randn(2, 3, 16, 8, generator=g).to(cuda)
    acc = torch.zeros(1, dtype=torch.float64, device=cuda)
    gr = ops.l1_loss(xa, xb, acc, gscale=2.0)
    assert abs(acc.item() - (xa - xb).abs().double().sum().item()) < 1e-6 * acc.item()
    assert torch.equal(gr, 2.0 * torch.sign(xb - xa))
    lg = torch.randn(2, 7, 5, 1, generator=g).to(cuda) * 2
    for sgn in (1.0, -1.0, 0.0):
        acc.zero_()
        gg = ops.hinge_loss(lg, acc, sgn, gscale=0.25)
        if sgn == 0.0:
            assert abs(acc.item() - lg.double().sum().item()) < 1e-9 + 1e-6 * abs(lg.double().sum().item())
            assert torch.equal(gg, torch.full_like(lg, 0.25))
        else:
            assert abs(acc.item() - F.relu(1 - sgn * lg).double().sum().item()) < 1e-5
            assert torch.equal(gg, torch.where(1 - sgn * lg > 0, torch.full_like(lg, -sgn * 0.25), torch.zeros_like(lg)))
    # DiffAugment forward + backward against the torch restatement with the same draws
    B, H, W = 3, 16, 8
    img = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(cuda).requires_grad_(True)
    torch.manual_seed(77)
    want = TR.diff_augment(img)
    torch.manual_seed(77)
    r = torch.cat([torch.rand(B, 1, 1, 1, device=cuda).view(B, 1) for _ in range(3)], 1).contiguous()
    sx, sy = int(H * 0.125 + 0.5), int(W * 0.125 + 0.5)
    tx = torch.randint(-sx, sx + 1, size=[B, 1, 1], device=cuda).view(B, 1)
    ty = torch.randint(-sy, sy + 1, size=[B, 1, 1], device=cuda).view(B, 1)
    t = torch.cat([tx, ty], 1).int().contiguous()
    got = ops.diffaug_fwd(img.detach(), r, t)
    assert _rel(got, want.detach()) < 1e-5
    dout = torch.randn(B, 3, H, W, generator=g).to(cuda)
    want.backward(dout)
    assert _rel(ops.diffaug_bwd(dout, r, t), img.grad) < 1e-5
    # adaptive weight + axpy with a device scalar
    rg, gg2 = torch.randn(500, generator=g).to(cuda) * 8, torch.randn(500, generator=g).to(cuda) * 8
    wdev = torch.zeros(1, device=cuda)
    ops.adaptive_weight(rg, gg2, wdev, 1.0 / 8, 10.0, 1.0)
    wwant = (rg / 8).norm() / ((gg2 / 8).norm() + 1e-4)
    assert abs(wdev.item() - wwant.item()) < 1e-5 * wwant.item()
    ops.adaptive_weight(rg, gg2, wdev, 1.0 / 8, 0.5, 1.0)
    assert wdev.item() == 0.5
    assert _rel(ops.axpy_dev(rg, gg2, wdev), rg + 0.5 * gg2) < 1e-7
    # LeakyReLU backward from the output's sign
    pre = torch.randn(2, 4, 4, 8, generator=g).to(cuda)
    yp = ops.split_planes(F.leaky_relu(pre, 0.2), 2)
    dyy = torch.randn(2, 4, 4, 8, generator=g).to(cuda)
    dpre, dprep = ops.lrelu_bwd(yp, dyy)
    assert torch.equal(dpre, dyy * torch.where(pre > 0, 1.0, 0.2)) and _rel(dprep.float().sum(0), dpre) < 1e-6
    # quantizer backward vs autograd of the restatement
    cb = R.codebooks(3, 18, 16, 32, "trained").to(cuda)
    z = R.latent(4, (2, 32, 4, 2)).to(cuda).requires_grad_(True)
    mask = R.blocky_mask(5, 2, 64, 32, 16, extra_ids=(20,)).to(cuda)
    books = [cb[k].clone().requires_grad_(True) for k in range(18)]
    zq, loss = TR.quantize_texture_train(books, z, mask)
    dzq = torch.randn(zq.shape, generator=g).to(cuda)
    (3.0 * loss + (zq * dzq).sum()).backward()
    zh = z.detach().permute(0, 2, 3, 1).contiguous()
    ids = ops.mask_to_ids(mask, 4, 2)
    rr = ops.vq_search(zh, cb, ids, cont_stride=1024)
    dcb = torch.zeros_like(cb)
    nel = z.numel()
    dz = ops.vq_bwd(zh, cb, rr["idx"], ids, dzq.permute(0, 2, 3, 1).contiguous(), dcb, 3.0 * 2 / nel, 3.0 * 2 * 0.25 / nel)
    assert _rel(dz.permute(0, 3, 1, 2), z.grad) < 1e-5
    want_dcb = torch.stack([bk.grad if bk.grad is not None else torch.zeros_like(bk) for bk in books])
    assert _rel(dcb, want_dcb) < 1e-5


def test_sample_step_draws_the_categorical_law(cuda):
    """t2h_sample_step: revealed rows are exactly {u < 1/t and still masked}; tokens follow softmax(logits/temp)
    (chi-square over many independent draws); unrevealed rows are untouched."""
    from text2human_b200 import ops
    ncls, M = 16, 4096
    g = torch.Generator().manual_seed(1)
    row_logits = torch.randn(ncls, generator=g) * 1.5
    logits = row_logits.repeat(M, 1).to(cuda).contiguous()
    temp = 0.8
    p = torch.softmax(row_logits.double() / temp, 0)
    u = torch.rand(M, generator=g).to(cuda)
    tex = torch.randint(0, 18, (M,), generator=g).to(cuda)
    x_t = torch.full((M,), 18432, dtype=torch.long, device=cuda)
    unm = torch.zeros(M, dtype=torch.uint8, device=cuda)
    unm[::7] = 1
    before = unm.clone()
    ops.sample_step(logits, u, tex, x_t, unm, t=2, temp=temp, seed=1234, step=5, n_heads=18)
    rev = (u < 0.5) & (before == 0)
    assert torch.equal(unm.bool(), rev | before.bool())
    assert bool((x_t[~rev] == 18432).all())
    tok = x_t[rev] - 1024 * tex[rev]
    assert int(tok.min()) >= 0 and int(tok.max()) < ncls
    # pool draws from several seeds / steps for the chi-square
    counts = torch.zeros(ncls, dtype=torch.float64)
    total = 0
    for seed in range(12):
        x2 = torch.full((M,), 18432, dtype=torch.long, device=cuda)
        un2 = torch.zeros(M, dtype=torch.uint8, device=cuda)
        ops.sample_step(logits, torch.zeros(M, device=cuda), tex, x2, un2, t=1, temp=temp, seed=seed, step=seed,
                        n_heads=18)
        tk = (x2 - 1024 * tex).cpu()
        counts += torch.bincount(tk, minlength=ncls).double()
        total += M
    chi2 = float(((counts - total * p) ** 2 / (total * p)).sum())
    assert chi2 < 45.0, chi2     # 15 degrees of freedom: P(chi2 > 45) < 1e-4
    # same (seed, step) -> same draws; different step -> different draws
    xa = torch.full((M,), 18432, dtype=torch.long, device=cuda)
    xb, xc = xa.clone(), xa.clone()
    for xx, st in ((xa, 3), (xb, 3), (xc, 4)):
        ops.sample_step(logits, torch.zeros(M, device=cuda), tex, xx, torch.zeros(M, dtype=torch.uint8, device=cuda),
                        t=1, temp=temp, seed=9, step=st, n_heads=18)
    assert torch.equal(xa, xb) and not torch.equal(xa, xc)


@pytest.mark.parametrize("N,H,W,C,act", [(2, 64, 32, 128, "swish"), (1, 40, 24, 256, "swish"), (2, 32, 16, 128, "lrelu"),
                                         (8, 32, 16, 512, "swish"), (3, 128, 64, 256, "swish"), (2, 256, 128, 128, "swish")])
def test_norm_backward_sums_fused_into_the_data_gradient_conv(cuda, N, H, W, C, act):
    """t2h_tapgemm_params.nb_sums: pass 1 of the GroupNorm backward (sum du, sum du*xhat per image and channel)
    accumulated in the epilogue of the 3x3 data-gradient conv that produces dy, against the stand-alone reduce pass of
    t2h_norm_bwd and an fp64 restatement; the conv's own output must not change."""
    from text2human_b200 import conv_grad as G
    from text2human_b200 import ops
    ops.set_precision("fp32")
    g = torch.Generator(device=cuda).manual_seed(N * H + C)
    Co = 128
    x = torch.randn(N, H, W, C, device=cuda, generator=g) * 1.5 + 0.3        # the norm's input
    gamma = torch.randn(C, device=cuda, generator=g)
    beta = torch.randn(C, device=cuda, generator=g)
    w = torch.randn(Co, C, 3, 3, device=cuda, generator=g) / (9 * C) ** 0.5     # the conv that consumed act(norm(x))
    gy = torch.randn(N, H, W, Co, device=cuda, generator=g)                   # gradient w.r.t. the conv output
    st = ops.norm_stats(x, 32)
    wn, wt = G.weight_planes(G.oihw_to_master(w), 2)
    gyp = ops.f32_to_planes(gy, ops.CVT_PLAIN)
    nb = ops.nb_context(x, st, gamma, beta, act=act, groups=32, eps=1e-6)
    assert nb is not None
    da_f = G.dgrad("k3", gyp, wt, n=N, in_hw=(H, W), nb=nb)
    da_p = G.dgrad("k3", gyp, wt, n=N, in_hw=(H, W))
    assert torch.equal(da_f, da_p)
    # fp64 restatement of pass 1
    xd = x.double()
    xg = xd.view(N, H * W, 32, C // 32)
    mean = xg.mean(dim=(1, 3), keepdim=True)
    var = xg.var(dim=(1, 3), unbiased=False, keepdim=True)
    xh = ((xg - mean) / torch.sqrt(var + 1e-6)).view(N, H, W, C)
    u = xh * gamma.double() + beta.double()
    if act == "swish":
        s = torch.sigmoid(u)
        dact = s * (1 + u * (1 - s))
    else:
        dact = torch.where(u > 0, torch.ones_like(u), torch.full_like(u, 0.2))
    du = da_p.double() * dact
    ref = torch.stack((du.sum(dim=(1, 2)), (du * xh).sum(dim=(1, 2))), dim=-1)   # [N, C, 2]
    got = nb["sums"].view(N, C, 2)
    e = ((got - ref).abs().max() / ref.abs().max()).item()
    print(f"[nb] fused pass-1 sums vs fp64: {e:.2e}")
    assert e < 2e-5
    # the full backward with and without the fused sums
    dx_f, dxp_f = ops.norm_bwd(x, st, gamma, beta, da_f, act=act, groups=32, eps=1e-6, want_planes=True, sums=nb["sums"])
    dx_p, dxp_p = ops.norm_bwd(x, st, gamma, beta, da_p, act=act, groups=32, eps=1e-6, want_planes=True)
    assert ((dx_f - dx_p).abs().max() / dx_p.abs().max()).item() < 1e-5
