"""CPU: the host logic of the training-step convolutions (text2human_b200/conv_grad.py) -- tap tables of every conv
kind, the negated / parity tap sets of the data gradients, the tap-major master layout and its OIHW view -- checked by
EMULATING the kernels' contract in plain torch (out[n,h,w] = sum_i W[slot_i] . a[n + off_i, h + dy_i, w + dx_i], zero
outside) against torch.nn.functional.conv2d and autograd.  No libt2h compute call is made."""
import pytest
import torch
import torch.nn.functional as F

from text2human_b200 import conv_grad as G


def emulate_tap_conv(a_imgs, w_slots, taps, n, out_hw, tap_w=None):
    """a_imgs [I,h,w,C] (fp32 stand-in for the planes), w_slots [S,Cout,C] -> [n,H,W,Cout]"""
    H, W = out_hw
    I, ah, aw, C = a_imgs.shape
    out = torch.zeros(n, H, W, w_slots.shape[1], dtype=a_imgs.dtype)
    for i, (dy, dx, off) in enumerate(taps):
        wi = w_slots[tap_w[i] if tap_w is not None else i]
        for img in range(n):
            src = torch.zeros(H, W, C, dtype=a_imgs.dtype)
            hs = [h for h in range(H) if 0 <= h + dy < ah]
            ws = [w for w in range(W) if 0 <= w + dx < aw]
            if hs and ws:
                src[hs[0]:hs[-1] + 1, ws[0]:ws[-1] + 1] = a_imgs[img + off, hs[0] + dy:hs[-1] + dy + 1,
                                                                 ws[0] + dx:ws[-1] + dx + 1]
            out[img] += src @ wi.t()
    return out


def s2d(x_nhwc):
    """[N,H,W,C] -> phases [4*N,H/2,W/2,C] in the T2H_CVT_S2D order (phase p*2+q = x[2i+p, 2j+q])"""
    return torch.cat([x_nhwc[:, p::2, q::2] for p in (0, 1) for q in (0, 1)], 0)


def torch_conv(kind, x, w):
    if kind == "k3":
        return F.conv2d(x, w, padding=1)
    if kind == "k1":
        return F.conv2d(x, w)
    if kind == "down":
        return F.conv2d(F.pad(x, (0, 1, 0, 1)), w, stride=2)
    if kind == "k4s2":
        return F.conv2d(x, w, stride=2, padding=1)
    return F.conv2d(x, w, stride=1, padding=1)


@pytest.mark.parametrize("kind,H,W", [("k3", 6, 5), ("k1", 4, 4), ("down", 8, 6), ("k4s2", 8, 6), ("k4s1", 7, 5)])
def test_tap_tables_forward_dgrad_wgrad(kind, H, W):
    g = torch.Generator().manual_seed(1)
    N, Ci, Co, K = 2, 5, 4, G.KSIZE[kind]
    x = torch.randn(N, Ci, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Co, Ci, K, K, generator=g, dtype=torch.float64, requires_grad=True)
    y = torch_conv(kind, x, w)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    m = G.oihw_to_master(w.detach().float()).double()
    m[:, :Co, :Ci] = w.detach().permute(2, 3, 0, 1).reshape(K * K, Co, Ci)       # exact fp64 copy of the layout
    assert torch.equal(G.master_as_oihw(m, Co, Ci, K), w.detach())
    assert m.shape == (K * K, 8, 8) and float(m[:, Co:].abs().max()) == 0 and float(m[:, :, Ci:].abs().max()) == 0
    xh = F.pad(x.detach().permute(0, 2, 3, 1), (0, 8 - Ci))                       # NHWC, channels padded to 8
    a = s2d(xh) if G.STRIDE[kind] == 2 else xh
    Ho, Wo = G.out_hw(kind, H, W)
    assert y.shape[2:] == (Ho, Wo)
    # forward
    out = emulate_tap_conv(a, m, G.fwd_taps(kind, N), N, (Ho, Wo))
    assert torch.allclose(out[..., :Co].permute(0, 3, 1, 2), y.detach(), atol=1e-10)
    # weight gradient: dW[tap] = sum_pixels dy^T . shifted x
    dyh = F.pad(dy.permute(0, 2, 3, 1), (0, 8 - Co))
    gw = torch.zeros_like(m)
    for t, (ty, tx, off) in enumerate(G.fwd_taps(kind, N)):
        sh = emulate_tap_conv(a, torch.eye(8, dtype=torch.float64).unsqueeze(0), ((ty, tx, off),), N, (Ho, Wo))
        gw[t] = torch.einsum("nhwo,nhwi->oi", dyh, sh)
    assert torch.allclose(G.master_as_oihw(gw, Co, Ci, K), w.grad, atol=1e-10)
    # data gradient: transposed weights, negated taps (stride 1) or four parity launches (stride 2)
    mt = m.transpose(1, 2).contiguous()
    if G.STRIDE[kind] == 1:
        taps = tuple((-ty, -tx, 0) for ty, tx, _ in G.fwd_taps(kind, N))
        dx = emulate_tap_conv(dyh, mt, taps, N, (H, W))
    else:
        dx = torch.zeros(N, H, W, 8, dtype=torch.float64)
        seen = set()
        for pa in (0, 1):
            for pb in (0, 1):
                taps, slots = G.dgrad_parity_taps(kind, pa, pb)
                seen.update(slots)
                dx[:, pa::2, pb::2] = emulate_tap_conv(dyh, mt, taps, N, (H // 2, W // 2), tap_w=slots)
        assert seen == set(range(K * K))            # every tap feeds exactly one input parity
    assert torch.allclose(dx[..., :Ci].permute(0, 3, 1, 2), x.grad, atol=1e-10)
