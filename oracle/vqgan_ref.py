"""Plain-PyTorch fp32 restatement of the reference VQGAN forward path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The product package never
imports this file.

The reference (/root/reference/models/archs/vqgan_arch.py) builds nn.Module
trees; this restatement is *functional and structure-driven*: each function takes
a flat ``state_dict`` (reference key names) plus a key prefix, and reads the
network shape (levels, blocks, attention, shortcuts, down/up-sampling) off the
keys that exist.  It runs on whatever device the tensors live on, always in fp32
with TF32 disabled by the caller.

Pinned against the real reference modules by tests/test_oracle_golden.py using
fixtures produced by oracle/make_golden.py (which imports /root/reference).
"""
import re

import torch
import torch.nn.functional as F


def _sub(sd, prefix):
    """keys under ``prefix`` (without it)"""
    n = len(prefix)
    return [k[n:] for k in sd if k.startswith(prefix)]


def _count(sd, prefix):
    """number of consecutive integer children ``prefix0.``, ``prefix1.``, ..."""
    ids = set()
    for k in _sub(sd, prefix):
        m = re.match(r"(\d+)\.", k)
        if m:
            ids.add(int(m.group(1)))
    return (max(ids) + 1) if ids else 0


def conv(sd, name, x, stride=1, padding=1):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def norm(sd, name, x):
    # Normalize(): GroupNorm(32, C, eps=1e-6, affine)            vqgan_arch.py:515-517
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)


def swish(x):
    # nonlinearity()                                             vqgan_arch.py:510-512
    return x * torch.sigmoid(x)


def resnet_block(sd, p, x):
    # ResnetBlock.forward (temb is None, dropout p=0)             vqgan_arch.py:597-617
    h = conv(sd, p + "conv1", swish(norm(sd, p + "norm1", x)))
    h = conv(sd, p + "conv2", swish(norm(sd, p + "norm2", h)))
    if (p + "nin_shortcut.weight") in sd:
        x = conv(sd, p + "nin_shortcut", x, padding=0)
    elif (p + "conv_shortcut.weight") in sd:
        x = conv(sd, p + "conv_shortcut", x)
    return x + h


def attn_block(sd, p, x):
    # AttnBlock.forward                                           vqgan_arch.py:636-661
    b, c, hh, ww = x.shape
    hn = norm(sd, p + "norm", x)
    q = conv(sd, p + "q", hn, padding=0).reshape(b, c, hh * ww)
    k = conv(sd, p + "k", hn, padding=0).reshape(b, c, hh * ww)
    v = conv(sd, p + "v", hn, padding=0).reshape(b, c, hh * ww)
    w = torch.bmm(q.transpose(1, 2), k) * (int(c) ** (-0.5))  # [b, i, j]
    w = F.softmax(w, dim=2)
    o = torch.bmm(v, w.transpose(1, 2)).reshape(b, c, hh, ww)  # o[b,c,i] = sum_j v[b,c,j] w[b,i,j]
    return x + conv(sd, p + "proj_out", o, padding=0)


def _mid(sd, p, h):
    h = resnet_block(sd, p + "mid.block_1.", h)
    h = attn_block(sd, p + "mid.attn_1.", h)
    return resnet_block(sd, p + "mid.block_2.", h)


def encoder(sd, x, p=""):
    # Encoder.forward                                             vqgan_arch.py:892-919
    h = conv(sd, p + "conv_in", x)
    levels = _count(sd, p + "down.")
    for lv in range(levels):
        lp = f"{p}down.{lv}."
        for bk in range(_count(sd, lp + "block.")):
            h = resnet_block(sd, f"{lp}block.{bk}.", h)
            if (f"{lp}attn.{bk}.norm.weight") in sd:
                h = attn_block(sd, f"{lp}attn.{bk}.", h)
        if (lp + "downsample.conv.weight") in sd:
            # Downsample: pad (0,1,0,1) then stride-2 conv          vqgan_arch.py:547-551
            h = conv(sd, lp + "downsample.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = _mid(sd, p, h)
    return conv(sd, p + "conv_out", swish(norm(sd, p + "norm_out", h)))


def decoder(sd, z, p="", bot_h=None, give_pre_end=False, stop_after_level=None, mid_h=None):
    # Decoder.forward / get_feature_top / get_feature_middle      vqgan_arch.py:1000-1087
    h = conv(sd, p + "conv_in", z)
    h = _mid(sd, p, h)
    levels = _count(sd, p + "up.")
    for lv in reversed(range(levels)):
        lp = f"{p}up.{lv}."
        for bk in range(_count(sd, lp + "block.")):
            h = resnet_block(sd, f"{lp}block.{bk}.", h)
            if (f"{lp}attn.{bk}.norm.weight") in sd:
                h = attn_block(sd, f"{lp}attn.{bk}.", h)
        if (lp + "upsample.conv.weight") in sd:
            # Upsample: nearest x2 then conv                        vqgan_arch.py:529-534
            h = conv(sd, lp + "upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
        if lv == 4 and bot_h is not None:
            h = h + bot_h                                          # :1023-1024
        if lv == 4 and mid_h is not None:
            h = h + mid_h                                          # :1084-1085
        if stop_after_level is not None and lv == stop_after_level:
            return h
    if give_pre_end:
        return h
    return conv(sd, p + "conv_out", swish(norm(sd, p + "norm_out", h)))


def decoder_res(sd, z, p=""):
    # DecoderRes.forward                                          vqgan_arch.py:1136-1151
    return _mid(sd, p, conv(sd, p + "conv_in", z))


# ----------------------------------------------------------------------------
# quantizers, written with the same torch expression the reference uses (this is
# the "torch-order" variant; oracle/vq_oracle.c is the fixed-order variant the
# CUDA kernel reproduces bit for bit)
# ----------------------------------------------------------------------------
def _nearest(z_rows, e):
    d = torch.sum(z_rows ** 2, dim=1, keepdim=True) + torch.sum(e ** 2, dim=1) - 2 * (z_rows @ e.t())
    return torch.argmin(d, dim=1)


def quantize_plain(codebook, z, beta=0.25):
    # VectorQuantizer.forward                                     vqgan_arch.py:79-122
    b, c, hh, ww = z.shape
    zf = z.permute(0, 2, 3, 1).reshape(-1, c)
    idx = _nearest(zf, codebook)
    zq = codebook[idx].view(b, hh, ww, c)
    zp = z.permute(0, 2, 3, 1)
    loss = torch.mean((zq - zp) ** 2) + beta * torch.mean((zq - zp) ** 2)
    zq = (zp + (zq - zp)).permute(0, 3, 1, 2).contiguous()
    return zq, loss, idx.view(b, hh, ww)


def quantize_texture(codebooks, z, segm_map, ps=1, cont_stride=1024, beta=0.25):
    """VectorQuantizerTexture.forward (ps=1, :212-287) and
    VectorQuantizerSpatialTextureAware.forward (ps=2, :375-461).
    codebooks: [18, n_e, D]."""
    b, c, hh, ww = z.shape
    hp, wp = hh // ps, ww // ps
    seg = F.interpolate(segm_map, size=(hp, wp), mode="nearest").reshape(-1)
    if ps == 1:
        rows = z.permute(0, 2, 3, 1).reshape(-1, c)
    else:
        rows = F.unfold(z, (ps, ps), stride=ps).permute(0, 2, 1).reshape(-1, c * ps * ps)
    zq = torch.zeros_like(rows)
    cont = torch.full((rows.shape[0],), -1, dtype=torch.long, device=z.device)
    per_book = []
    for k in range(codebooks.shape[0]):
        mine = torch.full((rows.shape[0],), -1, dtype=torch.long, device=z.device)
        sel = seg == k
        if sel.any():
            idx = _nearest(rows[sel], codebooks[k])
            zq[sel] = codebooks[k][idx]
            mine[sel] = idx
            cont[sel] = idx + cont_stride * k
        per_book.append(mine.view(b, hp, wp))
    if ps == 1:
        zq = zq.view(b, hh, ww, c).permute(0, 3, 1, 2)
    else:
        zq = F.fold(zq.view(b, hp * wp, -1).permute(0, 2, 1), (hh, ww), kernel_size=(ps, ps), stride=ps)
    loss = torch.mean((zq - z) ** 2) + beta * torch.mean((zq - z) ** 2)
    zq = (z + (zq - z)).contiguous()
    # the top quantizer reshapes the continual indices to [b,h,w] (:267); the spatial one does not (:460)
    return zq, loss, (cont.view(b, hp, wp) if ps == 1 else cont), per_book


def codebook_entry_texture(codebooks, indices_list, segm_map, shape, ps=1):
    # get_codebook_entry                                          vqgan_arch.py:289-309, :463-486
    b, hp, wp = shape[0], shape[1], shape[2]
    seg = F.interpolate(segm_map, size=(hp, wp), mode="nearest").reshape(-1)
    d = codebooks.shape[2]
    zq = torch.zeros(b * hp * wp, d, device=codebooks.device)
    for k in range(codebooks.shape[0]):
        sel = seg == k
        if sel.any():
            zq[sel] = codebooks[k][indices_list[k].reshape(-1)[sel]]
    if ps == 1:
        return zq.view(b, hp, wp, d).permute(0, 3, 1, 2).contiguous()
    return F.fold(zq.view(b, hp * wp, d).permute(0, 2, 1), (hp * ps, wp * ps), kernel_size=(ps, ps), stride=ps)


# ----------------------------------------------------------------------------
# wrapper-level sequences
# ----------------------------------------------------------------------------
def vq_forward_step(sd, codebooks, x, mask):
    """VQImageSegmTextureModel.forward_step (models/vqgan_model.py:532-551).
    sd keys: encoder.*, decoder.*, quant_conv.*, post_quant_conv.*"""
    h = encoder(sd, x, "encoder.")
    h = conv(sd, "quant_conv", h, padding=0)
    quant, emb_loss, cont, per_book = quantize_texture(codebooks, h, mask)
    dec = decoder(sd, conv(sd, "post_quant_conv", quant, padding=0), "decoder.")
    return dict(dec=dec, loss=emb_loss, idx_cont=cont, idx_list=per_book, z=h, quant=quant)


def hierarchy_forward_step(sd, top_codebooks, bot_codebooks, x, mask):
    """HierarchyVQSpatialTextureAwareModel.forward_step (models/hierarchy_vqgan_model.py:215-239).
    sd keys: top_encoder.*, top_quant_conv.*, top_post_quant_conv.*, bot_encoder.*, bot_quant_conv.*,
    bot_post_quant_conv.*, bot_decoder_res.*, decoder.*"""
    h = conv(sd, "top_quant_conv", encoder(sd, x, "top_encoder."), padding=0)
    quant_top, _, top_cont, _ = quantize_texture(top_codebooks, h, mask)
    quant_top = conv(sd, "top_post_quant_conv", quant_top, padding=0)
    hb = conv(sd, "bot_quant_conv", encoder(sd, x, "bot_encoder."), padding=0)
    quant_bot, emb_loss, bot_cont, _ = quantize_texture(bot_codebooks, hb, mask, ps=2,
                                                        cont_stride=bot_codebooks.shape[1])
    bot_dec_res = decoder_res(sd, conv(sd, "bot_post_quant_conv", quant_bot, padding=0), "bot_decoder_res.")
    dec = decoder(sd, quant_top, "decoder.", bot_h=bot_dec_res)
    return dict(dec=dec, loss=emb_loss, top_idx=top_cont, bot_idx=bot_cont, z_top=h, z_bot=hb)
