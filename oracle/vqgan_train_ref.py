"""Plain-PyTorch fp32 restatement of the VQGAN training step (SURVEY a18 / BASELINE config 5): generator loss with
the adaptive discriminator weight, hinge discriminator loss, DiffAugment, as a pure function of explicit state
dicts and the torch RNG.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The CUDA path for this step is NOT built yet; this
file and tests/golden/vqgan_train.npz pin what it will have to reproduce.

Restates /root/reference/models/vqgan_model.py VQImageSegmTextureModel.training_step (:438-488, same body as
VQModel.training_step :289-327) with forward_step (:532-551), /root/reference/models/losses/vqgan_loss.py
(calculate_adaptive_weight :5-12, adopt_weight :15-18, hinge_d_loss :21-26, DiffAugment 'color,translation'
:29-80), Discriminator (models/archs/vqgan_arch.py:1155-1203, BatchNorm in training mode) and the straight-through
/ legacy-beta codebook loss of VectorQuantizerTexture.forward (:270-281).

Third-party: `lpips.LPIPS(net="vgg")` (lpips==0.1.4, downloads VGG weights: unavailable offline).  It enters as
``perceptual(x, xrec) -> [B,1,1,1]``; the fixture uses the zero function (perceptual_weight therefore has no effect),
as BASELINE config 5 states ("LPIPS stubbed").
"""
import torch
import torch.nn.functional as F

from . import vqgan_ref as V


def quantize_texture_train(codebooks, z, segm_map, beta=0.25):
    """VectorQuantizerTexture.forward with its autograd semantics: nearest codes per texture, legacy loss
    mean((zq.detach()-z)^2) + beta*mean((zq-z.detach())^2), straight-through z + (zq - z).detach().
    codebooks: list/tensor of 18 [n_e, D] (requires_grad for the codebook gradient)."""
    b, c, hh, ww = z.shape
    seg = F.interpolate(segm_map, size=(hh, ww), mode="nearest").reshape(-1)
    zp = z.permute(0, 2, 3, 1).contiguous()
    rows = zp.view(-1, c)
    zq = torch.zeros_like(rows)
    for k in range(len(codebooks)):
        sel = seg == k
        if sel.any():
            e = codebooks[k]
            d = torch.sum(rows[sel] ** 2, dim=1, keepdim=True) + torch.sum(e ** 2, dim=1) - 2 * (rows[sel] @ e.t())
            zq = zq.index_put((sel.nonzero(as_tuple=True)[0],), e[torch.argmin(d, dim=1)])
    zq = zq.view(zp.shape)
    loss = torch.mean((zq.detach() - zp) ** 2) + beta * torch.mean((zq - zp.detach()) ** 2)
    zq = zp + (zq - zp).detach()
    return zq.permute(0, 3, 1, 2).contiguous(), loss


def discriminator(sd, x, prefix="main.", n_layers=3, train=True, bn_state=None):
    """Discriminator.forward (:1200-1203): conv4x4 s2 + LeakyReLU, (n_layers-1) x [conv4x4 s2, BN, LeakyReLU],
    [conv4x4 s1, BN, LeakyReLU], conv4x4 s1 -> 1 channel.  BatchNorm uses batch statistics when ``train``;
    running-statistic updates are returned in ``bn_state`` (dict) if given."""
    i = 0
    h = F.leaky_relu(F.conv2d(x, sd[f"{prefix}{i}.weight"], sd[f"{prefix}{i}.bias"], stride=2, padding=1), 0.2)
    i = 2
    for n in range(1, n_layers + 1):
        stride = 2 if n < n_layers else 1
        h = F.conv2d(h, sd[f"{prefix}{i}.weight"], None, stride=stride, padding=1)
        bn = f"{prefix}{i + 1}."
        rm, rv = sd[bn + "running_mean"].clone(), sd[bn + "running_var"].clone()
        h = F.batch_norm(h, rm, rv, sd[bn + "weight"], sd[bn + "bias"], train, 0.1, 1e-5)
        if bn_state is not None:
            bn_state[bn + "running_mean"], bn_state[bn + "running_var"] = rm, rv
        h = F.leaky_relu(h, 0.2)
        i += 3
    return F.conv2d(h, sd[f"{prefix}{i}.weight"], sd[f"{prefix}{i}.bias"], stride=1, padding=1)


def diff_augment(x):
    """DiffAugment(x, 'color,translation') with the reference's RNG call order: three torch.rand(B,1,1,1)
    (brightness, saturation, contrast) then two torch.randint translations"""
    b = x.size(0)
    x = x + (torch.rand(b, 1, 1, 1, dtype=x.dtype, device=x.device) - 0.5)
    m = x.mean(dim=1, keepdim=True)
    x = (x - m) * (torch.rand(b, 1, 1, 1, dtype=x.dtype, device=x.device) * 2) + m
    m = x.mean(dim=[1, 2, 3], keepdim=True)
    x = (x - m) * (torch.rand(b, 1, 1, 1, dtype=x.dtype, device=x.device) + 0.5) + m
    sx, sy = int(x.size(2) * 0.125 + 0.5), int(x.size(3) * 0.125 + 0.5)
    tx = torch.randint(-sx, sx + 1, size=[b, 1, 1], device=x.device)
    ty = torch.randint(-sy, sy + 1, size=[b, 1, 1], device=x.device)
    gb, gx, gy = torch.meshgrid(torch.arange(b, device=x.device), torch.arange(x.size(2), device=x.device),
                                torch.arange(x.size(3), device=x.device), indexing="ij")
    gx = torch.clamp(gx + tx + 1, 0, x.size(2) + 1)
    gy = torch.clamp(gy + ty + 1, 0, x.size(3) + 1)
    xp = F.pad(x, [1, 1, 1, 1, 0, 0, 0, 0])
    return xp.permute(0, 2, 3, 1).contiguous()[gb, gx, gy].permute(0, 3, 1, 2).contiguous()


def training_step(sd, codebooks, sd_disc, x, mask, step, *, perceptual=None, perceptual_weight=1.0,
                  disc_start_step=0, disc_weight_max=1.0, diff_aug=True, disc_layers=3):
    """-> dict(loss, d_loss or None, nll_loss, g_loss, d_weight, codebook_loss, xrec).  ``sd`` / ``codebooks`` /
    ``sd_disc`` tensors that require grad receive gradients from loss.backward() / d_loss.backward() exactly as the
    reference's two optimisers see them (the discriminator's parameters also collect the generator loss's
    gradient in the reference, until disc_optimizer.zero_grad() clears it: not restated)."""
    h = V.conv(sd, "quant_conv", V.encoder(sd, x, "encoder."), padding=0)
    quant, codebook_loss = quantize_texture_train(codebooks, h, mask)
    xrec = V.decoder(sd, V.conv(sd, "post_quant_conv", quant, padding=0), "decoder.")
    recon = torch.abs(x - xrec)
    p_loss = perceptual(x, xrec) if perceptual is not None else torch.zeros(x.size(0), 1, 1, 1)
    nll_loss = torch.mean(recon + perceptual_weight * p_loss)
    xr = diff_augment(xrec) if diff_aug else xrec
    g_loss = -torch.mean(discriminator(sd_disc, xr, n_layers=disc_layers))
    last = sd["decoder.conv_out.weight"]
    rg = torch.autograd.grad(nll_loss, last, retain_graph=True)[0]
    gg = torch.autograd.grad(g_loss, last, retain_graph=True)[0]
    d_weight = torch.clamp(torch.norm(rg) / (torch.norm(gg) + 1e-4), 0.0, disc_weight_max).detach()
    d_weight = d_weight * (1 if step >= disc_start_step else 0.0)
    loss = nll_loss + d_weight * g_loss + codebook_loss
    d_loss = None
    if step > disc_start_step:
        real_in = diff_augment(x.detach()) if diff_aug else x.detach()
        lr_ = discriminator(sd_disc, real_in, n_layers=disc_layers)
        lf_ = discriminator(sd_disc, xr.detach(), n_layers=disc_layers)
        d_loss = 0.5 * (torch.mean(F.relu(1.0 - lr_)) + torch.mean(F.relu(1.0 + lf_)))
    return dict(loss=loss, d_loss=d_loss, nll_loss=nll_loss, g_loss=g_loss, d_weight=d_weight,
                codebook_loss=codebook_loss, xrec=xrec)
