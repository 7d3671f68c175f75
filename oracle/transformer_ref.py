"""Plain-PyTorch fp32 restatement of the reference index-prediction transformer
and its absorbing-diffusion sampling loop.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates /root/reference/models/archs/transformer_arch.py (TransformerMultiHead
:184-273, Block :74-99, CausalSelfAttention :9-71 with causal=False) and
/root/reference/models/sample_model.py sample_fn (:256-328), functionally over a
flat reference-keyed ``state_dict``.
"""
import math
import re

import torch
import torch.nn.functional as F


def _n_blocks(sd):
    ids = {int(m.group(1)) for k in sd for m in [re.match(r"blocks\.(\d+)\.", k)] if m}
    return max(ids) + 1


def _n_heads_out(sd):
    ids = {int(m.group(1)) for k in sd for m in [re.match(r"head_list\.(\d+)\.", k)] if m}
    return max(ids) + 1


def attention(sd, p, x, n_head):
    # CausalSelfAttention.forward, causal=False                   transformer_arch.py:37-71
    B, T, C = x.shape
    hs = C // n_head
    k = F.linear(x, sd[p + "key.weight"], sd[p + "key.bias"]).view(B, T, n_head, hs).transpose(1, 2)
    q = F.linear(x, sd[p + "query.weight"], sd[p + "query.bias"]).view(B, T, n_head, hs).transpose(1, 2)
    v = F.linear(x, sd[p + "value.weight"], sd[p + "value.bias"]).view(B, T, n_head, hs).transpose(1, 2)
    att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(hs))
    att = F.softmax(att, dim=-1)
    y = (att @ v).transpose(1, 2).contiguous().view(B, T, C)
    return F.linear(y, sd[p + "proj.weight"], sd[p + "proj.bias"])


def block(sd, p, x, n_head):
    # Block.forward                                               transformer_arch.py:91-99
    C = x.shape[-1]
    x = x + attention(sd, p + "attn.", F.layer_norm(x, (C,), sd[p + "ln1.weight"], sd[p + "ln1.bias"]), n_head)
    h = F.layer_norm(x, (C,), sd[p + "ln2.weight"], sd[p + "ln2.bias"])
    h = F.linear(h, sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"])
    h = F.gelu(h)
    return x + F.linear(h, sd[p + "mlp.2.weight"], sd[p + "mlp.2.bias"])


def transformer_logits(sd, idx, segm_tokens, texture_tokens, n_head):
    """TransformerMultiHead.forward (:249-273) -> list of per-head logits [B,T,classes]."""
    T = idx.shape[1]
    x = sd["tok_emb.weight"][idx] + sd["pos_emb"][:, :T, :] + sd["segm_emb.weight"][segm_tokens] \
        + sd["texture_emb.weight"][texture_tokens]
    for i in range(_n_blocks(sd)):
        x = block(sd, f"blocks.{i}.", x, n_head)
    C = x.shape[-1]
    x = F.layer_norm(x, (C,), sd["ln_f.weight"], sd["ln_f.bias"])
    return [F.linear(x, sd[f"head_list.{i}.weight"]) for i in range(_n_heads_out(sd))]


def sample_fn(logits_fn, segm_tokens, texture_mask, latent_shape, mask_id, sample_steps, temp=1.0,
              trace=None, reveal_u=None):
    """BaseSampleModel.sample_fn (sample_model.py:256-328).

    logits_fn(x_t, segm_tokens, texture_tokens) -> list of 18 [B,T,1024] logits.
    Uses the global torch RNG exactly as the reference does: per step one torch.rand for the reveal
    mask, then one Categorical draw per codebook that has positions to reveal, ascending codebook order.
    If ``trace`` is a list, (x_t, changes) of every step are appended (for teacher-forced parity).
    ``reveal_u`` [steps, B, T] replaces the per-step torch.rand draws (step index 0 = the first, t = sample_steps),
    making the reveal schedule a pure function of its argument.

    Pinned: tests/test_sample_oracle.py reproduces, token for token, the fixture that oracle/make_golden_sample.py
    recorded from the REAL BaseSampleModel.sample_fn under the same global seed.
    """
    B = segm_tokens.shape[0]
    device = segm_tokens.device
    n = latent_shape[0] * latent_shape[1]
    x_t = torch.ones((B, n), device=device).long() * mask_id
    unmasked = torch.zeros_like(x_t).bool()
    texture_tokens = F.interpolate(texture_mask, tuple(latent_shape), mode="nearest").view(B, -1).long()
    tex_flat = texture_tokens.view(-1)
    out = [torch.full(tex_flat.size(), -1, dtype=torch.long, device=device) for _ in range(18)]
    for t in reversed(range(1, sample_steps + 1)):
        # the reference compares against 1 / t.float() computed in fp32 (:283-286)
        thr = 1 / torch.full((B,), t, device=device, dtype=torch.long).float().unsqueeze(-1)
        u = torch.rand(x_t.shape, device=device) if reveal_u is None else reveal_u[sample_steps - t]
        changes = u < thr
        changes = torch.bitwise_xor(changes, torch.bitwise_and(changes, unmasked))
        unmasked = torch.bitwise_or(unmasked, changes)
        if trace is not None:
            trace.append((x_t.clone(), changes.clone()))
        logits_list = logits_fn(x_t, segm_tokens, texture_tokens)
        ch = changes.view(-1)
        flat = x_t.view(-1).clone()
        for k, lg in enumerate(logits_list):
            if torch.sum(tex_flat[ch] == k) > 0:
                lg = lg / temp
                draw = torch.distributions.Categorical(logits=lg).sample().long().view(-1)
                sel = torch.bitwise_and(ch, tex_flat == k)
                flat[sel] = draw[sel] + 1024 * k
                out[k][sel] = draw[sel]
        x_t = flat.view(B, n)
    return [o.view(B, n) for o in out], x_t


def train_loss(sd, x_0, gt_list, segm_tokens, texture_tokens, t, mask, n_head, mask_id, num_timesteps=1000,
               loss_type="reweighted_elbo"):
    """TransformerTextureAwareModel._train_loss (transformer_model.py:232-271) with the random draws
    (t from sample_time :198-207, the mask from q_sample :212-230) passed in so that it is a pure function.
    x_0 [B,T] continual tokens; gt_list: per-texture ground-truth indices with -1 outside the texture.
    -> (loss, vb_loss) scalars with autograd history back to ``sd``."""
    x_t = x_0.clone()
    x_t[mask] = mask_id
    logits_list = transformer_logits(sd, x_t, segm_tokens, texture_tokens, n_head)
    ce = 0
    for lg, gt in zip(logits_list, gt_list):
        gt_ignore = gt.clone()
        gt_ignore[~mask] = -1
        ce = ce + F.cross_entropy(lg.permute(0, 2, 1), gt_ignore, ignore_index=-1, reduction="none").sum(1)
    pt = torch.ones_like(t).float() / num_timesteps
    ntok = x_0.shape[1:].numel()
    vb = ce / t / pt / (math.log(2) * ntok)
    if loss_type == "elbo":
        loss = vb
    elif loss_type == "reweighted_elbo":
        loss = (1 - (t / num_timesteps)) * ce / (math.log(2) * ntok)
    else:
        raise ValueError(loss_type)
    return loss.mean(), vb.mean()


def adam_update(p, g, m, v, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
    """one torch.optim.Adam step (weight_decay 0, amsgrad off) restated; returns new (p, m, v)"""
    m = betas[0] * m + (1 - betas[0]) * g
    v = betas[1] * v + (1 - betas[1]) * g * g
    bc1 = 1 - betas[0] ** step
    bc2 = 1 - betas[1] ** step
    p = p - (lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps)
    return p, m, v
