"""CPU: pins oracle/transformer_ref.sample_fn (the restatement of BaseSampleModel.sample_fn, sample_model.py:256-328)
to the fixture recorded from the REAL reference loop (oracle/make_golden_sample.py): same global seed, same RNG call
order (one torch.rand per step, then one Categorical draw per codebook that has positions to reveal, ascending) ->
the same tokens, exactly."""
import os

import numpy as np
import torch

import golden_recipes as R
from oracle import transformer_ref as TR

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sample_fn.npz")


def _net_sd():
    from text2human_b200.transformer_arch import TransformerMultiHead
    net = TransformerMultiHead(**R.SAMPLE_TRANSFORMER)
    return R.fill_state_dict(R.spec_of(net), 81)


def test_sample_fn_restatement_reproduces_reference_tokens():
    gold = np.load(GOLD)["lists"].astype(np.int64)
    sd = _net_sd()
    segm, mask = R.sample_inputs(82, R.SAMPLE_BATCH)
    torch.manual_seed(83)
    with torch.no_grad():
        out, x_t = TR.sample_fn(lambda x, s, t: TR.transformer_logits(sd, x, s, t, n_head=4), segm, mask, [32, 16],
                                18432, R.SAMPLE_STEPS)
    got = torch.stack(out).numpy()
    assert got.shape == gold.shape
    assert np.array_equal(got >= 0, gold >= 0), "reveal pattern differs from the reference loop"
    assert np.array_equal(got, gold), f"{int((got != gold).sum())} sampled tokens differ from the reference loop"
    tex = torch.nn.functional.interpolate(mask, (32, 16), mode="nearest").view(R.SAMPLE_BATCH, -1).long()
    assert bool(((x_t // 1024) == tex).all())


def test_sample_fn_reveal_schedule_is_a_function_of_the_uniforms():
    """with explicit reveal uniforms the set of positions revealed at each step does not depend on the drawn tokens"""
    sd = _net_sd()
    segm, mask = R.sample_inputs(82, R.SAMPLE_BATCH)
    g = torch.Generator().manual_seed(5)
    u = torch.rand((R.SAMPLE_STEPS, R.SAMPLE_BATCH, 512), generator=g)
    traces = []
    for seed in (1, 2):
        torch.manual_seed(seed)
        tr = []
        with torch.no_grad():
            TR.sample_fn(lambda x, s, t: TR.transformer_logits(sd, x, s, t, n_head=4), segm, mask, [32, 16], 18432,
                         R.SAMPLE_STEPS, trace=tr, reveal_u=u)
        traces.append(tr)
    for (xa, ca), (xb, cb) in zip(*traces):
        assert torch.equal(ca, cb)
    assert not all(torch.equal(a[0], b[0]) for a, b in zip(*traces))   # the tokens themselves do differ
