"""Generate tests/golden/sample_fn.npz by running the REAL reference sampling loop
(/root/reference/models/sample_model.py BaseSampleModel.sample_fn :256-328, unbound, on a stand-in ``self`` that
carries only the attributes it reads) around the REAL reference TransformerMultiHead on the CPU with a fixed seed.

The fixture stores the 18 returned index maps; the restatement oracle/transformer_ref.sample_fn, driven by the same
global torch RNG seed, must reproduce them exactly (tests/test_sample_oracle.py).
Run in the build container only:  python oracle/make_golden_sample.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R  # noqa: E402
from oracle import ref_loader as RL  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    torch.set_num_threads(8)
    ns = RL.install("reference", wrappers=("sample_model",))
    cfg = R.SAMPLE_TRANSFORMER
    net = ns.transformer_arch.TransformerMultiHead(**cfg)
    net.load_state_dict(R.fill_state_dict(R.spec_of(net), 81), strict=True)
    B, steps = R.SAMPLE_BATCH, R.SAMPLE_STEPS
    segm_tokens, texture_mask = R.sample_inputs(82, B)
    fake = types.SimpleNamespace(batch_size=B, shape=(32, 16), device=torch.device("cpu"),
                                 mask_id=cfg["codebook_size"], texture_mask=texture_mask,
                                 segm_tokens=segm_tokens, sampler_fn=net)
    torch.manual_seed(83)
    with torch.no_grad():
        out = ns.sample_model.BaseSampleModel.sample_fn(fake, temp=1.0, sample_steps=steps)
    arr = torch.stack(out).numpy().astype(np.int16)          # [18, B, 512]
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "sample_fn.npz")
    np.savez_compressed(path, lists=arr)
    print(path, os.path.getsize(path), "revealed", int((arr >= 0).sum()), "of", B * 512)


if __name__ == "__main__":
    main()
