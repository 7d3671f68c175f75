"""GPU: dataset-side preparation kernels (data/segm_attr_dataset.py:138-154) bit-exact against the fixture
recorded from the real reference dataset item code (oracle/make_golden_dataprep.py)."""
import os

import numpy as np
import pytest
import torch

import golden_recipes as R

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "data_prep.npz")


def test_texture_mask_and_image_normalisation_match_reference(cuda):
    from text2human_b200 import ops
    g = np.load(GOLD)
    imgs, segms, attrs = R.dataset_items(91, 6, 32, 16)
    segm = torch.from_numpy(segms[:, None].astype(np.float32)).to(cuda)
    mask = ops.texture_mask(segm, torch.from_numpy(attrs).to(cuda))
    assert np.array_equal(mask.cpu().numpy(), g["mask"])
    planes, nchw = ops.u8_to_planes(torch.from_numpy(imgs).to(cuda), want_nchw=True, terms=2)
    assert np.array_equal(nchw.cpu().numpy(), g["image"])
    rebuilt = planes.float().sum(0)[..., :3].permute(0, 3, 1, 2)            # hi + lo reproduces the fp32 value
    assert float((rebuilt.cpu() - torch.from_numpy(g["image"])).abs().max()) < 1e-6
    assert float(planes[..., 3:].abs().max()) == 0.0
    # the planes are exactly what the encoder's entry conversion makes from the normalised NCHW image
    assert torch.equal(planes, ops.nchw_to_planes(nchw, terms=2))
