"""CPU: oracle/dataprep_ref (texture-mask construction + image normalisation) against the fixture recorded from the
real reference `DeepFashionAttrSegmDataset.__getitem__` (oracle/make_golden_dataprep.py)."""
import os

import numpy as np

import golden_recipes as R
from oracle import dataprep_ref as DR

GOLD = os.path.join(os.path.dirname(__file__), "golden", "data_prep.npz")


def test_dataprep_restatement_matches_reference_items():
    g = np.load(GOLD)
    imgs, segms, attrs = R.dataset_items(91, 6, 32, 16)
    mask = DR.texture_mask(segms[:, None].astype(np.float32), attrs)
    assert np.array_equal(mask, g["mask"])
    assert np.array_equal(DR.normalize_image(imgs), g["image"])
    assert set(np.unique(mask)) <= set(range(19))
