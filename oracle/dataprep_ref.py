"""numpy restatement of the reference's dataset-side preparation (data/segm_attr_dataset.py:138-154).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned to the real `DeepFashionAttrSegmDataset.__getitem__` by
tests/test_dataprep_oracle.py (fixture from oracle/make_golden_dataprep.py)."""
import numpy as np

UPPER, LOWER, OUTER = (1., 4.), (3., 5., 21.), (2.,)


def texture_mask(segm, attrs):
    """segm float [B,1,H,W]; attrs int [B,3] (upper, lower, outer fused attribute, 17 = none) -> mask float [B,1,H,W]"""
    segm = np.asarray(segm, np.float32)
    mask = np.zeros_like(segm)
    for b in range(segm.shape[0]):
        for group, a in zip((UPPER, LOWER, OUTER), attrs[b]):
            if a != 17:
                for c in group:
                    mask[b][segm[b] == c] = a + 1
    return mask


def normalize_image(img_u8_hwc):
    """uint8 [B,H,W,3] -> float32 NCHW, image / 127.5 - 1 (:154; float32 division then subtraction)"""
    x = np.asarray(img_u8_hwc).transpose(0, 3, 1, 2).astype(np.float32)
    return (x / np.float32(127.5) - np.float32(1.0)).astype(np.float32)
