"""Generate tests/golden/data_prep.npz by running the REAL reference dataset item code
(/root/reference/data/segm_attr_dataset.py DeepFashionAttrSegmDataset.__getitem__ :120-164, unbound, on a stand-in
``self`` whose three loaders return synthetic arrays instead of reading files): the texture mask built from the parsing
map and the fused clothes attributes, and the image normalisation image / 127.5 - 1.
Run in the build container only:  python oracle/make_golden_dataprep.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    spec = importlib.util.spec_from_file_location("ref_segm_attr_dataset", "/root/reference/data/segm_attr_dataset.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cls = mod.DeepFashionAttrSegmDataset
    imgs, segms, attrs = R.dataset_items(91, 6, 32, 16)
    fake = types.SimpleNamespace(xflip=False, upper_cls=[1., 4.], lower_cls=[3., 5., 21.], outer_cls=[2.],
                                 upper_fused_attrs=attrs[:, 0].tolist(), lower_fused_attrs=attrs[:, 1].tolist(),
                                 outer_fused_attrs=attrs[:, 2].tolist(), _image_fnames=[f"{i}.png" for i in range(6)])
    fake._load_raw_image = lambda i: imgs[i].transpose(2, 0, 1).astype(np.float32)      # CHW float, as :74-88
    fake._load_densepose = lambda i: np.zeros((3, 32, 16), np.float32)
    fake._load_segm = lambda i: segms[i][np.newaxis].astype(np.float32)
    masks, images = [], []
    for i in range(6):
        item = cls.__getitem__(fake, i)
        masks.append(item["texture_mask"].numpy())
        images.append(item["image"].numpy())
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "data_prep.npz")
    np.savez_compressed(path, mask=np.stack(masks), image=np.stack(images))
    print(path, os.path.getsize(path), np.unique(np.stack(masks)))


if __name__ == "__main__":
    main()
