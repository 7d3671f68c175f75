"""Generate tests/golden/*.npz by running the REAL reference modules
(/root/reference/models/archs/{vqgan,transformer}_arch.py, loaded by file path because
`import models` needs mmcv/lpips) on the deterministic recipes in tests/golden_recipes.py.

Run in the build container only (the reference does not exist on the GPU box):
    python oracle/make_golden.py
The fixtures store reference OUTPUTS only; inputs and weights are re-created from the recipes.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_recipes as R  # noqa: E402

REF = "/root/reference/models/archs"
OUT = os.path.join(ROOT, "tests", "golden")


def _load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _loaded(module, seed):
    sd = R.fill_state_dict(R.spec_of(module), seed)
    module.load_state_dict(sd, strict=True)
    return module.eval()


def _set_books(q, cb):
    for k, emb in enumerate(q.embedding_list):
        emb.weight.data.copy_(cb[k])


def main():
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    vq = _load("vqgan_arch")
    tr = _load("transformer_arch")
    with torch.no_grad():
        # ---------------- quantizers (indices: the bit-exact contract) -------------
        out = {}
        for kind in ("default", "trained"):
            # top: VectorQuantizerTexture(n_e=1024 in the real config; 128 keeps the recipe light)
            cb = R.codebooks(11, 18, 128, 256, kind)
            q = vq.VectorQuantizerTexture(128, 256, beta=0.25)
            _set_books(q, cb)
            zscale = 1.0 if kind == "trained" else 0.02
            z = R.latent(12, (2, 256, 32, 16), zscale)
            for mname, mask in (("blocky", R.blocky_mask(13, 2, 512, 256, 64, extra_ids=(20,))),
                                ("iid", R.iid_mask(14, 2, 512, 256))):
                zq, loss, (_, cont, lst) = q(z, mask)
                tag = f"top_{kind}_{mname}"
                out[tag + "_cont"] = cont.numpy().astype(np.int32)
                out[tag + "_list"] = torch.stack(lst).numpy().astype(np.int32)
                out[tag + "_loss"] = np.float64(loss.item())
                out[tag + "_zq_sum"] = np.float64(zq.double().sum().item())
                out[tag + "_zq_abs"] = np.float64(zq.double().abs().sum().item())
            # entry lookup
            ent = q.get_codebook_entry(lst, mask, (2, 32, 16, 256))
            out[f"top_{kind}_entry_abs"] = np.float64(ent.double().abs().sum().item())
            # bottom: VectorQuantizerSpatialTextureAware on 2x2 patches
            cbb = R.codebooks(21, 18, 64, 32 * 4, kind)
            qb = vq.VectorQuantizerSpatialTextureAware(64, 32, beta=0.25, spatial_size=2)
            _set_books(qb, cbb)
            zb = R.latent(22, (2, 32, 32, 16), zscale)
            maskb = R.blocky_mask(23, 2, 256, 128, 32)
            zqb, lossb, (_, contb, lstb) = qb(zb, maskb)
            out[f"bot_{kind}_cont"] = contb.numpy().astype(np.int32)
            out[f"bot_{kind}_list"] = torch.stack(lstb).numpy().astype(np.int32)
            out[f"bot_{kind}_loss"] = np.float64(lossb.item())
            out[f"bot_{kind}_zq"] = zqb.numpy()
            entb = qb.get_codebook_entry(lstb, maskb, (2, 16, 8, 32))
            out[f"bot_{kind}_entry"] = entb.numpy()
            # plain VectorQuantizer (segm tokeniser)
            cbs = R.codebooks(31, 1, 128, 32, kind)
            qs = vq.VectorQuantizer(128, 32, beta=0.25, sane_index_shape=True)
            qs.embedding.weight.data.copy_(cbs[0])
            zs = R.latent(32, (2, 32, 32, 16), zscale)
            zqs, losss, (_, _, idxs) = qs(zs)
            out[f"plain_{kind}_idx"] = idxs.numpy().astype(np.int32)
            out[f"plain_{kind}_loss"] = np.float64(losss.item())
        np.savez_compressed(os.path.join(OUT, "quantizers.npz"), **out)

        # ---------------- float modules -------------------------------------------
        out = {}
        enc = _loaded(vq.Encoder(**R.TINY_ENC), 41)
        x = R.image(42, 2, 3, 32, 16)
        out["enc_z"] = enc(x).numpy()
        dec = _loaded(vq.Decoder(**R.TINY_DEC), 43)
        z = R.latent(44, (2, 32, 4, 2))
        bot_h = R.latent(45, (2, 64, 8, 4), name="bot_h")
        out["dec_plain"] = dec(z).numpy()
        out["dec_both"] = dec(z, bot_h=bot_h.clone()).numpy()
        out["dec_feature_top"] = dec.get_feature_top(z).numpy()
        dres = _loaded(vq.DecoderRes(**R.TINY_DECRES), 46)
        zr = R.latent(47, (2, 32, 8, 4))
        out["decres"] = dres(zr).numpy()
        # single blocks at awkward shapes
        rb = _loaded(vq.ResnetBlock(in_channels=64, out_channels=128, temb_channels=0, dropout=0.0), 48)
        xb = R.latent(49, (2, 64, 16, 8))
        out["resblock"] = rb(xb, None).numpy()
        ab = _loaded(vq.AttnBlock(64), 50)
        out["attnblock"] = ab(xb).numpy()
        up = _loaded(vq.Upsample(64, True), 51)
        out["upsample"] = up(xb).numpy()
        dn = _loaded(vq.Downsample(64, True), 52)
        out["downsample"] = dn(xb).numpy()
        np.savez_compressed(os.path.join(OUT, "vqgan_modules.npz"), **out)

        # ---------------- transformer ----------------------------------------------
        out = {}
        tf = _loaded(tr.TransformerMultiHead(**R.TINY_TRANSFORMER), 61)
        g = R._gen(62, "tokens")
        idx = torch.randint(0, 18 * 16 + 1, (2, 32), generator=g)
        segm = torch.randint(0, 32, (2, 32), generator=g)
        tex = torch.randint(0, 18, (2, 32), generator=g)
        out["logits"] = torch.stack(tf(idx, segm, tex)).numpy()
        np.savez_compressed(os.path.join(OUT, "transformer.npz"), **out)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
