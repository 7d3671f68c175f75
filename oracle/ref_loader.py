"""Load the UNMODIFIED reference sources (staged by oracle/make_ref.py into oracle/_ref/, or read from
/root/reference in the build container) as the `models.*` package the reference scripts import, with either

  archs="reference"  the reference's own arch modules (the oracle proper / the `--impl reference` bench arm), or
  archs="mirror"     text2human_b200's module mirrors installed AS `models.archs.vqgan_arch`, `.transformer_arch`,
                     `.unet_arch`, `.fcn_arch` — the drop-in boundary SURVEY.md §8b states: the reference's wrapper
                     files (`models/vqgan_model.py`, `hierarchy_vqgan_model.py`, `sample_model.py`,
                     `transformer_model.py`) then run unmodified on the B200 kernels.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).

Third-party packages the reference imports that are absent offline get stand-ins registered in sys.modules:
`lpips` (0.1.4; would download VGG weights) -> LPIPS returning zeros (BASELINE config 5: "LPIPS stubbed");
`mmcv` / `mmseg` (mmcv-full 1.2.1, mmsegmentation 0.9.0) -> the names unet_arch.py / fcn_arch.py import, with
`ConvModule` restating mmcv's published conv -> BN -> ReLU module (see oracle/make_golden_indexpred.py).
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = (os.path.join(ROOT, "oracle", "_ref"), "/root/reference")


def ref_root():
    """directory holding the reference's `models/` tree, or None"""
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "models", "archs", "vqgan_arch.py")):
            return c
    return None


def available():
    return ref_root() is not None


# ---------------------------------------------------------------- third-party stand-ins
class _ZeroLPIPS(nn.Module):
    def __init__(self, net="vgg", **kw):
        super().__init__()

    def forward(self, a, b):
        return torch.zeros(a.size(0), 1, 1, 1, device=a.device, dtype=a.dtype)


class ConvModule(nn.Module):
    """mmcv.cnn.ConvModule 1.2.1 for (conv_cfg=None, norm_cfg=BN, act_cfg=ReLU, order conv-norm-act)"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), **kw):
        super().__init__()
        assert conv_cfg is None
        with_norm = norm_cfg is not None
        if bias == "auto":
            bias = not with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.with_norm, self.with_act = with_norm, act_cfg is not None
        if with_norm:
            assert norm_cfg["type"] == "BN"
            self.bn = nn.BatchNorm2d(out_channels)
        if self.with_act:
            assert act_cfg["type"] == "ReLU"
            self.activate = nn.ReLU(inplace=True)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.bn(x)
        if self.with_act:
            x = self.activate(x)
        return x


class _Registry:
    def __init__(self):
        self.d = {}

    def register_module(self):
        def deco(cls):
            self.d[cls.__name__] = cls
            return cls
        return deco


def _third_party_stubs():
    reg = _Registry()

    def build_upsample_layer(cfg, *a, **kw):
        cfg = dict(cfg)
        return reg.d[cfg.pop("type")](*a, **kw, **cfg)

    noop = lambda *a, **k: None  # noqa: E731
    mods = {n: types.ModuleType(n) for n in ("lpips", "mmcv", "mmcv.cnn", "mmcv.runner", "mmcv.utils",
                                             "mmcv.utils.parrots_wrapper", "mmseg", "mmseg.utils", "mmseg.ops")}
    mods["lpips"].LPIPS = _ZeroLPIPS
    c = mods["mmcv.cnn"]
    c.UPSAMPLE_LAYERS, c.ConvModule = reg, ConvModule
    c.build_activation_layer = lambda cfg: nn.ReLU(inplace=True)
    c.build_norm_layer = lambda cfg, n: ("bn", nn.BatchNorm2d(n))
    c.build_upsample_layer = build_upsample_layer
    c.constant_init = c.kaiming_init = c.normal_init = noop
    mods["mmcv.runner"].load_checkpoint = noop
    mods["mmcv.utils.parrots_wrapper"]._BatchNorm = nn.modules.batchnorm._BatchNorm
    mods["mmseg.utils"].get_root_logger = noop
    mods["mmseg.ops"].resize = torch.nn.functional.interpolate
    for k, v in mods.items():
        sys.modules.setdefault(k, v)


def _exec(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def install(archs="reference", wrappers=("vqgan_model", "hierarchy_vqgan_model", "sample_model", "transformer_model")):
    """(Re)build the `models` package in sys.modules and return a namespace with the loaded modules:
    ns.vqgan_arch, ns.transformer_arch, ns.unet_arch, ns.fcn_arch, ns.vqgan_loss and one attribute per wrapper.
    Wrapper files are always the reference's own; ``archs`` picks what they find under models.archs."""
    root = ref_root()
    if root is None:
        raise RuntimeError("reference sources not staged: run `python oracle/make_ref.py` in the build container")
    assert archs in ("reference", "mirror")
    _third_party_stubs()
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    m = os.path.join(root, "models")
    for name in ("models", "models.archs", "models.losses"):
        pkg = types.ModuleType(name)
        pkg.__path__ = []
        sys.modules[name] = pkg
    ns = types.SimpleNamespace(root=root, archs=archs)
    if archs == "reference":
        for n in ("vqgan_arch", "transformer_arch", "unet_arch", "fcn_arch"):
            setattr(ns, n, _exec(f"models.archs.{n}", os.path.join(m, "archs", n + ".py")))
    else:
        import text2human_b200.index_pred_arch as ip
        import text2human_b200.transformer_arch as ta
        import text2human_b200.vqgan_arch as va
        unet = types.ModuleType("models.archs.unet_arch")
        unet.UNet = ip.UNet
        unet.ShapeUNet = None          # sample_from_pose only; not on the hot path
        fcn = types.ModuleType("models.archs.fcn_arch")
        fcn.MultiHeadFCNHead = ip.MultiHeadFCNHead
        fcn.FCNHead = None
        for n, mod in (("vqgan_arch", va), ("transformer_arch", ta), ("unet_arch", unet), ("fcn_arch", fcn)):
            sys.modules[f"models.archs.{n}"] = mod
            setattr(ns, n, mod)
    _exec("models.archs.shape_attr_embedding_arch", os.path.join(m, "archs", "shape_attr_embedding_arch.py"))
    for n in ("accuracy", "cross_entropy_loss", "segmentation_loss", "vqgan_loss"):
        p = os.path.join(m, "losses", n + ".py")
        if os.path.exists(p):
            try:
                setattr(ns, n, _exec(f"models.losses.{n}", p))
            except ImportError:        # cross_entropy_loss imports mmseg registries that are not on the hot path
                sys.modules.pop(f"models.losses.{n}", None)
    for w in wrappers:
        setattr(ns, w, _exec(f"models.{w}", os.path.join(m, w + ".py")))
    return ns


# ---------------------------------------------------------------- wrapper stand-ins
def bind(cls, obj, names):
    """bind the reference class's (unmodified) methods to a stand-in object"""
    for n in names:
        setattr(obj, n, types.MethodType(getattr(cls, n), obj))
    return obj


def vq_top_wrapper(ns, opt, device, with_disc=False, ndf=64, disc_layers=3):
    """A stand-in `self` carrying exactly the attributes VQImageSegmTextureModel's methods read, with the modules
    built by the SAME constructor calls as its __init__ (models/vqgan_model.py:391-422) -- __init__ itself hard-codes
    torch.device('cuda') and is therefore not usable on the CPU arm.  The methods (feed_data, encode, decode,
    forward_step, training_step, optimize_parameters) are the reference's own, unmodified."""
    va = ns.vqgan_arch
    import contextlib
    import io
    w = types.SimpleNamespace(opt=opt, device=torch.device(device))
    with contextlib.redirect_stdout(io.StringIO()):
        w.encoder = va.Encoder(ch=opt['ch'], num_res_blocks=opt['num_res_blocks'],
                               attn_resolutions=opt['attn_resolutions'], ch_mult=opt['ch_mult'],
                               in_channels=opt['in_channels'], resolution=opt['resolution'],
                               z_channels=opt['z_channels'], double_z=opt['double_z'], dropout=opt['dropout'])
        w.decoder = va.Decoder(in_channels=opt['in_channels'], resolution=opt['resolution'],
                               z_channels=opt['z_channels'], ch=opt['ch'], out_ch=opt['out_ch'],
                               num_res_blocks=opt['num_res_blocks'], attn_resolutions=opt['attn_resolutions'],
                               ch_mult=opt['ch_mult'], dropout=opt['dropout'], resamp_with_conv=True,
                               give_pre_end=False)
    w.quantize = va.VectorQuantizerTexture(opt['n_embed'], opt['embed_dim'], beta=0.25)
    w.quant_conv = torch.nn.Conv2d(opt["z_channels"], opt['embed_dim'], 1)
    w.post_quant_conv = torch.nn.Conv2d(opt['embed_dim'], opt["z_channels"], 1)
    mods = ["encoder", "decoder", "quantize", "quant_conv", "post_quant_conv"]
    if with_disc:
        w.disc = va.Discriminator(3, ndf, n_layers=disc_layers)
        w.perceptual = _ZeroLPIPS()
        w.perceptual_weight, w.disc_start_step, w.disc_weight_max = 1.0, opt.get('disc_start_step', 0), 1.0
        w.diff_aug, w.policy, w.log_dict = True, "color,translation", {}
        mods.append("disc")
    for n in mods:
        setattr(w, n, getattr(w, n).to(device))
    w.modules = mods
    cls = ns.vqgan_model.VQImageSegmTextureModel
    names = ["feed_data", "encode", "decode", "forward_step"]
    if with_disc:
        names += ["training_step", "optimize_parameters", "configure_optimizers"]
    return bind(cls, w, names)
