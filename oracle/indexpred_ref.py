"""Plain-PyTorch fp32 restatement of the index-prediction path (top features -> bottom-level indices):
UNet guidance encoder + MultiHeadFCNHead + the per-texture argmax of bot_index_prediction.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates /root/reference/models/archs/unet_arch.py (UNet :317-481, UpConvBlock :12-110, BasicConvBlock
:113-181, InterpConv :244-314), /root/reference/models/archs/fcn_arch.py (MultiHeadFCNHead :228-348) and
/root/reference/models/sample_model.py bot_index_prediction (:183-213), functionally over reference-keyed
state dicts, in eval mode.

Third-party arithmetic on this path: `mmcv.cnn.ConvModule` (mmcv-full==1.2.1, README.md:55; absent from
/root/reference and from this image).  Its published behaviour for the configuration used here
(conv_cfg=None, norm_cfg=dict(type='BN'), act_cfg=dict(type='ReLU'), order conv-norm-act, bias='auto'):
`nn.Conv2d(..., bias=False)` registered as `conv`, `nn.BatchNorm2d(out)` registered as `bn`, `nn.ReLU`
registered as `activate`; forward = activate(bn(conv(x))).  `conv_module` below restates exactly that.
"""
import torch
import torch.nn.functional as F


def conv_module(sd, p, x, padding):
    """mmcv ConvModule (conv without bias -> BatchNorm2d in eval mode -> ReLU)"""
    x = F.conv2d(x, sd[p + "conv.weight"], sd.get(p + "conv.bias"), padding=padding)
    x = F.batch_norm(x, sd[p + "bn.running_mean"], sd[p + "bn.running_var"], sd[p + "bn.weight"], sd[p + "bn.bias"],
                     False, 0.0, 1e-5)
    return F.relu(x)


def _n(sd, pattern):
    import re
    ids = {int(m.group(1)) for k in sd for m in [re.match(pattern, k)] if m}
    return max(ids) + 1


def unet(sd, x):
    """UNet.forward (:460-471) with the defaults used by sample_model.py:68-69: 5 stages, stride 1,
    MaxPool2d(2) before stages 1..4, two 3x3 ConvModules per block, InterpConv upsampling (bilinear x2,
    align_corners=False, then a 1x1 ConvModule).  Returns dec_outs (5 tensors, coarsest first)."""
    n_stage = _n(sd, r"encoder\.(\d+)\.")
    enc_outs = []
    for i in range(n_stage):
        if i != 0:
            x = F.max_pool2d(x, 2)                                             # :441
        blk = f"encoder.{i}.{0 if i == 0 else 1}.convs."
        for j in range(_n(sd, rf"encoder\.{i}\.\d+\.convs\.(\d+)\.")):
            x = conv_module(sd, f"{blk}{j}.", x, 1)                            # BasicConvBlock :160-170
        enc_outs.append(x)
    dec_outs = [x]
    for i in reversed(range(n_stage - 1)):
        p = f"decoder.{i}."
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)   # InterpConv :303
        x = conv_module(sd, p + "upsample.interp_upsample.1.", x, 0)           # 1x1 ConvModule :293-302
        x = torch.cat([enc_outs[i], x], dim=1)                                 # UpConvBlock.forward :105-108
        for j in range(_n(sd, rf"decoder\.{i}\.conv_block\.convs\.(\d+)\.")):
            x = conv_module(sd, f"{p}conv_block.convs.{j}.", x, 1)
        dec_outs.append(x)
    return dec_outs


def fcn_heads(sd, inputs, in_index=4):
    """MultiHeadFCNHead.forward (:313-327), num_convs=1, concat_input=False, eval (Dropout2d is identity):
    per head a 3x3 ConvModule then a 1x1 Conv2d with bias -> list of [B, num_classes, H, W]"""
    x = inputs[in_index]
    out = []
    for k in range(_n(sd, r"conv_seg_head_list\.(\d+)\.")):
        y = conv_module(sd, f"convs_list.{k}.0.", x, 1)
        out.append(F.conv2d(y, sd[f"conv_seg_head_list.{k}.weight"], sd[f"conv_seg_head_list.{k}.bias"]))
    return out


def bot_index_prediction(sd_unet, sd_fcn, feature_top, texture_mask, latent_hw=(32, 16)):
    """sample_model.py:183-213, batched: per position the argmax of its own texture's head, -1 elsewhere.
    -> list of 18 int64 [B, h, w]"""
    B = feature_top.shape[0]
    tex = F.interpolate(texture_mask, latent_hw, mode="nearest").view(-1).long()
    logits = fcn_heads(sd_fcn, unet(sd_unet, feature_top))
    out = [torch.full(tex.size(), -1, dtype=torch.long) for _ in range(len(logits))]
    for k, lg in enumerate(logits):
        roi = tex == k
        if torch.sum(roi) > 0:
            pred = lg.argmax(dim=1).view(-1)
            out[k][roi] = pred[roi]
    return [o.view(B, *latent_hw) for o in out]
