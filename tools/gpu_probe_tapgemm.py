"""Bring-up probe for t2h_tapgemm on a real B200: progressively harder cases,
each synchronised and printed so a hang or a wrong descriptor is localised.
Usage (GPU box): timeout 600 python tools/gpu_probe_tapgemm.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text2human_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"


def report(name, got, ref):
    err = (got.double() - ref.double()).abs().max().item()
    scale = ref.double().abs().max().item()
    print(f"  {name}: max|err|={err:.3e} max|ref|={scale:.3e} rel={err / max(scale, 1e-30):.3e}", flush=True)
    return err / max(scale, 1e-30)


def planes_to_f32(p):
    return p.float().sum(0)


def case_linear(M, K, N, terms, bias=True, planes_out=False, gelu=False, residual=False):
    print(f"linear M={M} K={K} N={N} terms={terms} bias={bias} planes_out={planes_out} gelu={gelu} res={residual}",
          flush=True)
    g = torch.Generator(device=dev).manual_seed(M * 7 + K * 3 + N)
    x = torch.randn(M, K, device=dev, generator=g)
    w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
    b = torch.randn(N, device=dev, generator=g) if bias else None
    r = torch.randn(M, N, device=dev, generator=g) if residual else None
    a = ops.split_planes(x, terms)
    wp = ops.pack_linear_weight(w, terms)
    out = ops.linear(a, wp, b, planes_out=planes_out, act=ops.ACT_GELU if gelu else ops.ACT_NONE, residual=r)
    torch.cuda.synchronize()
    xe, we = planes_to_f32(a), planes_to_f32(wp)[0]
    ref = xe.double() @ we.double().t()
    if bias:
        ref = ref + b.double()
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if residual:
        ref = ref + r.double()
    got = planes_to_f32(out) if planes_out else out
    return report("vs fp64 of the split operands", got, ref)


def case_conv(N, H, W, Cin, Cout, terms, nchw_out=False, residual=False):
    print(f"conv3x3 N={N} H={H} W={W} Cin={Cin} Cout={Cout} terms={terms} nchw_out={nchw_out} res={residual}",
          flush=True)
    g = torch.Generator(device=dev).manual_seed(H * 5 + Cin + Cout)
    x = torch.randn(N, Cin, H, W, device=dev, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device=dev, generator=g) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, device=dev, generator=g)
    a = ops.nchw_to_planes(x, terms=terms)
    wp = ops.pack_conv_weight(w, terms, c_pad=a.shape[-1])
    r = torch.randn(N, H, W, Cout, device=dev, generator=g) if residual else None
    out = ops.conv3x3(a, wp, b, nchw_out=nchw_out, residual=r)
    torch.cuda.synchronize()
    xe = planes_to_f32(a)[..., :Cin].permute(0, 3, 1, 2).double()
    we = planes_to_f32(wp)[..., :Cin].reshape(3, 3, Cout, Cin).permute(2, 3, 0, 1).double()
    ref = torch.nn.functional.conv2d(xe, we, b.double(), padding=1)
    got = out if nchw_out else out.permute(0, 3, 1, 2)
    if residual:
        ref = ref + r.permute(0, 3, 1, 2).double()
    return report("vs fp64 conv of the split operands", got, ref)


def case_conv_s2(N, H, W, Cc, Cout, terms):
    print(f"conv3x3 s2 N={N} H={H} W={W} C={Cc} Cout={Cout} terms={terms}", flush=True)
    g = torch.Generator(device=dev).manual_seed(H + Cc)
    x = torch.randn(N, H, W, Cc, device=dev, generator=g)
    w = torch.randn(Cout, Cc, 3, 3, device=dev, generator=g) / (9 * Cc) ** 0.5
    b = torch.randn(Cout, device=dev, generator=g)
    a = ops.f32_to_planes(x, ops.CVT_S2D, terms)
    wp = ops.pack_conv_weight(w, terms)
    out = ops.conv3x3_s2(a, wp, b)
    torch.cuda.synchronize()
    xs = ops.split_planes(x, terms).float().sum(0).permute(0, 3, 1, 2).double()
    we = planes_to_f32(wp).reshape(3, 3, Cout, Cc).permute(2, 3, 0, 1).double()
    ref = torch.nn.functional.conv2d(torch.nn.functional.pad(xs, (0, 1, 0, 1)), we, b.double(), stride=2)
    return report("vs fp64 conv", out.permute(0, 3, 1, 2), ref)


def case_bmm(G, M, N, K, terms):
    print(f"bmm_nt G={G} M={M} N={N} K={K} terms={terms}", flush=True)
    g = torch.Generator(device=dev).manual_seed(G + M + N)
    a = torch.randn(G, M, K, device=dev, generator=g)
    b = torch.randn(G, N, K, device=dev, generator=g)
    ap, bp = ops.split_planes(a, terms), ops.split_planes(b, terms)
    out = ops.bmm_nt(ap, bp, alpha=0.125)
    torch.cuda.synchronize()
    ref = 0.125 * planes_to_f32(ap).double() @ planes_to_f32(bp).double().transpose(1, 2)
    return report("vs fp64", out, ref)


def main():
    from text2human_b200 import _lib
    lib = _lib.load()
    print("libt2h version", lib.t2h_version(), torch.cuda.get_device_name(0), flush=True)
    worst = 0.0
    t0 = time.time()
    worst = max(worst, case_linear(128, 64, 64, 1, bias=False))
    worst = max(worst, case_linear(128, 64, 16, 1, bias=False))
    worst = max(worst, case_linear(256, 128, 128, 1))
    worst = max(worst, case_linear(300, 512, 256, 1))
    worst = max(worst, case_linear(2048, 512, 1536, 1, planes_out=True))
    worst = max(worst, case_linear(2048, 512, 2048, 2, gelu=True, planes_out=True))
    worst = max(worst, case_linear(2048, 2048, 512, 2, residual=True))
    worst = max(worst, case_linear(4096, 32, 32, 2))
    worst = max(worst, case_conv(2, 32, 16, 64, 128, 1))
    worst = max(worst, case_conv(2, 32, 16, 128, 256, 2, residual=True))
    worst = max(worst, case_conv(1, 64, 32, 3, 128, 2))
    worst = max(worst, case_conv(2, 64, 32, 128, 3, 2, nchw_out=True))
    worst = max(worst, case_conv(4, 256, 128, 128, 128, 1))  # MBLK=2 path
    worst = max(worst, case_conv(1, 16, 8, 512, 512, 2))
    worst = max(worst, case_conv_s2(2, 64, 32, 128, 128, 2))
    worst = max(worst, case_bmm(4, 512, 512, 512, 2))
    worst = max(worst, case_bmm(8, 512, 512, 64, 1))
    print(f"worst rel err {worst:.3e}  ({time.time() - t0:.1f}s)", flush=True)
    print("PROBE_OK" if worst < 1e-3 else "PROBE_BAD", flush=True)


if __name__ == "__main__":
    main()
