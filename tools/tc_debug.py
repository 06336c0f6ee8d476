"""GPU debug harness for the tcgen05 conv path: runs a few 1x1 / 3x3 problems
at precision 1 (TF32) and 3 (3xTF32) against torch fp32 and prints error
structure (which rows / columns / k-ranges are wrong) to localise descriptor or
swizzle mistakes.  Not part of the product or the test-suite."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "epipolarpose_b200"))
from epipolarpose_b200 import net, ops
import torch.nn.functional as F

dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(kind, cin, cout, k, s, p, hw, N, precision, affine=True, verbose=True):
    conv = net.Conv("t", kind, cin, cout, k, s, p, 0)
    eng = net.Engine(None, precision=precision)
    eng.dev = dev
    w = torch.randn(*((cout, cin, k, k) if kind == "conv" else (cin, cout, k, k)), device=dev) * 0.1
    x = torch.randn(N, cin, hw, hw, device=dev)
    sc = torch.rand(cin, device=dev) + 0.5
    sh = torch.randn(cin, device=dev) * 0.1
    xa = torch.relu(x * sc[None, :, None, None] + sh[None, :, None, None]) if affine else x
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = (F.conv2d(xa.double(), w.double(), None, s, p) if kind == "conv"
           else F.conv_transpose2d(xa.double(), w.double(), None, s, p)).float()
    xn = torch.zeros(N, hw, hw, conv.cin_p, device=dev)
    ops.nchw_to_nhwc(x.contiguous(), xn, N, cin, hw, hw, conv.cin_p)
    wf, wd = conv.pack(ops, w)
    stats = torch.zeros(2 * conv.cout_p, device=dev, dtype=torch.float64)
    try:
        out, Ho, Wo = eng._conv_fwd(conv, xn, N, hw, hw, wf, affine=(sc, sh) if affine else None,
                                    stats=stats)
        torch.cuda.synchronize()
    except Exception as e:
        print("FAILED", kind, cin, cout, k, s, precision, repr(e)[:300])
        return None
    o = out[..., :cout].permute(0, 3, 1, 2)
    err = (o - ref).abs()
    rel = (err.max() / ref.abs().max()).item()
    st_ref = torch.cat([ref.double().sum((0, 2, 3)), (ref.double() ** 2).sum((0, 2, 3))])
    st = torch.cat([stats[:cout], stats[conv.cout_p:conv.cout_p + cout]])
    srel = ((st - st_ref).abs().max() / st_ref.abs().max()).item()
    # dgrad + wgrad through the same geometry
    gout = torch.randn_like(ref)
    xa_ = xa.detach().clone().requires_grad_(True); w_ = w.detach().clone().requires_grad_(True)
    r2 = (F.conv2d(xa_.double(), w_.double(), None, s, p) if kind == "conv"
          else F.conv_transpose2d(xa_.double(), w_.double(), None, s, p))
    gx, gw_ref = torch.autograd.grad(r2, (xa_, w_), gout.double())
    gn = torch.zeros(N, Ho, Wo, conv.cout_p, device=dev)
    ops.nchw_to_nhwc(gout.contiguous(), gn, N, cout, Ho, Wo, conv.cout_p)
    drel = wrel = float("nan")
    try:
        din = eng._conv_dgrad(conv, gn, N, hw, hw, wd)
        torch.cuda.synchronize()
        drel = ((din[..., :cin].permute(0, 3, 1, 2) - gx.float()).abs().max() / gx.abs().max()).item()
        gw = torch.zeros_like(w)
        eng._conv_wgrad(conv, xn, gn, N, hw, hw, gw, affine=(sc, sh) if affine else None)
        torch.cuda.synchronize()
        wrel = ((gw - gw_ref.float()).abs().max() / gw_ref.abs().max()).item()
    except Exception as e:
        print("   dgrad/wgrad FAILED", repr(e)[:300])
    print("%-6s cin %4d cout %4d k%d s%d hw %3d N %d prec %d : fprop %.3e stats %.3e dgrad %.3e wgrad %.3e"
          % (kind, cin, cout, k, s, hw, N, precision, rel, srel, drel, wrel), flush=True)
    if verbose and wrel > 5e-3:
        e = (gw - gw_ref.float()).abs()
        print("   wgrad bad: shape", tuple(gw.shape), "ratio sample",
              (gw.flatten()[:8] / gw_ref.float().flatten()[:8]).cpu().numpy().round(3).tolist(),
              "bad frac %.3f" % (e > 1e-2 * gw_ref.abs().max()).float().mean().item())
    if verbose and rel > 5e-3:
        e2 = err.permute(0, 2, 3, 1).reshape(-1, cout)      # [pixels][cout]
        r2 = ref.permute(0, 2, 3, 1).reshape(-1, cout)
        o2 = o.permute(0, 2, 3, 1).reshape(-1, cout)
        bad = (e2 > 1e-2 * ref.abs().max())
        print("   bad fraction %.4f ; bad rows (first 32 of 128-row tile):" % bad.float().mean().item(),
              bad.any(1)[:128].int().tolist()[:32])
        print("   bad cols (first 64):", bad.any(0)[:64].int().tolist())
        print("   ratio out/ref sample:", (o2[:4, :8] / r2[:4, :8]).cpu().numpy().round(3).tolist())
    return rel


if __name__ == "__main__":
    for prec in (1, 3):
        run("conv", 32, 64, 1, 1, 0, 16, 1, prec, affine=False)     # single tile, K = 32
        run("conv", 64, 64, 1, 1, 0, 16, 1, prec, affine=False)     # K = 64 (2 k-blocks)
        run("conv", 64, 128, 1, 1, 0, 16, 2, prec)
        run("conv", 256, 256, 1, 1, 0, 16, 2, prec)
        run("conv", 64, 64, 3, 1, 1, 14, 3, prec)
        run("conv", 64, 128, 3, 2, 1, 14, 3, prec)
        run("conv", 256, 1088, 1, 1, 0, 16, 2, prec)                # N tail
        run("deconv", 64, 32, 4, 2, 1, 7, 3, prec)
        run("conv", 512, 512, 3, 1, 1, 8, 8, prec)
        run("conv", 256, 64, 1, 1, 0, 64, 8, prec)                  # many tiles (persistent loop)
