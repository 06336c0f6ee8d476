"""CPU study for the round-2 operand format: end-to-end heat-map error of PoseResNet (training-mode
BN, float64 accumulation, so that only the OPERAND representation is modelled) when every
convolution multiplies split operands
   tf32     : a_hi*b_hi                                  (single pass)
   tf32x3   : a_hi*b_hi + a_lo*b_hi + a_hi*b_lo          (today's kernels)
   f16x3    : the same three products with fp16 hi / lo planes and STATIC power-of-two scales
              (activations 2^4, weights 2^10) -- kind::f16 runs at twice the TF32 rate
against the exact float64 forward.  Also reports the largest scaled activation (fp16 overflows at
65504) per run.

    python tools/split_precision_study.py [layers] [HW] [N]
"""
import os, sys, types
import numpy as np, torch
import torch.nn.functional as TF
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import restate_net
from tests import golden_inputs as gi

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 50
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 128
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4
J, D = 3, HW // 4


def tf32_rn(f):
    u = f.contiguous().view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)


def tf32_tr(f):
    return (f.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


stats = {"amax": 0.0}


def split(t, mode, scale):
    f = t.float()
    if mode == "tf32":
        return tf32_rn(f).double(), None, 1.0
    if mode == "tf32x3":
        hi = tf32_rn(f)
        return hi.double(), tf32_tr(f - hi).double(), 1.0
    s = f * scale
    stats["amax"] = max(stats["amax"], float(s.abs().max()))
    hi = s.half()
    lo = (s - hi.float()).half()
    return hi.double(), lo.double(), scale


class Shim(types.SimpleNamespace):
    def __init__(self, mode, sa, sw):
        self.mode, self.sa, self.sw = mode, sa, sw

    def __getattr__(self, k):
        return getattr(TF, k)

    def _run(self, fn, x, w, bias, *a):
        ah, al, s1 = split(x, self.mode, self.sa)
        bh, bl, s2 = split(w, self.mode, self.sw)
        y = fn(ah, bh, None, *a)
        if al is not None:
            y = y + fn(al, bh, None, *a) + fn(ah, bl, None, *a)
        y = y / (s1 * s2)
        if bias is not None:
            y = y + bias.reshape(1, -1, 1, 1)
        return y

    def conv2d(self, x, w, bias=None, *a):
        return self._run(TF.conv2d, x, w, bias, *a)

    def conv_transpose2d(self, x, w, bias=None, *a):
        return self._run(TF.conv_transpose2d, x, w, bias, *a)


sd = restate_net.init_state(restate_net.param_shapes(layers, J, True, D), 5)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
x = torch.from_numpy(gi.images(N, HW, 5)).double()
with torch.no_grad():
    ref = restate_net.forward(sd64, x, num_layers=layers, training=True)
    ref32 = restate_net.forward(sd, x.float(), num_layers=layers, training=True)
    print("R%d %dx%d N%d heat-map rel err (max|d|/max|ref|) vs exact float64:" % (layers, HW, HW, N))
    print("  fp32 torch-CPU             %.3e" % float((ref32.double() - ref).abs().max() / ref.abs().max()))
    for mode, sa, sw in (("tf32", 1, 1), ("tf32x3", 1, 1), ("f16x3", 1.0, 1.0), ("f16x3", 16.0, 1024.0)):
        stats["amax"] = 0.0
        restate_net.F = Shim(mode, sa, sw)
        try:
            o = restate_net.forward(sd64, x, num_layers=layers, training=True)
        finally:
            restate_net.F = TF
        e = float((o - ref).abs().max() / ref.abs().max())
        extra = "" if mode != "f16x3" else "   scales (%g, %g), largest scaled operand %.1f" % (sa, sw, stats["amax"])
        print("  %-8s                   %.3e%s" % (mode, e, extra))
