"""One conv layer on the split-fp16 path, run a few times (target for ncu and A/B timing):

    python tools/one_conv16.py N H W Cin Cout k stride [what=fprop|dgrad|wgrad] [reps] [kind=conv|deconv]
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarpose_b200 import net, ops

N, H, W, cin, cout, k, s = [int(a) for a in sys.argv[1:8]]
what = sys.argv[8] if len(sys.argv) > 8 else "fprop"
reps = int(sys.argv[9]) if len(sys.argv) > 9 else 5
kind = sys.argv[10] if len(sys.argv) > 10 else "conv"
dev = torch.device("cuda:0")
conv = net.Conv("t", kind, cin, cout, k, s, (k // 2) if kind == "conv" else 1, 0)
torch.manual_seed(0)
Ho, Wo = conv.out_hw(H, W)
T = k * k


def split_of(t):
    h = torch.empty(2 * t.numel(), device=dev, dtype=torch.float16)
    sc = torch.ones(2, device=dev)
    ops.split16_batch(ops.SplitBatch([(t.reshape(-1), h, sc)]))
    return h, sc


x, x_sc = split_of(torch.relu(torch.randn(N, H, W, conv.cin_p, device=dev)))
x = x.view(2, N, H, W, conv.cin_p)
dz, dz_sc = split_of(torch.randn(N, Ho, Wo, conv.cout_p, device=dev) * 1e-4)
dz = dz.view(2, N, Ho, Wo, conv.cout_p)
wf, wf_sc = split_of(torch.randn(conv.cout_p * T * conv.cin_p, device=dev) * 0.05)
wd, wd_sc = split_of(torch.randn(conv.cin_p * T * conv.cout_p, device=dev) * 0.05)
out = torch.empty(N, Ho, Wo, conv.cout_p, device=dev)
din = torch.empty(N, H, W, conv.cin_p, device=dev)
dw = torch.zeros(conv.cout_p * T * conv.cin_p, device=dev)
ws = torch.empty(48 << 20, device=dev)
stats = torch.zeros(2 * conv.cout_p, device=dev, dtype=torch.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run():
    if what == "fprop":
        for g in conv.fprop_geoms(ops, N, H, W, 3):
            if g is not None:
                g.in_relu, g.accumulate = 0, 0
                ops.conv16_fprop(g, x, x_sc, wf, wf_sc, out, None, stats)
    elif what == "dgrad":
        for g in conv.dgrad_geoms(ops, N, H, W, 3):
            if g is not None:
                g.in_relu, g.accumulate = 0, 0
                ops.conv16_fprop(g, dz, dz_sc, wd, wd_sc, din, None, None)
    else:
        for g in conv.fprop_geoms(ops, N, H, W, 3):
            if g is not None:
                g.in_relu, g.accumulate = 0, 0
                ops.conv16_wgrad(g, x, x_sc, dz, dz_sc, dw, ws)


run(); run()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
M = N * (Ho * Wo if kind == "conv" else H * W)
fl = 2.0 * M * cin * cout * k * k
t = sorted(ts)[len(ts) // 2]
print("%s16 %s N%d %dx%d %d->%d k%d s%d: %.3f ms (median of %d), %.1f TFLOP/s algorithmic"
      % (what, kind, N, H, W, cin, cout, k, s, t, reps, fl / t / 1e9))
