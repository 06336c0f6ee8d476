"""One conv layer of the bench workload, fprop (and optionally dgrad / wgrad) run a few times:
a target for `ncu -k regex:... --set full` and for A/B timing of kernel variants.

    python tools/one_conv.py N H W Cin Cout k stride [what=fprop|dgrad|wgrad] [reps]
"""
import os, sys
os.environ.setdefault("EPB_OVERLAP_WGRAD", "0")
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarpose_b200 import net, ops

N, H, W, cin, cout, k, s = [int(a) for a in sys.argv[1:8]]
what = sys.argv[8] if len(sys.argv) > 8 else "fprop"
reps = int(sys.argv[9]) if len(sys.argv) > 9 else 5
dev = torch.device("cuda:0")
conv = net.Conv("t", "conv", cin, cout, k, s, k // 2, 0)
eng = net.Engine(None, precision=3)
eng.dev = dev
torch.manual_seed(0)
w = torch.randn(cout, cin, k, k, device=dev) * 0.05
x = torch.randn(N, H, W, conv.cin_p, device=dev)
sc, sh = torch.rand(conv.cin_p, device=dev) + 0.5, torch.randn(conv.cin_p, device=dev) * 0.1
wf, wd = conv.pack(ops, w)
Ho, Wo = conv.out_hw(H, W)
gout = torch.randn(N, Ho, Wo, conv.cout_p, device=dev)
gw = torch.zeros_like(w)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run():
    if what == "fprop":
        eng._conv_fwd(conv, x, N, H, W, wf, affine=(sc, sh))
    elif what == "dgrad":
        eng._conv_dgrad(conv, gout, N, H, W, wd)
    else:
        eng._conv_wgrad(conv, x, gout, N, H, W, gw, affine=(sc, sh))


run(); run()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()                      # evict the operands from L2 between repetitions
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
M = N * Ho * Wo
fl = 2.0 * M * cin * cout * k * k
t = sorted(ts)[len(ts) // 2]
print("%s N%d %dx%d %d->%d k%d s%d: %.3f ms (median of %d, incl. weight prep), %.1f TFLOP/s algorithmic"
      % (what, N, H, W, cin, cout, k, s, t, reps, fl / t / 1e9))
