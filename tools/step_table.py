"""Per-C-ABI-call timing of ONE eager training step of the bench workload (CUDA events around
every libepb.so call, weight gradients serialised), any precision mode:

    python tools/step_table.py [tuples=32] [precision=f16x3] [layers=50] [hw=256]

Prints the step time, the time per entry point, and the conv calls grouped by shape with
their algorithmic TFLOP/s.  Markdown on stdout (profiles/*_step_table.md)."""
import collections
import os
import sys

os.environ.setdefault("EPB_OVERLAP_WGRAD", "0")   # serialise wgrad: clean per-call times
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "epipolarpose_b200"))
from epipolarpose_b200 import ops
import lib.models as models
import lib.core.integral_loss as il
import lib.utils.utils as U
from tools.bench_cfg import make_cfg

tuples = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "f16x3"
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 50
HW = int(sys.argv[4]) if len(sys.argv) > 4 else 256
J, D, V = 16, 64, 4
dev = torch.device("cuda:0")
cfg = make_cfg(num_layers=layers, num_joints=J, volume=True, depth_res=D, image_size=(HW, HW))
torch.manual_seed(0)
model = models.pose3d_resnet.get_pose_net(cfg, False, precision=prec).to(dev).train()
model.fused_head_gradient = True          # as inside train_integral / GraphedTrainStep
crit = il.SmoothL1JointLocationLoss(J)
opt = U.FusedAdam(list(model.parameters()), lr=1e-3)
n = tuples * V
x = torch.randn(n, 3, HW, HW, device=dev)
lab = torch.rand(n, J * 3, device=dev) - 0.5
wt = torch.ones(n, J * 3, device=dev)
rec = []
orig_call = ops._call


# elementwise passes: name -> (index of M, index of C, algorithmic bytes per element,
#                               [(index of an optional pointer argument, extra bytes per element)])
ELEM = {"epb_bn_bwd_reduce_mx": (8, 9, 8, [(2, 2)]),
        "epb_bn_bwd_apply_split": (11, 12, 12, [(2, 2), (15, 4)]),
        "epb_bn_act_split": (9, 10, 8, [(3, 4), (6, 4)]),
        "epb_bn_bwd_split": (10, 11, 20, [(2, 4), (14, 4)])}        # bit masks: 1/4 byte, not counted


def timed_call(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig_call(name, *args)
    e1.record()
    g = getattr(args[0], "_obj", None) if args else None
    shape = None
    if g is not None and hasattr(g, "Hp"):
        shape = (g.N * g.Hp * g.Wp, g.Cin, g.Cout, g.T, g.os, g.is_)
    elif name in ELEM:
        im, ic, b, opt_ = ELEM[name]
        extra = sum(bb for (ia, bb) in opt_ if args[ia] is not None)
        shape = ("elem", int(args[im]), int(args[ic]), b + extra)
    rec.append((name, shape, e0, e1))


def step():
    opt.zero_grad()
    loss = crit(model(x), lab, wt)
    loss.backward()
    opt.step()
    return loss


step(); step()
torch.cuda.synchronize()
ops._call = timed_call
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); step(); t1.record()
torch.cuda.synchronize()
ops._call = orig_call
print("# one eager step, %d tuples x %d views, R%d %dx%d, precision %s (%s)\n" %
      (tuples, V, layers, HW, HW, prec, type(model._engine()).__name__))
print("step %.2f ms (eager, wgrad serialised), %d C-ABI calls\n" % (t0.elapsed_time(t1), len(rec)))
by = collections.OrderedDict()
for name, shape, e0, e1 in rec:
    a = by.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
tot = sum(a[1] for a in by.values())
print("| entry point | calls | ms | share |\n|---|---:|---:|---:|")
for name, (c, ms) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% |" % (name, c, ms, 100 * ms / tot))
print("| **total** | %d | %.3f | |\n" % (len(rec), tot))
agg = collections.OrderedDict()
for name, shape, e0, e1 in rec:
    if shape is None or shape[0] == "elem":
        continue
    a = agg.setdefault((name,) + shape, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
print("| call | M | Cin | Cout | T | os | is | calls | ms | alg. TFLOP/s |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
fam = collections.OrderedDict()
for (name, M, ci, co, T, os_, is_), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    fl = 2.0 * M * ci * co * T * cnt
    print("| %s | %d | %d | %d | %d | %d | %d | %d | %.3f | %.1f |" %
          (name.replace("epb_", ""), M, ci, co, T, os_, is_, cnt, ms, fl / ms / 1e9))
    f = fam.setdefault(name, [0.0, 0.0])
    f[0] += fl
    f[1] += ms
print()
for name, (fl, ms) in fam.items():
    print("* %s: %.2f TFLOP in %.2f ms = %.1f TFLOP/s algorithmic" % (name, fl / 1e12, ms, fl / ms / 1e9))

print("\n| elementwise pass | M | C | B/elem | calls | ms | GB/s |\n|---|---:|---:|---:|---:|---:|---:|")
el = collections.OrderedDict()
for name, shape, e0, e1 in rec:
    if shape is None or shape[0] != "elem":
        continue
    a = el.setdefault((name,) + shape[1:], [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
for (name, M, C, b), (cnt, ms) in sorted(el.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %d | %d | %d | %.3f | %.0f |" % (name.replace("epb_", ""), M, C, b, cnt, ms,
                                                      float(M) * C * b * cnt / ms / 1e6))
