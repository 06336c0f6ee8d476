"""Per-conv-call timing table of one training step of the bench workload
(CUDA events around each epb_conv_fprop / epb_conv_wgrad call)."""
import os, sys, collections
os.environ.setdefault("EPB_OVERLAP_WGRAD", "0")   # serialise wgrad: clean per-call times
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "epipolarpose_b200"))
from oracle import refshim
from epipolarpose_b200 import ops
import lib.models as models, lib.core.integral_loss as il, lib.utils.utils as U

tuples = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
J, D, HW, V = 16, 64, 256, 4
dev = torch.device("cuda:0")
cfg = refshim.make_cfg(num_layers=50, num_joints=J, volume=True, depth_res=D, image_size=(HW, HW))
torch.manual_seed(0)
model = models.pose3d_resnet.get_pose_net(cfg, False, precision=prec).to(dev).train()
crit = il.SmoothL1JointLocationLoss(J)
opt = U.FusedAdam(list(model.parameters()), lr=1e-3)
n = tuples * V
x = torch.randn(n, 3, HW, HW, device=dev)
lab = torch.rand(n, J * 3, device=dev) - 0.5
wt = torch.ones(n, J * 3, device=dev)
rec = []
of, ow = ops.conv_fprop, ops.conv_wgrad
def wrap(f, kind):
    def w(g, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(g, *a, **k); e1.record()
        M = g.N * g.Hp * g.Wp
        rec.append((kind, M, g.Cin, g.Cout, g.T, g.os, g.is_, e0, e1))
    return w
def step():
    opt.zero_grad(); loss = crit(model(x), lab, wt); loss.backward(); opt.step(); return loss
step(); step(); torch.cuda.synchronize()
ops.conv_fprop, ops.conv_wgrad = wrap(of, "fprop"), wrap(ow, "wgrad")
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); step(); t1.record(); torch.cuda.synchronize()
print("step %.2f ms" % t0.elapsed_time(t1))
agg = collections.OrderedDict()
for kind, M, ci, co, T, os_, is_, e0, e1 in rec:
    key = (kind, M, ci, co, T, os_, is_)
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1)
tot = sum(a[1] for a in agg.values())
print("conv total %.2f ms over %d calls" % (tot, len(rec)))
print("%-6s %8s %5s %5s %3s %2s %2s %4s %8s %8s" % ("kind", "M", "Cin", "Cout", "T", "os", "is", "n", "ms", "TFLOP/s"))
for (kind, M, ci, co, T, os_, is_), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    fl = 2.0 * M * ci * co * T * cnt
    print("%-6s %8d %5d %5d %3d %2d %2d %4d %8.3f %8.1f" % (kind, M, ci, co, T, os_, is_, cnt, ms, fl / ms / 1e9))
