"""Two-stream timeline of ONE eager training step (main stream + the weight-gradient side stream):
CUDA events around every libepb.so call, timestamps relative to the start of the step.

    python tools/timeline.py [tuples=32] [out.csv]

Prints how much of the step has both streams busy and how much longer each entry point runs when it
shares the GPU (ratio to the same call in a serialised step)."""
import collections
import os
import sys

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
out_csv = sys.argv[2] if len(sys.argv) > 2 else None
J, D, V, HW = 16, 64, 4, 256
dev = torch.device("cuda:0")
cfg = make_cfg(num_layers=50, num_joints=J, volume=True, depth_res=D, image_size=(HW, HW))
torch.manual_seed(0)
model = models.pose3d_resnet.get_pose_net(cfg, False, precision="f16x3").to(dev).train()
model.fused_head_gradient = True
crit = il.SmoothL1JointLocationLoss(J)
opt = U.FusedAdam(list(model.parameters()), lr=1e-3)
n = tuples * V
x = torch.randn(n, 3, HW, HW, device=dev)
lab = torch.rand(n, J * 3, device=dev) - 0.5
wt = torch.ones(n, J * 3, device=dev)
orig_call = ops._call
rec = []


def timed_call(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream
    e0.record()
    orig_call(name, *args)
    e1.record()
    rec.append((name, st, e0, e1))


def step():
    opt.zero_grad()
    loss = crit(model(x), lab, wt)
    loss.backward()
    opt.step()


def run(overlap):
    os.environ["EPB_OVERLAP_WGRAD"] = "1" if overlap else "0"
    step(); step()
    torch.cuda.synchronize()
    del rec[:]
    ops._call = timed_call
    base = torch.cuda.Event(enable_timing=True)
    base.record()
    step()
    torch.cuda.synchronize()
    ops._call = orig_call
    main = torch.cuda.current_stream().cuda_stream
    return [(nm, 0 if st == main else 1, base.elapsed_time(e0), base.elapsed_time(e1)) for nm, st, e0, e1 in rec]


ser = run(False)
ovl = run(True)
span = lambda rows: max(r[3] for r in rows) - min(r[2] for r in rows)
print("serialised step %.2f ms, overlapped step %.2f ms (eager, %d calls)" % (span(ser), span(ovl), len(ovl)))


def busy(rows, s):
    iv = sorted((r[2], r[3]) for r in rows if r[1] == s)
    return iv


def total(iv):
    return sum(b - a for a, b in iv)


def inter(a, b):
    i = j = 0
    t = 0.0
    while i < len(a) and j < len(b):
        lo, hi = max(a[i][0], b[j][0]), min(a[i][1], b[j][1])
        if hi > lo:
            t += hi - lo
        if a[i][1] < b[j][1]:
            i += 1
        else:
            j += 1
    return t


m, s = busy(ovl, 0), busy(ovl, 1)
print("main stream busy %.2f ms, side stream busy %.2f ms, both busy %.2f ms" % (total(m), total(s), inter(m, s)))
# per entry point: time in the overlapped step / time in the serialised step (same call order per stream)
dur = lambda rows: collections.Counter()
a, b = collections.OrderedDict(), collections.OrderedDict()
for rows, acc in ((ser, a), (ovl, b)):
    for nm, st, t0, t1 in rows:
        acc[nm] = acc.get(nm, 0.0) + (t1 - t0)
print("\n| entry point | serialised ms | overlapped ms | ratio |\n|---|---:|---:|---:|")
for nm in sorted(a, key=lambda k: -a[k]):
    print("| `%s` | %.3f | %.3f | %.2f |" % (nm, a[nm], b.get(nm, 0.0), b.get(nm, 0.0) / a[nm] if a[nm] else 0))
# what runs on the main stream while each weight-gradient call is in flight
ov = collections.Counter()
for nm, st, t0, t1 in ovl:
    if st != 1:
        continue
    for nm2, st2, u0, u1 in ovl:
        if st2 == 0:
            lo, hi = max(t0, u0), min(t1, u1)
            if hi > lo:
                ov[nm2] += hi - lo
print("\nmain-stream work concurrent with side-stream calls (ms): " +
      ", ".join("%s %.2f" % (k.replace("epb_", ""), v) for k, v in ov.most_common()))
if out_csv:
    with open(out_csv, "w") as f:
        f.write("name,stream,t0_ms,t1_ms\n")
        for nm, st, t0, t1 in ovl:
            f.write("%s,%d,%.4f,%.4f\n" % (nm, st, t0, t1))
