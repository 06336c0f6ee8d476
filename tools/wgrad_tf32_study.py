"""CPU study: what does single-pass TF32 (instead of 3xTF32) in the WEIGHT-GRADIENT contraction cost
in gradient parity?  Weight gradients are leaves of the backward pass (their rounding error is not
propagated), so only the contraction itself matters: dW = sum over N*H*W pixels of
tf32(dout) * tf32(act).  The reference PoseResNet (float64, training-mode BN) gives exact
activations / output gradients per conv through hooks; dW is recomputed from TF32-rounded operands
with float64 accumulation and compared per tensor (max|d| / max|ref|, the parity metric).

    PYTHONDONTWRITEBYTECODE=1 python tools/wgrad_tf32_study.py [layers] [HW] [N]     (build container only)
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim, restate_net
from tests import golden_inputs as gi

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 50
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 128
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4
J, D = 3, HW // 4
r = refshim.ref()
cfg = refshim.make_cfg(num_layers=layers, num_joints=J, volume=True, depth_res=D, image_size=(HW, HW))
model = r.pose3d_resnet.get_pose_net(cfg, False)
model.load_state_dict(restate_net.init_state(restate_net.param_shapes(layers, J, True, D), 5))
model = model.double().train()


def tf32(t):
    f = t.float().contiguous()
    u = f.view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32).double()


rec = {}
for name, m in model.named_modules():
    if isinstance(m, torch.nn.Conv2d):
        m.register_forward_hook(lambda mod, inp, out, name=name: rec.setdefault(name, {}).update(x=inp[0].detach()))
        m.register_full_backward_hook(lambda mod, gin, gout, name=name: rec[name].update(g=gout[0].detach()))
x = torch.from_numpy(gi.images(N, HW, 5)).double()
gt, wt = gi.labels(N, J, 5)
out = model(x)
sm = torch.softmax(out.reshape(N, J, -1), 2).reshape(N, J, D, D, D)
ar = torch.arange(D, dtype=torch.float64)
c = torch.stack([(sm.sum((2, 3)) * ar).sum(2) / D - 0.5, (sm.sum((2, 4)) * ar).sum(2) / D - 0.5,
                 (sm.sum((3, 4)) * ar).sum(2) / D - 0.5], 2).reshape(N, J * 3)
((c - torch.from_numpy(gt).double()).abs() * torch.from_numpy(wt).double()).sum().div(N).backward()
worst = []
for name, m in model.named_modules():
    if isinstance(m, torch.nn.Conv2d) and name in rec and "g" in rec[name]:
        ref = m.weight.grad
        dw = torch.nn.grad.conv2d_weight(tf32(rec[name]["x"]), m.weight.shape, tf32(rec[name]["g"]),
                                         stride=m.stride, padding=m.padding)
        e = float((dw - ref).abs().max() / ref.abs().max())
        worst.append((e, name, tuple(m.weight.shape), rec[name]["x"].shape[0] * rec[name]["g"].shape[2] * rec[name]["g"].shape[3]))
worst.sort(reverse=True)
print("R%d %dx%d N%d: single-pass TF32 wgrad, rel err per weight tensor (max|d|/max|ref|); worst 8 of %d:" % (layers, HW, HW, N, len(worst)))
for e, name, shp, k in worst[:8]:
    print("  %.3e  %-28s %s  reduction length %d" % (e, name, shp, k))
print("  median %.3e" % np.median([w[0] for w in worst]))
