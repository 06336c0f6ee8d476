"""Per-parameter gradient error listing (ours vs fp64 oracle, fp32 oracle vs fp64)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "epipolarpose_b200"))
from oracle import refshim, restate_net
from tests import golden_inputs as gi
import lib.models as models
import lib.core.integral_loss as il
dev = torch.device("cuda:0")
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 18
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
J, D, N = 3, HW // 4, 4
cfg = refshim.make_cfg(num_layers=layers, num_joints=J, volume=True, depth_res=D, image_size=(HW, HW))
sd = restate_net.init_state(restate_net.param_shapes(layers, J, True, D), 5)
x = gi.images(N, HW, 5); gt, wt = gi.labels(N, J, 5)
def oracle(dt):
    p = {k: (v.to(dt).clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
             else (v.to(dt) if v.is_floating_point() else v)) for k, v in sd.items()}
    o = restate_net.forward(p, torch.from_numpy(x).to(dt), num_layers=layers, training=True)
    sm = torch.softmax(o.reshape(N, J, -1), 2).reshape(N, J, D, D, D)
    ar = torch.arange(D, dtype=dt)
    c = torch.stack([(sm.sum((2, 3)) * ar).sum(2) / D - 0.5, (sm.sum((2, 4)) * ar).sum(2) / D - 0.5,
                     (sm.sum((3, 4)) * ar).sum(2) / D - 0.5], 2).reshape(N, J * 3)
    loss = ((c - torch.from_numpy(gt).to(dt)).abs() * torch.from_numpy(wt).to(dt)).sum() / N
    loss.backward()
    return loss.item(), {k: v.grad for k, v in p.items() if getattr(v, "grad", None) is not None}
l64, g64 = oracle(torch.float64); l32, g32 = oracle(torch.float32)
model = models.pose3d_resnet.get_pose_net(cfg, False, precision=prec)
model.load_state_dict(sd); model = model.to(dev).train()
loss = il.L1JointLocationLoss(J)(model(torch.from_numpy(x).to(dev)), torch.from_numpy(gt).to(dev), torch.from_numpy(wt).to(dev))
loss.backward()
print("loss ours %.7f fp64 %.7f fp32 %.7f" % (loss.item(), l64, l32))
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
for k, p in model.named_parameters():
    print("%-36s ours %.2e  fp32 %.2e" % (k, rel(p.grad.cpu().numpy().astype(np.float64), g64[k].numpy()), rel(g32[k].numpy().astype(np.float64), g64[k].numpy())))
