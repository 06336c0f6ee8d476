"""BASELINE config C5 sanity on the GPU: ResNet-101, 384x384, J = 17, DEPTH_RES = 96 (1632 x 96 x 96
logits), training-mode forward against the float64 oracle -- next to the distance of the torch-CPU
fp32 oracle from float64, which is the noise floor of the comparison: at N = 2 the batch statistics
of a 101-layer net are ill conditioned (fp32 vs fp64: 5e-4, fp32 with 1 thread vs 16 threads: 2.5e-4;
round-1 GPU run: 2.1e-3 vs the fp32 oracle) -- and one full backward + fused Adam step through the
public surface (finite, non-zero gradients).  Report only; fails on gross errors (> 1e-2).

    python tools/c5_check.py [N]
"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "epipolarpose_b200"))
from oracle import refshim, restate_net
from tests import golden_inputs as gi
import lib.models as models, lib.core.integral_loss as il, lib.utils.utils as U

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
J, D, HW, L = 17, 96, 384, 101
dev = torch.device("cuda:0")
cfg = refshim.make_cfg(num_layers=L, num_joints=J, volume=True, depth_res=D, image_size=(HW, HW))
model = models.pose3d_resnet.get_pose_net(cfg, False)
sd = restate_net.init_state(restate_net.param_shapes(L, J, True, D), 11)
model.load_state_dict(sd)
model = model.to(dev).train()
x = torch.from_numpy(gi.images(N, HW, 11))
opt = U.FusedAdam(list(model.parameters()), lr=1e-3)
t0 = time.time()
preds = model(x.to(dev))
gt, wt = gi.labels(N, J, 11)
loss = il.SmoothL1JointLocationLoss(J)(preds, torch.from_numpy(gt).to(dev), torch.from_numpy(wt).to(dev))
loss.backward()
opt.step()
torch.cuda.synchronize()
print("gpu step %.2f s, logits %s, loss %.6f" % (time.time() - t0, tuple(preds.shape), loss.item()))
gn = [float(p.grad.abs().max()) for p in model.parameters() if p.grad is not None]
assert all(np.isfinite(gn)) and min(gn) >= 0 and max(gn) > 0, "gradients"
with torch.no_grad():
    ref32 = restate_net.forward(sd, x, num_layers=L, training=True)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    ref64 = restate_net.forward(sd64, x.double(), num_layers=L, training=True)
rel = lambda a: float((a.double() - ref64).abs().max() / ref64.abs().max())
err, floor = rel(preds.detach().cpu()), rel(ref32)
print("C5 R101 384x384 J17 D96 N%d: heat-map rel err vs float64 oracle: GPU %.3e, torch-CPU fp32 %.3e; "
      "%d parameter tensors with finite gradients, max |g| %.3e" % (N, err, floor, len(gn), max(gn)))
assert err <= 1e-2
print("C5 ok")
