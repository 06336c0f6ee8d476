"""One warm-up + one profiled training step of the bench workload (used under
ncu with --profile-from-start off).  Usage:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py [tuples] [precision]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "epipolarpose_b200"))
from tools.bench_cfg import make_cfg
import lib.models as models, lib.core.integral_loss as il, lib.utils.img_utils as iu, lib.utils.utils as U
from lib.dataset.synthetic import ring_camera

tuples = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "f16x3"
J, D, HW, V = 16, 64, 256, 4
dev = torch.device("cuda:0")
cfg = make_cfg(num_layers=50, num_joints=J, volume=True, depth_res=D, image_size=(HW, HW))
torch.manual_seed(0)
model = models.pose3d_resnet.get_pose_net(cfg, False, precision=prec).to(dev).train()
crit = il.SmoothL1JointLocationLoss(J)
opt = U.FusedAdam(list(model.parameters()), lr=1e-3)
n = tuples * V
rng = np.random.default_rng(1000)
order = [(t, 0) for t in range(tuples)] + [(t, 3) for t in range(tuples)] + [(t, 1) for t in range(tuples)] + [(t, 2) for t in range(tuples)]
cams = {(t, v): ring_camera(rng, v) for t in range(tuples) for v in range(V)}
meta = {"center_x": torch.tensor(500 + rng.uniform(-50, 50, n)), "center_y": torch.tensor(500 + rng.uniform(-50, 50, n)),
        "width": torch.tensor(800 + rng.uniform(-100, 100, n)), "height": torch.tensor(800 + rng.uniform(-100, 100, n)),
        "scale": torch.ones(n, dtype=torch.float64), "rot": torch.zeros(n, dtype=torch.float64),
        "R": torch.tensor(np.stack([cams[o][0] for o in order])), "T": torch.tensor(np.stack([cams[o][1] for o in order])),
        "f": torch.tensor(np.stack([cams[o][2] for o in order])), "c": torch.tensor(np.stack([cams[o][3] for o in order])),
        "projection_matrix": torch.tensor(np.stack([cams[o][4] for o in order]))}
meta = {k: v.to(dev) for k, v in meta.items()}
x = torch.randn(n, 3, HW, HW, device=dev)

from lib.core.function import online_epipolar_loss, _fused_head


def step():                                   # the body of GraphedTrainStep._eager / train_integral
    opt.zero_grad()
    with _fused_head(model):
        preds = model(x)
    loss = online_epipolar_loss(crit, preds, meta, "iterative")
    loss.backward()
    opt.step()
    return loss

step(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
l = step(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", l.item())
