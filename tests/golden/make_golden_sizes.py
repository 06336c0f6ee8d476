"""Golden vectors at the BASELINE.json configuration sizes, produced by the UNMODIFIED
reference module (/root/reference lib/models/pose3d_resnet.py through oracle/refshim.py) on the
CPU in float32.  Run in the build container only:
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_sizes.py [c1] [c2]

  c1  configs[0]: ResNet-50, 256x256, batch 1, .eval(), J=16 (experiments/mpii), VOLUME D=64
  c2  configs[1] slice: ResNet-50, 256x256, J=17 D=64, .train(), N=8, forward + backward

The tensors are too large to store whole (C2: 142 MB of logits, 137 MB of gradients), so the
files hold strided samples plus per-channel / per-tensor sums; tests/golden_inputs.py has the
sampling rules so the GPU tests take the same samples."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import refshim, restate_net  # noqa: E402
from tests import golden_inputs as gi  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
r = refshim.ref()
which = sys.argv[1:] or ["c1", "c2"]
torch.set_num_threads(os.cpu_count())


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, len(kw), "arrays")


for tag in which:
    c = gi.SIZE_CASES[tag]
    t0 = time.time()
    cfg = refshim.make_cfg(num_layers=c["layers"], num_joints=c["J"], volume=True, depth_res=c["D"],
                           image_size=(c["HW"], c["HW"]))
    model = r.pose3d_resnet.get_pose_net(cfg, False)
    shapes = restate_net.param_shapes(num_layers=c["layers"], num_joints=c["J"], volume=True,
                                      depth_res=c["D"])
    model.load_state_dict(restate_net.init_state(shapes, c["seed"]))
    x = torch.from_numpy(gi.images(c["N"], c["HW"], c["seed"]))
    rec = {}
    if c["train"]:
        model.train()
        out = model(x)
        g = torch.from_numpy(gi.grad_like_big(out.shape, c["seed"] + 1))
        (out * g).sum().backward()
        rec.update(gi.sample_output(out.detach().numpy()))
        for k, p in model.named_parameters():
            s, tot = gi.sample_grad(p.grad.numpy())
            rec["grad/" + k] = s
            rec["gsum/" + k] = tot
        sd = model.state_dict()
        rec["bn1.running_mean"] = sd["bn1.running_mean"].numpy()
        rec["bn1.running_var"] = sd["bn1.running_var"].numpy()
    else:
        model.eval()
        with torch.no_grad():
            out = model(x)
        rec.update(gi.sample_output(out.numpy()))
    save("net_" + tag, **rec)
    print(tag, "%.1f s" % (time.time() - t0))
