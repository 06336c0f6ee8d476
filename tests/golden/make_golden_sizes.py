"""Golden vectors at the BASELINE.json configuration sizes.  Run in the build container only:
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_sizes.py [c1] [c2] [c5]

  c1  configs[0]: ResNet-50, 256x256, batch 1, .eval(), J=16 (experiments/mpii), VOLUME D=64
  c2  configs[1] slice: ResNet-50, 256x256, J=17 D=64, .train(), N=8, forward + backward
  c5  configs[4] slice: ResNet-101, 384x384, J=17 D=96, .train(), N=16 (4 tuples x 4 views)

Two evaluations of the same seeded inputs are stored per case:
  ref/...  the UNMODIFIED reference module (/root/reference lib/models/pose3d_resnet.py through
           oracle/refshim.py), float32 on the CPU -- the parity target of north_star;
  f64/...  the oracle restatement (oracle/restate_net.py, pinned to the reference by
           tests/test_oracle_pinned.py) in float64 -- the yardstick that tells how far the
           float32 reference run itself is from the exact result.  Random-init 50/101-layer
           BatchNorm networks amplify rounding through the backward pass: at these sizes the
           reference's own float32 gradients sit up to several 1e-2 from the float64 ones (and
           from a second float32 evaluation with another summation order), so a gradient test
           can only ask an implementation to be as close to float64 as the reference is.

The tensors are too large to store whole, so the files hold strided samples plus sums
(tests/golden_inputs.py: sample_output / sample_grad)."""
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
which = sys.argv[1:] or ["c1", "c2", "c5"]
torch.set_num_threads(os.cpu_count())


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, len(kw), "arrays")


def record(rec, prefix, out, grads, state=None):
    for k, v in gi.sample_output(out).items():
        rec[prefix + k] = v
    for k, g in (grads or {}).items():
        s, tot = gi.sample_grad(g)
        rec[prefix + "grad/" + k] = s
        rec[prefix + "gsum/" + k] = tot
    if state is not None:
        rec[prefix + "bn1.running_mean"] = state["bn1.running_mean"].numpy()
        rec[prefix + "bn1.running_var"] = state["bn1.running_var"].numpy()


for tag in which:
    c = gi.SIZE_CASES[tag]
    t0 = time.time()
    cfg = refshim.make_cfg(num_layers=c["layers"], num_joints=c["J"], volume=True, depth_res=c["D"],
                           image_size=(c["HW"], c["HW"]))
    shapes = restate_net.param_shapes(num_layers=c["layers"], num_joints=c["J"], volume=True,
                                      depth_res=c["D"])
    sd = restate_net.init_state(shapes, c["seed"])
    x = torch.from_numpy(gi.images(c["N"], c["HW"], c["seed"]))
    rec = {}
    # ---- the unmodified reference, float32
    model = r.pose3d_resnet.get_pose_net(cfg, False)
    model.load_state_dict(sd)
    if c["train"]:
        model.train()
        out = model(x)
        g = torch.from_numpy(gi.grad_like_big(out.shape, c["seed"] + 1))
        (out * g).sum().backward()
        record(rec, "ref/", out.detach().numpy(), {k: p.grad.numpy() for k, p in model.named_parameters()},
               model.state_dict())
    else:
        model.eval()
        with torch.no_grad():
            out = model(x)
        record(rec, "ref/", out.numpy(), None)
    del model, out
    print(tag, "reference %.1f s" % (time.time() - t0), flush=True)
    # ---- the oracle restatement, float64
    t0 = time.time()
    p = {k: (v.double().clone().requires_grad_(c["train"]) if v.is_floating_point() and "running" not in k
             else (v.double() if v.is_floating_point() else v)) for k, v in sd.items()}
    if c["train"]:
        out = restate_net.forward(p, x.double(), num_layers=c["layers"], volume=True,
                                  image_size=(c["HW"], c["HW"]))
        (out * g.double()).sum().backward()
        record(rec, "f64/", out.detach().numpy(),
               {k: v.grad.numpy() for k, v in p.items() if torch.is_tensor(v) and v.requires_grad})
    else:
        with torch.no_grad():
            out = restate_net.forward(p, x.double(), num_layers=c["layers"], volume=True,
                                      image_size=(c["HW"], c["HW"]), training=False)
        record(rec, "f64/", out.numpy(), None)
    print(tag, "float64 oracle %.1f s" % (time.time() - t0), flush=True)
    save("net_" + tag, **rec)
