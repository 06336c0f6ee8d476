"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported read-only through oracle/refshim.py) on seeded
synthetic inputs.  Run in the build container only:
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
Inputs are regenerated from the seeds in tests/golden_inputs.py, so only the
reference OUTPUTS are stored (small files)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import refshim, restate_net  # noqa: E402
from tests import golden_inputs as gi  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
r = refshim.ref()
il, tri, iu, inf = r.integral_loss, r.triangulation, r.img_utils, r.inference


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: np.asarray(v).shape for k, v in kw.items()})


# ---- soft-argmax + losses (integral_loss.py:7-160)
for tag, (N, J, D, H, W, seed, scale) in gi.SOFTARGMAX_CASES.items():
    logits = torch.from_numpy(gi.logits(N, J, D, H, W, seed, scale)).requires_grad_(True)
    coords = il.softmax_integral_tensor(logits, J, True, W, H, D)
    gt, wt = gi.labels(N, J, seed)
    out = {"coords": coords.detach().numpy()}
    for cls, key in ((il.L1JointLocationLoss, "l1"), (il.SmoothL1JointLocationLoss, "smoothl1")):
        for norm in (False, True):
            logits.grad = None
            crit = cls(J, norm=norm)
            loss = crit(logits, torch.from_numpy(gt), torch.from_numpy(wt))
            loss.backward()
            k = key + ("_norm" if norm else "")
            out[k + "_loss"] = loss.detach().numpy()
            out[k + "_grad_sum_abs"] = logits.grad.abs().sum((2, 3)).numpy()   # [N, J*D]
            out[k + "_grad_sample"] = logits.grad[:, :, ::3, ::3].numpy().copy()
    logits.grad = None
    mse = il.weighted_mse_loss(il.softmax_integral_tensor(logits, J, True, W, H, D),
                               torch.from_numpy(gt), torch.from_numpy(wt), True)
    out["mse_loss"] = mse.detach().numpy()
    out["result"] = il.get_joint_location_result(256, 256, logits.detach()) if D == W else np.zeros(0)
    save("softargmax_" + tag, **out)

# ---- hard argmax (inference.py:12-40)
hm = gi.argmax_heatmaps()
preds, maxvals = inf.get_max_preds(hm)
save("argmax", preds=preds, maxvals=maxvals)

# ---- triangulators (triangulation.py:8-181)
u1, u2, P1, P2, Xtrue = gi.triangulation_case()
res = {}
for name in ("linear_eigen_triangulation", "linear_LS_triangulation", "iterative_LS_triangulation"):
    xs, sts = [], []
    for i in range(u1.shape[0]):
        x, st = getattr(tri, name)(u1[i], P1[i], u2[i], P2[i])
        xs.append(x)
        sts.append(np.asarray(st).astype(np.int64))
    res[name + "_x"] = np.asarray(xs)
    res[name + "_status"] = np.asarray(sts)
# exact (noise-free) projections: known-answer recovery
u1e, u2e = gi.exact_projections(P1, P2, Xtrue)
res["exact_eigen"] = np.asarray([tri.linear_eigen_triangulation(u1e[i], P1[i], u2e[i], P2[i])[0]
                                 for i in range(u1.shape[0])])
res["exact_iter"] = np.asarray([tri.iterative_LS_triangulation(u1e[i], P1[i], u2e[i], P2[i])[0]
                                for i in range(u1.shape[0])])
res["Xtrue"] = Xtrue
save("triangulation", **res)

# ---- patch -> image affine with rotation / scale (img_utils.py:141-155)
coords, boxes = gi.patch_case()
outs = [iu.trans_coords_from_patch_to_org_3d(coords[i], boxes[i, 0], boxes[i, 1], boxes[i, 2],
                                             boxes[i, 3], 256, 256, 2000, 2000, scale=boxes[i, 4],
                                             rot=boxes[i, 5]) for i in range(len(coords))]
save("patch_to_image", kps=np.asarray(outs))

# ---- self_supervision chain (img_utils.py:166-243)
logits_np, meta_np = gi.selfsup_case()
meta = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in meta_np.items()}
label, weight = iu.self_supervision(torch.from_numpy(logits_np), meta)
kp = il.get_joint_location_result(256, 256, torch.from_numpy(logits_np))
save("selfsup", label=label, weight=weight, result=kp)

# ---- PoseResNet forward/backward (pose3d_resnet.py) pins oracle/restate_net.py
for tag, c in gi.NET_CASES.items():
    cfg = refshim.make_cfg(num_layers=c["layers"], num_joints=c["J"], volume=c["volume"],
                           depth_res=c["D"], image_size=(c["HW"], c["HW"]))
    model = r.pose3d_resnet.get_pose_net(cfg, False)
    shapes = restate_net.param_shapes(num_layers=c["layers"], num_joints=c["J"], volume=c["volume"],
                                      depth_res=c["D"])
    sd = restate_net.init_state(shapes, c["seed"])
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(sd)
    model.train()
    x = torch.from_numpy(gi.images(c["N"], c["HW"], c["seed"]))
    out = model(x)
    outs = out if isinstance(out, tuple) else (out,)
    g = [torch.from_numpy(gi.grad_like(o.shape, c["seed"] + 1 + i)) for i, o in enumerate(outs)]
    sum((o * gg).sum() for o, gg in zip(outs, g)).backward()
    named = dict(model.named_parameters())
    rec = {"out%d" % i: o.detach().numpy() for i, o in enumerate(outs)}
    for k in c["grad_keys"]:
        rec["grad/" + k] = named[k].grad.numpy()
    rec["bn1.running_mean"] = model.state_dict()["bn1.running_mean"].numpy()
    rec["bn1.running_var"] = model.state_dict()["bn1.running_var"].numpy()
    model.eval()
    with torch.no_grad():
        e = model(x)
    rec["eval_out0"] = (e[0] if isinstance(e, tuple) else e).numpy()
    save("net_" + tag, **rec)
