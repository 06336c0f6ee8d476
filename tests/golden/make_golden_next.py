"""Golden vectors of the SURVEY.md 8(f) "next" rows built so far -- the H36M evaluation protocol
(row 2) and polynomial triangulation (row 3) -- produced by the UNMODIFIED reference:
lib/utils/triangulation.py::polynomial_triangulation (which calls cv2.correctMatches of the
installed OpenCV) on the seeded camera pairs of tests/golden_inputs.py::triangulation_case; lib/dataset/h36m.py::H36M_Integral.evaluate is called as an unbound
method on a stand-in `self` carrying only the fields it reads (db, cfg.DATASET.MPII_ORDER,
cfg.DEBUG.DEBUG, root), and lib/utils/prep_h36m.py::compute_similarity_transform directly.
Run in the build container only:
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_next.py"""
import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import refshim, restate  # noqa: E402
from tests import golden_inputs as gi  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
refshim.install()
h36m = importlib.import_module("lib.dataset.h36m")
prep = importlib.import_module("lib.utils.prep_h36m")
S = types.SimpleNamespace

pred, gt, pelvis, fl, c_p = gi.eval_case()
assert np.array_equal(h36m.H36M_TO_MPII_PERM, restate.H36M_TO_MPII_PERM)
rec = {}
for mpii in (False, True):
    db = [dict(fl=fl[i], c_p=c_p[i], pelvis=pelvis[i], joints_3d=gt[i],
               joints_3d_vis=np.ones((17, 3))) for i in range(len(pred))]
    fake = S(db=db, cfg=S(DATASET=S(MPII_ORDER=mpii), DEBUG=S(DEBUG=False)), root='')
    p = pred[:, h36m.H36M_TO_MPII_PERM, :] if mpii else pred
    with contextlib.redirect_stdout(io.StringIO()) as buf:       # evaluate prints the per-joint means
        name_value, mean = h36m.H36M_Integral.evaluate(fake, p.copy())
    tag = "mpii" if mpii else "h36m"
    rec[tag + "_values"] = np.array([v for _, v in name_value])
    rec[tag + "_mean"] = np.array(mean)
    rec[tag + "_per_joint"] = np.array([float(l.split()[-1]) for l in buf.getvalue().strip().splitlines()])
proc = [prep.compute_similarity_transform(gt[i], pred[i][:, :3], compute_optimal_scale=True)
        for i in range(4)]
rec["proc_d"] = np.array([q[0] for q in proc])
rec["proc_Z"] = np.stack([q[1] for q in proc])
rec["proc_T"] = np.stack([q[2] for q in proc])
rec["proc_b"] = np.array([q[3] for q in proc])
rec["proc_c"] = np.stack([q[4] for q in proc])
np.savez_compressed(os.path.join(OUT, "h36m_eval.npz"), **rec)
print("wrote h36m_eval", {k: v.shape for k, v in rec.items()})

# ---- polynomial (optimal) triangulation, triangulation.py:184-220
import cv2  # noqa: E402
tri = importlib.import_module("lib.utils.triangulation")
u1, u2, P1, P2, Xtrue = gi.triangulation_case()
xs, sts, c1s, c2s = [], [], [], []
for i in range(len(u1)):
    x, st = tri.polynomial_triangulation(u1[i], P1[i], u2[i], P2[i])
    xs.append(x)
    sts.append(st)
    F = restate.fundamental_from_projections(P1[i], P2[i])
    c1, c2 = cv2.correctMatches(F, u1[i].reshape(1, -1, 2), u2[i].reshape(1, -1, 2))
    c1s.append(c1[0])
    c2s.append(c2[0])
u1e, u2e = gi.exact_projections(P1, P2, Xtrue)
exact = np.asarray([tri.polynomial_triangulation(u1e[i], P1[i], u2e[i], P2[i])[0] for i in range(len(u1))])
np.savez_compressed(os.path.join(OUT, "triangulation_poly.npz"), x=np.asarray(xs),
                    status=np.asarray(sts).astype(np.int64), corrected_u1=np.asarray(c1s),
                    corrected_u2=np.asarray(c2s), exact=exact, cv2_version=np.array(cv2.__version__))
print("wrote triangulation_poly", np.asarray(xs).shape, "exact-recovery error",
      float(np.abs(exact - Xtrue).max()))

# ---- input pipeline: get_single_patch_sample (img_utils.py:246-298) on synthetic frames
import random  # noqa: E402
import tempfile  # noqa: E402
iu = importlib.import_module("lib.utils.img_utils")
il = importlib.import_module("lib.core.integral_loss")
MEAN = np.array([123.675, 116.280, 103.530])          # lib/dataset/JointIntegralDataset.py:67-68
STD = np.array([58.395, 57.120, 57.375])
rec = {}
with tempfile.TemporaryDirectory() as tmp:
    for tag in gi.PATCH_CASES:
        img, box, joints, joints_vis, pw, ph, seed = gi.frame_case(tag)
        path = os.path.join(tmp, tag + ".png")              # lossless: imread returns the same BGR bytes
        assert cv2.imwrite(path, img)
        for aug in (False, True):
            np.random.seed(seed); random.seed(seed)
            scale, rot, do_flip, color_scale = iu.do_augmentation() if aug else (1.0, 0, False, [1.0, 1.0, 1.0])
            np.random.seed(seed); random.seed(seed)          # the call below draws the same parameters
            patch, label, weight, s2, r2 = iu.get_single_patch_sample(
                path, box[0], box[1], box[2], box[3], joints.copy(), joints_vis.copy(), [], None, pw, ph,
                2000.0, 2000.0, MEAN, STD, aug, il.generate_joint_location_label)
            assert (s2, r2) == (scale, rot) and not do_flip
            k = tag + ("_aug" if aug else "")
            rec[k + "_patch"] = patch
            rec[k + "_label"] = np.asarray(label, dtype=np.float64)
            rec[k + "_weight"] = np.asarray(weight, dtype=np.float64)
            rec[k + "_aug"] = np.array([scale, rot, float(do_flip)] + list(color_scale), dtype=np.float64)
np.savez_compressed(os.path.join(OUT, "patch_sample.npz"), cv2_version=np.array(cv2.__version__),
                    numpy_version=np.array(np.__version__), **rec)
print("wrote patch_sample", {k: v.shape for k, v in rec.items() if k.endswith("_patch")})

# ---- refiner MLP (refiner/model.py), imported by path: forward / backward in training mode with
# p_dropout = 0 (torch's dropout stream cannot be reproduced) and the eval-mode forward
import importlib.util  # noqa: E402
import torch  # noqa: E402
from oracle import restate_refiner  # noqa: E402
spec = importlib.util.spec_from_file_location("ref_refiner_model", os.path.join(refshim.REF_ROOT, "refiner", "model.py"))
refm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(refm)
LS, NB = 128, 24
model = refm.LinearModelPG(linear_size=LS, p_dropout=0.0, input_size=45, output_size=45)
shapes = restate_refiner.param_shapes(LS, 45, 45)
sd = restate_refiner.init_state(shapes, 17)
assert list(model.state_dict().keys()) == list(sd.keys())
assert all(tuple(v.shape) == tuple(shapes[k]) for k, v in model.state_dict().items())
model.load_state_dict(sd)
model.train()
x = torch.from_numpy(gi.grad_like((NB, 45), 18)).requires_grad_(True)
tgt = torch.from_numpy(gi.grad_like((NB, 45), 19))
p1, p2 = model(x)
loss = torch.nn.functional.mse_loss(p1, tgt) + torch.nn.functional.mse_loss(p2, tgt)      # refiner/main.py:50
loss.backward()
rec = {"p1": p1.detach().numpy(), "p2": p2.detach().numpy(), "loss": loss.detach().numpy(), "dx": x.grad.numpy()}
named = dict(model.named_parameters())
for k in ("w1.weight", "w2.weight", "w4.bias", "linear_stages.0.w3.weight", "linear_stages.1.batch_norm2.weight",
          "linear_stages.1.batch_norm4.bias", "batch_norm3.weight"):
    rec["grad/" + k] = named[k].grad.numpy()
rec["batch_norm1.running_var"] = model.state_dict()["batch_norm1.running_var"].numpy()
model.eval()
with torch.no_grad():
    e1, e2 = model(x.detach())
rec["eval_p1"], rec["eval_p2"] = e1.numpy(), e2.numpy()
np.savez_compressed(os.path.join(OUT, "refiner.npz"), **rec)
print("wrote refiner", {k: v.shape for k, v in rec.items()})
