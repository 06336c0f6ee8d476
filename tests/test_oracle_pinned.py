"""CPU: pins the oracle restatements (oracle/restate.py, oracle/restate_net.py)
against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  The reference ships no tests of its own
(SURVEY.md section 4), so these vectors are the pin."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import restate, restate_net
from tests import golden_inputs as gi
from tests.conftest import ROOT, relerr


@pytest.mark.parametrize("tag", list(gi.SOFTARGMAX_CASES))
def test_softargmax_and_losses(golden, tag):
    N, J, D, H, W, seed, scale = gi.SOFTARGMAX_CASES[tag]
    g = golden("softargmax_" + tag)
    logits = gi.logits(N, J, D, H, W, seed, scale)
    coords = restate.softmax_integral(logits, J, W, H, D)
    assert np.max(np.abs(coords - g["coords"])) <= 2e-6          # fp32 accumulation order only
    gt, wt = gi.labels(N, J, seed)
    for kind, key in (("l1", "l1"), ("smoothl1", "smoothl1")):
        for norm in (False, True):
            k = key + ("_norm" if norm else "")
            loss, dcoords = restate.weighted_loss(kind, g["coords"], gt, wt, True, norm)
            assert abs(loss - float(g[k + "_loss"])) <= 1e-5 * max(1.0, abs(loss))
            grad = restate.softmax_integral_grad(logits, dcoords, J, W, H, D)
            ref_sample = g[k + "_grad_sample"]
            assert relerr(grad[:, :, ::3, ::3], ref_sample) <= 2e-4
            assert relerr(np.abs(grad).sum((2, 3)), g[k + "_grad_sum_abs"]) <= 2e-4
    loss, _ = restate.weighted_loss("mse", g["coords"], gt, wt, True, False)
    assert abs(loss - float(g["mse_loss"])) <= 1e-5 * max(1.0, abs(loss))
    if D == W:
        res = restate.joint_location_result(256, 256, g["coords"])
        assert np.max(np.abs(res - g["result"])) <= 1e-4


def test_argmax(golden):
    g = golden("argmax")
    preds, maxvals, _ = restate.get_max_preds(gi.argmax_heatmaps())
    assert np.array_equal(preds, g["preds"])                     # bit-exact indices
    assert np.array_equal(maxvals, g["maxvals"])


def test_triangulators(golden):
    g = golden("triangulation")
    u1, u2, P1, P2, X = gi.triangulation_case()
    for name, fn in (("linear_eigen_triangulation", restate.linear_eigen_triangulation),
                     ("linear_LS_triangulation", restate.linear_LS_triangulation),
                     ("iterative_LS_triangulation", restate.iterative_LS_triangulation)):
        for i in range(len(u1)):
            x, st = fn(u1[i], P1[i], u2[i], P2[i])
            assert np.max(np.abs(x - g[name + "_x"][i])) <= 1e-6, name   # mm; bar is 1e-4
            assert np.array_equal(np.asarray(st).astype(np.int64), g[name + "_status"][i])
    u1e, u2e = gi.exact_projections(P1, P2, X)
    for i in range(len(u1)):   # known-answer: noise-free projections triangulate back
        assert np.max(np.abs(restate.linear_eigen_triangulation(u1e[i], P1[i], u2e[i], P2[i])[0] - X[i])) <= 1e-7
        assert np.max(np.abs(restate.iterative_LS_triangulation(u1e[i], P1[i], u2e[i], P2[i])[0] - X[i])) <= 1e-7
    assert np.max(np.abs(g["exact_eigen"] - X)) <= 1e-7 and np.max(np.abs(g["exact_iter"] - X)) <= 1e-7


def test_patch_to_image(golden):
    g = golden("patch_to_image")
    coords, boxes = gi.patch_case()
    for i in range(len(coords)):
        out = restate.trans_coords_from_patch_to_org_3d(coords[i], *boxes[i, :4], 256, 256, 2000, 2000,
                                                        scale=boxes[i, 4], rot=boxes[i, 5])
        assert np.max(np.abs(out - g["kps"][i])) <= 1e-9


def test_self_supervision_chain(golden):
    g = golden("selfsup")
    logits, meta = gi.selfsup_case()
    J, D = 4, 16
    coords = restate.softmax_integral(logits, J, D, D, D)
    label, weight, X, img = restate.self_supervision(coords, meta)
    assert np.max(np.abs(label - g["label"])) <= 5e-6
    assert np.array_equal(weight, g["weight"])


@pytest.mark.parametrize("tag", list(gi.NET_CASES))
def test_network_restatement(golden, tag):
    c = gi.NET_CASES[tag]
    g = golden("net_" + tag)
    shapes = restate_net.param_shapes(num_layers=c["layers"], num_joints=c["J"], volume=c["volume"],
                                      depth_res=c["D"])
    sd = restate_net.init_state(shapes, c["seed"])
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
         for k, v in sd.items()}
    x = torch.from_numpy(gi.images(c["N"], c["HW"], c["seed"]))
    new_stats = {}
    out = restate_net.forward(p, x, num_layers=c["layers"], volume=c["volume"],
                              image_size=(c["HW"], c["HW"]), training=True, new_stats=new_stats)
    outs = out if isinstance(out, tuple) else (out,)
    for i, o in enumerate(outs):
        assert relerr(o.detach().numpy(), g["out%d" % i]) <= 5e-4   # fp32 reassociation (tiny-batch BN)
    gs = [torch.from_numpy(gi.grad_like(o.shape, c["seed"] + 1 + i)) for i, o in enumerate(outs)]
    sum((o * gg).sum() for o, gg in zip(outs, gs)).backward()
    # heads are well conditioned; deep-trunk gradients of a random-init net are chaotic in fp32
    for k in ("final_layer.bias",):
        assert relerr(p[k].grad.numpy(), g["grad/" + k]) <= 1e-3
    assert relerr(new_stats["bn1.running_mean"].numpy(), g["bn1.running_mean"]) <= 1e-5
    assert relerr(new_stats["bn1.running_var"].numpy(), g["bn1.running_var"]) <= 1e-5
    sd_eval = dict(sd)
    sd_eval.update(new_stats)          # the reference's eval pass ran after the running-stat update
    with torch.no_grad():
        e = restate_net.forward(sd_eval, x, num_layers=c["layers"], volume=c["volume"],
                                image_size=(c["HW"], c["HW"]), training=False)
    e = e[0] if isinstance(e, tuple) else e
    assert relerr(e.numpy(), g["eval_out0"]) <= 5e-4


def test_h36m_evaluation_protocol(golden):
    """oracle/restate.py::h36m_evaluate / compute_similarity_transform against the outputs of
    the unmodified H36M_Integral.evaluate (tests/golden/make_golden_next.py)."""
    g = golden("h36m_eval")
    pred, gt, pelvis, fl, c_p = gi.eval_case()
    for mpii, tag in ((False, "h36m"), (True, "mpii")):
        p = pred[:, restate.H36M_TO_MPII_PERM, :] if mpii else pred
        o = restate.h36m_evaluate(p, gt, pelvis[:, 2], fl, c_p, mpii_order=mpii)
        vals = np.array([v for _, v in o["name_value"]])
        assert np.max(np.abs(vals - g[tag + "_values"])) <= 1e-9
        assert abs(o["mean"] - float(g[tag + "_mean"])) <= 1e-9
        assert np.max(np.abs(o["per_joint"].mean(0) - g[tag + "_per_joint"])) <= 1e-9
    for i in range(4):
        d, Z, T, b, c = restate.compute_similarity_transform(gt[i], pred[i][:, :3], True)
        assert abs(d - g["proc_d"][i]) <= 1e-12 and abs(b - g["proc_b"][i]) <= 1e-12
        assert np.max(np.abs(Z - g["proc_Z"][i])) <= 1e-9 and np.max(np.abs(T - g["proc_T"][i])) <= 1e-12
        assert np.max(np.abs(c - g["proc_c"][i])) <= 1e-9


def test_polynomial_triangulation(golden):
    """restate.correct_matches / polynomial_triangulation against cv2.correctMatches and the
    unmodified reference polynomial_triangulation (tests/golden/make_golden_next.py)."""
    g = golden("triangulation_poly")
    u1, u2, P1, P2, X = gi.triangulation_case()
    for i in range(len(u1)):
        F = restate.fundamental_from_projections(P1[i], P2[i])
        n1, n2 = restate.correct_matches(F, u1[i], u2[i])
        assert np.max(np.abs(n1 - g["corrected_u1"][i])) <= 1e-9          # px
        assert np.max(np.abs(n2 - g["corrected_u2"][i])) <= 1e-9
        x, st = restate.polynomial_triangulation(u1[i], P1[i], u2[i], P2[i])
        assert np.max(np.abs(x - g["x"][i])) <= 1e-6                      # mm
        assert np.array_equal(np.asarray(st).astype(np.int64), g["status"][i])
        # corrected matches satisfy the epipolar constraint exactly
        h1 = np.concatenate([n1, np.ones((len(n1), 1))], 1)
        h2 = np.concatenate([n2, np.ones((len(n2), 1))], 1)
        assert np.max(np.abs(np.einsum("ni,ij,nj->n", h2, F, h1))) <= 1e-9 * np.abs(F).max() * 1e6


MEAN = np.array([123.675, 116.280, 103.530])          # reference lib/dataset/JointIntegralDataset.py:67-68
STD = np.array([58.395, 57.120, 57.375])


@pytest.mark.parametrize("tag", list(gi.PATCH_CASES))
def test_input_pipeline_bit_exact(golden, tag):
    """restate.patch_sample (warpAffine fixed-point restatement, getAffineTransform LU) against the
    unmodified get_single_patch_sample: patches BIT-EXACT, labels to rounding."""
    g = golden("patch_sample")
    img, box, joints, vis, pw, ph, seed = gi.frame_case(tag)
    for aug in (False, True):
        k = tag + ("_aug" if aug else "")
        sc, rot, fl, c0, c1, c2 = g[k + "_aug"]
        t, lab, wt, tr = restate.patch_sample(img, box[0], box[1], box[2], box[3], joints, vis, pw, ph, 2000.0,
                                              MEAN, STD, sc, rot, bool(fl), (c0, c1, c2))
        assert np.array_equal(t, g[k + "_patch"]), k
        assert np.max(np.abs(lab - g[k + "_label"])) <= 1e-12
        assert np.array_equal(wt, g[k + "_weight"])


def test_nview_dlt_reduces_to_the_reference_pair_case(golden):
    """the V-view DLT oracle with V = 2 is the reference's linear_eigen_triangulation (golden);
    with four exact views it recovers the 3-D points."""
    g = golden("triangulation")
    u1, u2, P1, P2, X = gi.triangulation_case()
    for i in range(len(u1)):
        x, st = restate.linear_eigen_triangulation_nview(np.stack([u1[i], u2[i]]), np.stack([P1[i], P2[i]]))
        assert np.max(np.abs(x - g["linear_eigen_triangulation_x"][i])) <= 1e-6
    rng = np.random.default_rng(3)
    R, T, f, c, P = restate.synthetic_cameras(rng, 2, 4)
    Xw = rng.normal(0, 400, (17, 3))
    us = np.stack([restate.project(P[0, v], Xw) for v in range(4)])
    x, st = restate.linear_eigen_triangulation_nview(us, P[0])
    assert np.max(np.abs(x - Xw)) <= 1e-8 and st.all()


def test_refiner_oracle(golden):
    """oracle/restate_refiner.py against the unmodified refiner/model.py (training mode without
    dropout: forward, input / parameter gradients, running statistics; eval-mode forward)."""
    from oracle import restate_refiner as rr
    g = golden("refiner")
    sd = rr.init_state(rr.param_shapes(128, 45, 45), 17)
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    x = torch.from_numpy(gi.grad_like((24, 45), 18)).requires_grad_(True)
    tgt = torch.from_numpy(gi.grad_like((24, 45), 19))
    stats = {}
    p1, p2 = rr.forward(p, x, training=True, new_stats=stats)
    loss = torch.nn.functional.mse_loss(p1, tgt) + torch.nn.functional.mse_loss(p2, tgt)
    loss.backward()
    assert relerr(p1.detach().numpy(), g["p1"]) <= 1e-5 and relerr(p2.detach().numpy(), g["p2"]) <= 1e-5
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert relerr(x.grad.numpy(), g["dx"]) <= 1e-4
    for k in [k[5:] for k in g if k.startswith("grad/")]:
        assert relerr(p[k].grad.numpy(), g["grad/" + k]) <= 1e-4, k
    assert relerr(stats["batch_norm1.running_var"].numpy(), g["batch_norm1.running_var"]) <= 1e-5
    sd_after = dict(sd)
    sd_after.update(stats)                   # the reference evaluates after its training-mode forward
    with torch.no_grad():
        e1, e2 = rr.forward(sd_after, x.detach(), training=False)
    assert relerr(e1.numpy(), g["eval_p1"]) <= 1e-5 and relerr(e2.numpy(), g["eval_p2"]) <= 1e-5


def test_final_preds_bit_exact(golden):
    """get_final_preds (lib/core/inference.py:43-68): argmax, +-0.25 px refinement and the
    cv2.getAffineTransform-based transform_preds, restated -- bit-equal to the reference run."""
    g = golden("final_preds")
    hm, center, scale = gi.final_preds_case()
    for pp in (1, 0):
        p, m = restate.final_preds(hm, center, scale, bool(pp))
        assert np.array_equal(p, g["preds_pp%d" % pp])
        assert np.array_equal(m, g["maxvals_pp%d" % pp])


def test_eight_point_fallback(golden):
    """The fallback of polynomial_triangulation (lib/utils/triangulation.py:213-217):
    restate.fundamental_8point against cv2.findFundamentalMat(FM_8POINT), the branch on its own
    against the same composition of reference / OpenCV calls, and the natural trigger
    (P2 == P1: the correction is all-NaN, the reference itself falls back)."""
    g = golden("triangulation_8point")
    u1, u2, P1, P2, X = gi.triangulation_case()
    for i in range(len(u1)):
        F = restate.fundamental_8point(u1[i], u2[i])
        assert np.max(np.abs(F - g["f8"][i])) <= 1e-10 * np.abs(g["f8"][i]).max()
        x, st = restate.polynomial_triangulation_8point(u1[i], P1[i], u2[i], P2[i])
        assert np.max(np.abs(x - g["x_8pt"][i])) <= 1e-4
        x, st = restate.polynomial_triangulation(u1[i], P1[i], u2[i], P1[i])
        assert np.max(np.abs(x - g["x_same"][i])) <= 1e-4 and np.array_equal(st.astype(np.int64), g["st_same"][i])


def test_occluder_paste_bit_exact(golden):
    """Synthetic-occlusion augmentation (lib/utils/augmentation.py:61-123 inside
    get_single_patch_sample, img_utils.py:269-270): the mirror's draws (draw_occluders: same
    np.random / random calls, cv2.resize) + restate.paste_over reproduce the unmodified
    reference's patches BIT-EXACTLY."""
    import random
    cv2 = pytest.importorskip("cv2")
    sys.path.insert(0, os.path.join(ROOT, "epipolarpose_b200"))
    import lib.utils.img_utils as iu
    from lib.utils.augmentation import draw_occluders
    g = golden("patch_occluders")
    occ = gi.occluder_set()
    mean, std = np.array([123.675, 116.280, 103.530]), np.array([58.395, 57.120, 57.375])
    for tag in gi.PATCH_CASES:
        img, box, joints, vis, pw, ph, seed = gi.frame_case(tag)
        for aug in (False, True):
            np.random.seed(seed + 7); random.seed(seed + 7)
            scale, rot, fl, cs = iu.do_augmentation() if aug else (1.0, 0, False, [1.0, 1.0, 1.0])
            lst = draw_occluders(pw, ph, occ)
            t, lab, wt, tr = restate.patch_sample(img, box[0], box[1], box[2], box[3], joints, vis, pw, ph,
                                                  2000.0, mean, std, scale, rot, fl, cs, occluders=lst)
            k = tag + ("_aug" if aug else "")
            assert np.array_equal(t, g[k + "_patch"]), k
            assert np.max(np.abs(lab - g[k + "_label"])) <= 1e-9
