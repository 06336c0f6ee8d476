"""GPU (-m gpu): parity of the CUDA path (through the C ABI / the reference-
shaped Python surface) against the oracle restatement and the golden vectors
the unmodified reference produced.  Tolerances per BASELINE.json north_star:
argmax indices bit-exact; triangulated joints <= 1e-4 mm; soft-argmax coords
<= 1e-5 abs; heatmaps / gradients <= 1e-3 rel (max|d|/max|ref| per tensor)."""
import numpy as np
import pytest
import torch

from oracle import restate, restate_net
from tests import golden_inputs as gi
from tests.conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from epipolarpose_b200 import ops
    ops.device_check()
    return torch.device("cuda:0")


# ------------------------------------------------------------------ soft-argmax + losses
@pytest.mark.parametrize("tag", list(gi.SOFTARGMAX_CASES))
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_softargmax_loss_golden(golden, dev, tag, layout):
    import lib.core.integral_loss as il
    N, J, D, H, W, seed, scale = gi.SOFTARGMAX_CASES[tag]
    g = golden("softargmax_" + tag)
    x = torch.from_numpy(gi.logits(N, J, D, H, W, seed, scale)).to(dev)
    if layout == "nhwc":
        x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)   # channels_last view
    x.requires_grad_(True)
    coords = il.softmax_integral_tensor(x, J, True, W, H, D)
    assert np.max(np.abs(coords.detach().cpu().numpy() - g["coords"])) <= 1e-5
    gt, wt = gi.labels(N, J, seed)
    gt, wt = torch.from_numpy(gt).to(dev), torch.from_numpy(wt).to(dev)
    for cls, key in ((il.L1JointLocationLoss, "l1"), (il.SmoothL1JointLocationLoss, "smoothl1")):
        for norm in (False, True):
            x.grad = None
            loss = cls(J, norm=norm)(x, gt, wt)
            loss.backward()
            k = key + ("_norm" if norm else "")
            assert abs(loss.item() - float(g[k + "_loss"])) <= 1e-5 * max(1.0, abs(float(g[k + "_loss"])))
            grad = x.grad.cpu().numpy()
            assert relerr(grad[:, :, ::3, ::3], g[k + "_grad_sample"]) <= 1e-3
            assert relerr(np.abs(grad).sum((2, 3)), g[k + "_grad_sum_abs"]) <= 1e-3
    if D == W:
        res = il.get_joint_location_result(256, 256, x.detach())
        assert np.max(np.abs(res - g["result"])) <= 256 * 1e-5


def test_softargmax_full_size_properties(dev):
    """BASELINE size (J=17, 64^3) on a few images: planted delta peaks decode to
    their voxel; uniform logits decode to the volume centre; gradients of each
    (n,j) volume sum to zero (softmax Jacobian annihilates constants)."""
    import lib.core.integral_loss as il
    N, J, D = 4, 17, 64
    rng = np.random.default_rng(5)
    x = torch.zeros((N, J * D, D, D), device=dev)
    pos = rng.integers(0, D, size=(N, J, 3))
    for n in range(N):
        for j in range(J):
            x[n, j * D + pos[n, j, 2], pos[n, j, 1], pos[n, j, 0]] = 60.0
    for view in (x, x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)):
        c = il.softmax_integral_tensor(view, J, True, D, D, D).cpu().numpy().reshape(N, J, 3)
        assert np.max(np.abs((c + 0.5) * D - pos)) <= 1e-3
    u = torch.zeros((1, J * D, D, D), device=dev)
    c = il.softmax_integral_tensor(u, J, True, D, D, D).cpu().numpy()
    assert np.max(np.abs(c - ((D - 1) / 2 / D - 0.5))) <= 1e-5
    y = (2 * torch.randn((2, J * D, D, D), device=dev)).requires_grad_(True)
    il.softmax_integral_tensor(y, J, True, D, D, D).sum().backward()
    s = y.grad.reshape(2, J, -1).sum(-1).abs().max().item()
    assert s <= 1e-5


def test_softargmax_vs_oracle_medium(dev):
    import lib.core.integral_loss as il
    N, J, D = 2, 17, 32
    logits = gi.logits(N, J, D, D, D, 77, 3.0)
    ref = restate.softmax_integral(logits, J, D, D, D)
    for lay in (0, 1):
        x = torch.from_numpy(logits).to(dev)
        if lay:
            x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        c = il.softmax_integral_tensor(x, J, True, D, D, D).cpu().numpy()
        assert np.max(np.abs(c - ref)) <= 1e-5
        g = np.random.default_rng(3).standard_normal((N, J * 3)).astype(np.float32)
        x.requires_grad_(True)
        il.softmax_integral_tensor(x, J, True, D, D, D).backward(torch.from_numpy(g).to(dev))
        gref = restate.softmax_integral_grad(logits, g, J, D, D, D)
        assert relerr(x.grad.cpu().numpy(), gref) <= 1e-3


@pytest.mark.parametrize("shape", [(2, 3, 6, 5, 7), (1, 17, 10, 9, 13), (2, 2, 5, 8, 8)])
@pytest.mark.parametrize("memory", ["nchw", "channels_last", "sliced"])
def test_softargmax_any_volume_shape(dev, shape, memory):
    """The reference accepts every J/D/H/W and any memory format (integral_loss.py:71-86): widths that are
    not a multiple of 4, depths that are not (channels_last falls back to the NCHW kernels), and a
    mis-aligned slice take the scalar-load kernels; same 1e-5 / 1e-3 bars as the vector paths."""
    import lib.core.integral_loss as il
    N, J, D, H, W = shape
    logits = gi.logits(N, J, D, H, W, 91, 3.0)
    ref = restate.softmax_integral(logits, J, W, H, D)
    x = torch.from_numpy(logits).to(dev)
    if memory == "channels_last":
        x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    elif memory == "sliced":
        big = torch.zeros(N * J * D * H * W + 1, device=dev)
        big[1:] = x.reshape(-1)
        x = big[1:].view(N, J * D, H, W)                    # 4-byte aligned only
    x.requires_grad_(True)
    c = il.softmax_integral_tensor(x, J, True, W, H, D)
    assert np.max(np.abs(c.detach().cpu().numpy() - ref)) <= 1e-5
    g = np.random.default_rng(4).standard_normal((N, J * 3)).astype(np.float32)
    c.backward(torch.from_numpy(g).to(dev))
    gref = restate.softmax_integral_grad(logits, g, J, W, H, D)
    assert relerr(x.grad.cpu().numpy(), gref) <= 1e-3


# ------------------------------------------------------------------ argmax
@pytest.mark.parametrize("shape", [(3, 5, 16, 24), (2, 17, 64, 64), (1, 2, 5, 7)])
@pytest.mark.parametrize("kind", ["l1", "smoothl1", "mse"])
def test_heatmap_joint_loss_vs_oracle(dev, shape, kind):
    """epb_heatmap_joint_loss (one launch: heat-map MSE + joint loss, value and both gradients)
    against the float64 oracle; losses <= 1e-5 rel, gradients <= 1e-5 of the tensor maximum;
    deterministic across repeated launches.  (1,2,5,7): HW not a multiple of 4 (scalar path)."""
    import lib.core.integral_loss as il
    N, J, H, W = shape
    hm, tgt, wh, x, t, w = gi.heatmap_case(N, J, H, W, 72 + N)
    o_hm, o_jt, o_tot, o_dhm, o_dx = restate.heatmap_joint_loss(hm, tgt, wh, x, t, w, kind, 0.5, 2.0)
    th = torch.from_numpy(hm).to(dev).requires_grad_(True)
    tx = torch.from_numpy(x).to(dev).requires_grad_(True)
    crit = il.HeatmapJointLoss(J, kind=kind, hm_scale=0.5, jt_scale=2.0)
    args = (torch.from_numpy(tgt).to(dev), torch.from_numpy(t).to(dev), torch.from_numpy(w).to(dev))
    tot = crit((th, tx), *args, hm_weight=torch.from_numpy(wh).to(dev))
    tot.backward()
    assert abs(tot.item() - o_tot) <= 1e-5 * abs(o_tot)
    assert abs(crit.last_parts[0].item() - o_hm) <= 1e-5 * o_hm
    assert abs(crit.last_parts[1].item() - o_jt) <= 1e-5 * max(o_jt, 1e-30)
    assert relerr(th.grad.cpu().numpy(), o_dhm) <= 1e-5
    assert relerr(tx.grad.cpu().numpy(), o_dx) <= 1e-5
    tot2 = crit((th, tx), *args, hm_weight=torch.from_numpy(wh).to(dev))
    assert tot2.item() == tot.item()
    # heat-map loss alone, unweighted == F.mse_loss
    l2 = il.HeatmapMSELoss()(th, args[0])
    ref = float(((hm.astype(np.float64) - tgt) ** 2).mean())
    assert abs(l2.item() - ref) <= 1e-5 * ref


@pytest.mark.parametrize("mpii", [False, True])
def test_h36m_eval_vs_oracle_and_reference(golden, dev, mpii):
    """epb_h36m_eval (back-projection, Procrustes with optimal scale, root alignment, protocol
    means) against the numpy oracle per sample (<= 1e-8 mm) and against the aggregate values the
    unmodified H36M_Integral.evaluate produced (<= 1e-8 mm); PCK flags bit-exact; a larger batch
    through size-independent properties (alignment removes any similarity transform)."""
    import lib.dataset.h36m_eval as he
    g = golden("h36m_eval")
    tag = "mpii" if mpii else "h36m"
    pred, gt, pelvis, fl, c_p = gi.eval_case()
    p = pred[:, he.H36M_TO_MPII_PERM, :] if mpii else pred
    nv, perf, det = he.evaluate_h36m(p, gt, pelvis, fl, c_p, mpii_order=mpii, return_poses=True)
    o = restate.h36m_evaluate(p, gt, pelvis[:, 2], fl, c_p, mpii_order=mpii)
    assert np.max(np.abs(det["metrics"] - o["metrics"])) <= 1e-8
    assert np.max(np.abs(det["per_joint"] - o["per_joint"])) <= 1e-8
    assert np.array_equal(det["pck"], o["pck"])
    assert np.max(np.abs(det["poses"] - o["poses"])) <= 1e-8
    assert np.max(np.abs(np.array([v for _, v in nv]) - g[tag + "_values"])) <= 1e-8
    assert abs(perf - float(g[tag + "_mean"])) <= 1e-8
    # property at scale: predictions that are an exact similarity transform of the ground truth
    # in camera space align to zero error (4096 samples)
    if not mpii:
        rng = np.random.default_rng(5)
        S, J = 4096, 17
        pb, gb, pel, f2, c2 = gi.eval_case(S, J, 83)
        X = np.zeros((S, J, 3))
        d = gb[:, :, 2] + pel[:, 2:3]
        X[:, :, 0] = (gb[:, :, 0] - c2[:, 0:1]) / f2[:, 0:1] * d
        X[:, :, 1] = (gb[:, :, 1] - c2[:, 1:2]) / f2[:, 1:2] * d
        X[:, :, 2] = d
        ang = rng.uniform(-0.3, 0.3, S)
        R = np.stack([np.stack([np.cos(ang), -np.sin(ang), 0 * ang], 1),
                      np.stack([np.sin(ang), np.cos(ang), 0 * ang], 1),
                      np.stack([0 * ang, 0 * ang, 1 + 0 * ang], 1)], 1)
        Y = 1.1 * np.einsum("sjk,skl->sjl", X - X[:, :1], R) + X[:, :1]
        pp = np.zeros((S, J, 3))
        pp[:, :, 0] = Y[:, :, 0] / Y[:, :, 2] * f2[:, 0:1] + c2[:, 0:1]
        pp[:, :, 1] = Y[:, :, 1] / Y[:, :, 2] * f2[:, 1:2] + c2[:, 1:2]
        pp[:, :, 2] = Y[:, :, 2] - pel[:, 2:3]
        _, _, dd = he.evaluate_h36m(pp, gb, pel, f2, c2)
        assert np.max(dd["metrics"][:, 1]) <= 1e-6          # aligned error vanishes
        assert np.min(dd["metrics"][:, 0]) > 1.0            # un-aligned error does not


MEAN = np.array([123.675, 116.280, 103.530])          # reference lib/dataset/JointIntegralDataset.py:67-68
STD = np.array([58.395, 57.120, 57.375])


def test_input_pipeline_bit_exact(golden, dev):
    """epb_patch_sample / epb_patch_joints through the reference-named get_single_patch_sample:
    patches BIT-EXACT against the unmodified reference (cv2.warpAffine INTER_LINEAR + colour scale +
    normalisation), labels to rounding, same augmentation draws; batched launch with frames of
    different sizes and a mirrored frame against the oracle; a full-size batch (64 frames of
    1000x1002 -> 256x256) spot-checked bit-exactly against the oracle."""
    import random
    import lib.utils.img_utils as iu
    g = golden("patch_sample")
    for tag in gi.PATCH_CASES:
        img, box, joints, vis, pw, ph, seed = gi.frame_case(tag)
        for aug in (False, True):
            k = tag + ("_aug" if aug else "")
            np.random.seed(seed); random.seed(seed)
            patch, label, weight, scale, rot = iu.get_single_patch_sample(
                img, box[0], box[1], box[2], box[3], joints.copy(), vis.copy(), [], None, pw, ph, 2000.0, 2000.0,
                MEAN, STD, aug, None)
            assert (scale, rot) == (g[k + "_aug"][0], g[k + "_aug"][1])
            assert np.array_equal(patch, g[k + "_patch"]), k
            assert np.max(np.abs(label - g[k + "_label"])) <= 1e-12
            assert np.array_equal(weight, g[k + "_weight"])
    a, b = gi.frame_case("noise64"), gi.frame_case("edge48")
    out, trans, _ = iu.generate_patch_batch_device(
        [a[0], b[0], a[0]], [a[1][0], b[1][0], a[1][0]], [a[1][1], b[1][1], a[1][1]], [a[1][2], b[1][2], a[1][2]],
        [a[1][3], b[1][3], a[1][3]], 48, 48, scale=[1.1, 1.0, 0.9], rot=[12.0, 0.0, -20.0],
        do_flip=[False, False, True], color_scale=[[1.1, 0.9, 1.0]] * 3, mean=MEAN, std=STD)
    out = out.cpu().numpy()
    for i, (c, sc, rot, fl) in enumerate(((a, 1.1, 12.0, False), (b, 1.0, 0.0, False), (a, 0.9, -20.0, True))):
        t, _, _, tr = restate.patch_sample(c[0], c[1][0], c[1][1], c[1][2], c[1][3], c[2], c[3], 48, 48, 2000.0,
                                           MEAN, STD, sc, rot, fl, (1.1, 0.9, 1.0))
        assert np.array_equal(out[i], t), i
        assert np.array_equal(trans[i].cpu().numpy(), tr)
    rng = np.random.default_rng(7)
    frames = [a[0]] * 64
    cx, cy = 500 + rng.uniform(-50, 50, 64), 500 + rng.uniform(-50, 50, 64)
    w, h = 800 + rng.uniform(-100, 100, 64), 800 + rng.uniform(-100, 100, 64)
    sc, rot = 1 + rng.uniform(-0.25, 0.25, 64), rng.uniform(-60, 60, 64)
    big, _, _ = iu.generate_patch_batch_device(frames, cx, cy, w, h, 256, 256, scale=sc, rot=rot, mean=MEAN, std=STD)
    for i in (0, 31, 63):
        t, _, _, _ = restate.patch_sample(a[0], cx[i], cy[i], w[i], h[i], a[2], a[3], 256, 256, 2000.0, MEAN, STD,
                                          sc[i], rot[i])
        assert np.array_equal(big[i].cpu().numpy(), t), i


def test_occluder_paste_bit_exact(golden, dev):
    """Synthetic-occlusion augmentation (lib/utils/augmentation.py:61-123) fused into the patch
    kernel (epb_patch_sample_occ): get_single_patch_sample(..., occluder=...) BIT-EXACT against
    the unmodified reference (same draws, cv2.resize of the occluders on the host, blend on the
    device); occlude_with_objects / paste_over in their numpy form; a batch of 64 samples with
    1..7 occluders each against the oracle."""
    import random
    pytest.importorskip("cv2")
    import lib.utils.img_utils as iu
    import lib.utils.augmentation as aug_m
    g = golden("patch_occluders")
    occ = gi.occluder_set()
    for tag in gi.PATCH_CASES:
        img, box, joints, vis, pw, ph, seed = gi.frame_case(tag)
        for aug in (False, True):
            k = tag + ("_aug" if aug else "")
            np.random.seed(seed + 7); random.seed(seed + 7)
            patch, label, weight, scale, rot = iu.get_single_patch_sample(
                img, box[0], box[1], box[2], box[3], joints.copy(), vis.copy(), [], None, pw, ph, 2000.0, 2000.0,
                MEAN, STD, aug, None, occluder=occ)
            assert np.array_equal(patch, g[k + "_patch"]), k
            assert np.max(np.abs(label - g[k + "_label"])) <= 1e-12
    # numpy-form API: occlude_with_objects / paste_over == the oracle's paste_over
    rng = np.random.default_rng(11)
    im = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    np.random.seed(5); random.seed(5)
    lst = aug_m.draw_occluders(128, 96, occ)
    np.random.seed(5); random.seed(5)
    got = aug_m.occlude_with_objects(im, occ)
    want = im.copy()
    for rgba, c in lst:
        restate.paste_over(rgba, want, np.asarray(c, dtype=np.float64))
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    dst = im.copy()
    aug_m.paste_over(occ[0], dst, np.array([10.4, 90.6]))                 # partly outside the image
    want = im.copy()
    restate.paste_over(occ[0], want, np.array([10.4, 90.6]))
    assert np.array_equal(dst, want)
    # a loader-sized batch
    a = gi.frame_case("noise64")
    B = 64
    cx, cy = 500 + rng.uniform(-50, 50, B), 500 + rng.uniform(-50, 50, B)
    w, h = 800 + rng.uniform(-100, 100, B), 800 + rng.uniform(-100, 100, B)
    sc, rot = 1 + rng.uniform(-0.25, 0.25, B), rng.uniform(-60, 60, B)
    np.random.seed(9); random.seed(9)
    per = [aug_m.draw_occluders(256, 256, occ) for _ in range(B)]
    big, _, _ = iu.generate_patch_batch_device([a[0]] * B, cx, cy, w, h, 256, 256, scale=sc, rot=rot, mean=MEAN,
                                               std=STD, occluders=per)
    for i in (0, 17, 63):
        t, _, _, _ = restate.patch_sample(a[0], cx[i], cy[i], w[i], h[i], a[2], a[3], 256, 256, 2000.0, MEAN, STD,
                                          sc[i], rot[i], occluders=per[i])
        assert np.array_equal(big[i].cpu().numpy(), t), i


def test_final_preds_bit_exact(golden, dev):
    """lib/core/inference.py:43-68 on the device (epb_final_preds) against the unmodified
    reference: coordinates bit-exact (float32), through the reference-shaped numpy API and
    the tensor API."""
    import types
    import lib.core.inference as inf
    g = golden("final_preds")
    hm, center, scale = gi.final_preds_case()
    for pp in (1, 0):
        cfg = types.SimpleNamespace(TEST=types.SimpleNamespace(POST_PROCESS=bool(pp)))
        preds, maxvals = inf.get_final_preds(cfg, hm.copy(), center, scale)
        assert preds.dtype == np.float32 and np.array_equal(preds, g["preds_pp%d" % pp])
        assert np.array_equal(maxvals, g["maxvals_pp%d" % pp])
        pd, md = inf.get_final_preds_device(torch.from_numpy(hm).to(dev), center, scale, bool(pp))
        assert np.array_equal(pd.cpu().numpy(), g["preds_pp%d" % pp])
    # a batch the size of a validation step: equals the oracle on every map
    rng = np.random.default_rng(3)
    big = rng.standard_normal((64, 16, 64, 64)).astype(np.float32)
    c = np.stack([500 + rng.uniform(-50, 50, 64), 500 + rng.uniform(-50, 50, 64)], 1)
    s = np.stack([4 + rng.uniform(-1, 1, 64)] * 2, 1)
    pd, md = inf.get_final_preds_device(torch.from_numpy(big).to(dev), c, s, True)
    pr, mr = restate.final_preds(big, c, s, True)
    assert np.array_equal(pd.cpu().numpy(), pr) and np.array_equal(md.cpu().numpy(), mr)


def test_argmax_bit_exact(golden, dev):
    import lib.core.inference as inf
    g = golden("argmax")
    hm = gi.argmax_heatmaps()
    preds, maxvals = inf.get_max_preds(hm)
    assert np.array_equal(preds, g["preds"]) and np.array_equal(maxvals, g["maxvals"])
    big = np.random.default_rng(9).standard_normal((32, 17, 64, 64)).astype(np.float32)
    big[:, :, 10, 10] = big.max() + 1          # ties across maps at a fixed location
    big[3, 2, 5, 5] = big[3, 2, 10, 10]        # earlier tie wins
    p, m, idx = inf.get_max_preds_device(torch.from_numpy(big).to(dev))
    rp, rm, ridx = restate.get_max_preds(big)
    assert np.array_equal(idx.cpu().numpy(), ridx.astype(np.int32))
    assert np.array_equal(p.cpu().numpy(), rp) and np.array_equal(m.cpu().numpy(), rm)
    e = inf.get_max_preds_device(torch.zeros((0, 17, 64, 64), device=dev))
    assert e[0].shape == (0, 17, 2)


# ------------------------------------------------------------------ geometry (fp64)
def test_triangulators_golden(golden, dev):
    import lib.utils.triangulation as tri
    g = golden("triangulation")
    u1, u2, P1, P2, X = gi.triangulation_case()
    for name in ("linear_eigen_triangulation", "linear_LS_triangulation", "iterative_LS_triangulation"):
        for i in range(len(u1)):
            x, st = getattr(tri, name)(u1[i], P1[i], u2[i], P2[i])
            assert np.max(np.abs(x - g[name + "_x"][i])) <= 1e-4, name     # mm
            assert np.array_equal(np.asarray(st).astype(np.int64), g[name + "_status"][i])
    u1e, u2e = gi.exact_projections(P1, P2, X)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for m in ("linear_eigen", "linear_LS", "iterative_LS"):
        Xg, _ = tri.triangulate_pairs(t(u1e), t(u2e), t(P1), t(P2), m)
        assert np.max(np.abs(Xg.cpu().numpy() - X)) <= 1e-6             # known answer


def test_eight_point_fallback_golden(golden, dev):
    """polynomial_triangulation's fallback (lib/utils/triangulation.py:213-217) on the device:
    "polynomial_8point" (the branch on its own: 8-point F from the matches, correction, DLT)
    against the same composition of OpenCV / reference calls, and the natural trigger --
    identical cameras, F = 0, all-NaN correction -- against the reference's own output."""
    import lib.utils.triangulation as tri
    g = golden("triangulation_8point")
    u1, u2, P1, P2, X = gi.triangulation_case()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Xg, st = tri.triangulate_pairs(t(u1), t(u2), t(P1), t(P2), "polynomial_8point")
    assert np.max(np.abs(Xg.cpu().numpy() - g["x_8pt"])) <= 1e-4          # mm
    Xs, st = tri.triangulate_pairs(t(u1), t(u2), t(P1), t(P1), "polynomial")
    assert np.max(np.abs(Xs.cpu().numpy() - g["x_same"])) <= 1e-4
    assert np.array_equal(st.cpu().numpy().astype(np.int64), g["st_same"])
    x, s1 = tri.polynomial_triangulation(u1[3], P1[3], u2[3], P1[3])     # reference-shaped API
    assert np.max(np.abs(x - g["x_same"][3])) <= 1e-4 and s1.all()


def test_polynomial_triangulation_golden(golden, dev):
    """method "polynomial" (F from the projection matrices, Hartley-Sturm correction with the
    Laguerre root finder, homogeneous DLT) against the unmodified reference
    polynomial_triangulation / cv2.correctMatches: <= 1e-4 mm, status equal; exact projections
    recover the 3-D points; corrected matches agree with the other triangulators' input when the
    observations are noise free."""
    import lib.utils.triangulation as tri
    g = golden("triangulation_poly")
    u1, u2, P1, P2, X = gi.triangulation_case()
    for i in range(len(u1)):
        x, st = tri.polynomial_triangulation(u1[i], P1[i], u2[i], P2[i])
        assert np.max(np.abs(x - g["x"][i])) <= 1e-4                        # mm
        assert np.array_equal(np.asarray(st).astype(np.int64), g["status"][i])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    u1e, u2e = gi.exact_projections(P1, P2, X)
    Xg, _ = tri.triangulate_pairs(t(u1e), t(u2e), t(P1), t(P2), "polynomial")
    assert np.max(np.abs(Xg.cpu().numpy() - X)) <= 1e-6
    assert np.max(np.abs(Xg.cpu().numpy() - g["exact"])) <= 1e-6
    # batched: 64 pairs against the numpy oracle
    u1b, u2b, P1b, P2b, _ = gi.triangulation_case(n_pairs=64, J=17, seed=98)
    Xb, _ = tri.triangulate_pairs(t(u1b), t(u2b), t(P1b), t(P2b), "polynomial")
    for i in range(0, 64, 9):
        xr, _ = restate.polynomial_triangulation(u1b[i], P1b[i], u2b[i], P2b[i])
        assert np.max(np.abs(Xb[i].cpu().numpy() - xr)) <= 1e-4


@pytest.mark.parametrize("V", [2, 3, 4])
def test_nview_dlt_vs_oracle(dev, V):
    """epb_triangulate_nview against the numpy-SVD oracle (<= 1e-4 mm, 3 px noise), exact recovery
    from noise-free views, and V = 2 equal to the pair kernel (method 0)."""
    import lib.utils.triangulation as tri
    rng = np.random.default_rng(40 + V)
    NT, J = 16, 17
    R, T, f, c, P = restate.synthetic_cameras(rng, NT, 4)
    X = rng.normal(0, 400, (NT, J, 3))
    ue = np.stack([[restate.project(P[t, v], X[t]) for v in range(V)] for t in range(NT)])
    un = ue + rng.normal(0, 3, ue.shape)
    t64 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Xg, st = tri.triangulate_views(t64(un), t64(P[:, :V]))
    for t in range(NT):
        xo, so = restate.linear_eigen_triangulation_nview(un[t], P[t, :V])
        assert np.max(np.abs(Xg[t].cpu().numpy() - xo)) <= 1e-4
        assert np.array_equal(st[t].cpu().numpy().astype(bool), so)
    Xe, _ = tri.triangulate_views(t64(ue), t64(P[:, :V]))
    assert np.max(np.abs(Xe.cpu().numpy() - X)) <= 1e-6
    if V == 2:
        Xp, _ = tri.triangulate_pairs(t64(un[:, 0]), t64(un[:, 1]), t64(P[:, 0]), t64(P[:, 1]), "linear_eigen")
        assert np.max(np.abs(Xp.cpu().numpy() - Xg.cpu().numpy())) <= 1e-9
    e, _ = tri.triangulate_views(t64(un[:0]), t64(P[:0, :V]))
    assert e.shape == (0, J, 3)


def test_triangulation_large_vs_oracle(dev):
    import lib.utils.triangulation as tri
    u1, u2, P1, P2, X = gi.triangulation_case(n_pairs=64, J=17, seed=99)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for m, fn in (("linear_eigen", restate.linear_eigen_triangulation),
                  ("iterative_LS", restate.iterative_LS_triangulation)):
        Xg, st = tri.triangulate_pairs(t(u1), t(u2), t(P1), t(P2), m)
        Xg = Xg.cpu().numpy()
        for i in range(0, 64, 7):
            xr, sr = fn(u1[i], P1[i], u2[i], P2[i])
            assert np.max(np.abs(Xg[i] - xr)) <= 1e-4
    e, _ = tri.triangulate_pairs(t(u1[:0]), t(u2[:0]), t(P1[:0]), t(P2[:0]))
    assert e.shape == (0, 17, 3)
    # a point behind camera 1 gets the reference's negative status
    Xb = X[:1].copy()
    x, st = tri.iterative_LS_triangulation(u1[0], P1[0], u2[0], P2[0])
    xr, sr = restate.iterative_LS_triangulation(u1[0], P1[0], u2[0], P2[0])
    assert np.array_equal(st, sr)


def test_patch_to_image_and_selfsup_golden(golden, dev):
    import lib.utils.img_utils as iu
    g = golden("patch_to_image")
    coords, boxes = gi.patch_case()
    out = iu.trans_coords_from_patch_to_org_3d_batch(coords, boxes[:, 0], boxes[:, 1], boxes[:, 2],
                                                     boxes[:, 3], 256, 256, 2000, boxes[:, 4], boxes[:, 5])
    assert np.max(np.abs(out - g["kps"])) <= 5e-3      # inputs pass through float32 patch units
    one = iu.trans_coords_from_patch_to_org_3d(coords[1], *boxes[1, :4], 256, 256, 2000, 2000,
                                               scale=boxes[1, 4], rot=boxes[1, 5])
    assert np.max(np.abs(one - g["kps"][1])) <= 5e-3
    gs = golden("selfsup")
    logits, meta = gi.selfsup_case()
    mt = {k: torch.from_numpy(v) for k, v in meta.items()}
    label, weight = iu.self_supervision(torch.from_numpy(logits).to(dev), mt)
    assert np.max(np.abs(label - gs["label"])) <= 2e-5
    assert np.array_equal(weight, gs["weight"])


# ------------------------------------------------------------------ conv / BN kernels vs torch fp32
def _rand(dev, *s):
    return torch.randn(*s, device=dev)


@pytest.mark.parametrize("cfg", [
    ("conv", 32, 64, 1, 1, 0, 14), ("conv", 64, 64, 3, 1, 1, 14), ("conv", 64, 128, 3, 2, 1, 14),
    ("conv", 64, 256, 1, 2, 0, 14), ("conv", 3, 64, 7, 2, 3, 30), ("deconv", 64, 32, 4, 2, 1, 7),
    ("conv", 32, 40, 3, 1, 1, 9),
])
@pytest.mark.parametrize("precision", [0, 1, 3])
def test_conv_family_vs_torch(dev, cfg, precision):
    """fprop / dgrad / wgrad of one layer (through net.Conv geometry + C ABI)
    against torch fp32 on identical tensors: <= 1e-3 rel per the north star
    (the fp32 and 3xTF32 paths sit near 1e-5)."""
    from epipolarpose_b200 import net, ops
    import torch.nn.functional as F
    kind, cin, cout, k, s, p, hw = cfg
    N = 3
    tol = 3e-3 if precision == 1 else 1e-3      # single-pass TF32 carries 2^-11 operand rounding
    conv = net.Conv("t", kind, cin, cout, k, s, p, 0)
    eng = net.Engine(None, precision=precision)
    eng.dev = dev
    w = _rand(dev, *((cout, cin, k, k) if kind == "conv" else (cin, cout, k, k))) * 0.1
    x = _rand(dev, N, cin, hw, hw)
    sc, sh = torch.rand(cin, device=dev) + 0.5, _rand(dev, cin) * 0.1
    xa = torch.relu(x * sc[None, :, None, None] + sh[None, :, None, None]).requires_grad_(True)
    wt = w.clone().requires_grad_(True)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = F.conv2d(xa, wt, None, s, p) if kind == "conv" else F.conv_transpose2d(xa, wt, None, s, p)
        gout = _rand(dev, *ref.shape)
        ref.backward(gout)
    xn = torch.zeros(N, hw, hw, conv.cin_p, device=dev)
    ops.nchw_to_nhwc(x.contiguous(), xn, N, cin, hw, hw, conv.cin_p)
    scp = torch.ones(conv.cin_p, device=dev); scp[:cin] = sc
    shp = torch.zeros(conv.cin_p, device=dev); shp[:cin] = sh
    wf, wd = conv.pack(ops, w)
    stats = torch.zeros(2 * conv.cout_p, device=dev, dtype=torch.float64)
    out, Ho, Wo = eng._conv_fwd(conv, xn, N, hw, hw, wf, affine=(scp, shp), stats=stats)
    o = out[..., :cout].permute(0, 3, 1, 2)
    assert relerr(o.cpu().numpy(), ref.detach().cpu().numpy()) <= tol
    st_ref = torch.cat([ref.detach().double().sum((0, 2, 3)), (ref.detach().double() ** 2).sum((0, 2, 3))])
    st = torch.cat([stats[:cout], stats[conv.cout_p:conv.cout_p + cout]])
    assert relerr(st.cpu().numpy(), st_ref.cpu().numpy()) <= tol
    gn = torch.zeros(N, Ho, Wo, conv.cout_p, device=dev)
    ops.nchw_to_nhwc(gout.contiguous(), gn, N, cout, Ho, Wo, conv.cout_p)
    din = eng._conv_dgrad(conv, gn, N, hw, hw, wd)
    assert relerr(din[..., :cin].permute(0, 3, 1, 2).cpu().numpy(), xa.grad.cpu().numpy()) <= tol
    gw = torch.zeros_like(w)
    eng._conv_wgrad(conv, xn, gn, N, hw, hw, gw, affine=(scp, shp))
    assert relerr(gw.cpu().numpy(), wt.grad.cpu().numpy()) <= tol


def test_bn_pool_kernels_vs_torch(dev):
    from epipolarpose_b200 import ops
    import torch.nn.functional as F
    N, H, W, C = 3, 18, 14, 64
    x = _rand(dev, N, H, W, C) * 2 + 0.3
    M = N * H * W
    stats = torch.zeros(2 * C, device=dev, dtype=torch.float64)
    ops.channel_stats(x, M, C, stats)
    xd = x.double().reshape(M, C)
    assert relerr(stats[:C].cpu().numpy(), xd.sum(0).cpu().numpy()) <= 1e-6
    assert relerr(stats[C:].cpu().numpy(), (xd * xd).sum(0).cpu().numpy()) <= 1e-6
    gamma, beta = torch.rand(C, device=dev) + 0.5, _rand(dev, C)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    sc, sh, mu, inv = (torch.empty(C, device=dev) for _ in range(4))
    ops.bn_finalize(stats, M, C, gamma, beta, 1e-5, 0.1, rm, rv, sc, sh, mu, inv)
    xc = x.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    gt, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yref = F.batch_norm(xc, rm2, rv2, gt, bt, True, 0.1, 1e-5)
    assert relerr(rm.cpu().numpy(), rm2.cpu().numpy()) <= 1e-5 and relerr(rv.cpu().numpy(), rv2.cpu().numpy()) <= 1e-5
    # stem: bn + relu + maxpool, forward and backward
    pref = F.max_pool2d(torch.relu(yref), 3, 2, 1)
    Ho, Wo = pref.shape[2], pref.shape[3]
    y = torch.empty(N, Ho, Wo, C, device=dev)
    arg = torch.empty(N, Ho, Wo, C, device=dev, dtype=torch.uint8)
    ops.bn_relu_maxpool(x, sc, sh, y, arg, N, H, W, C)
    assert relerr(y.permute(0, 3, 1, 2).cpu().numpy(), pref.detach().cpu().numpy()) <= 1e-5
    g = _rand(dev, *pref.shape)
    pref.backward(g)
    gp = torch.empty(N, H, W, C, device=dev)
    ops.maxpool_bwd(g.permute(0, 2, 3, 1).contiguous(), arg, gp, N, H, W, C)
    sums = torch.zeros(2 * C, device=dev, dtype=torch.float64)
    ops.bn_bwd_reduce(gp, x, None, sc, sh, mu, inv, 1, M, C, sums)
    dx, dg, db = torch.empty_like(x), torch.empty(C, device=dev), torch.empty(C, device=dev)
    ops.bn_bwd_apply(gp, x, None, sc, sh, mu, inv, gamma, 1, sums, M, C, dx, dg, db)
    assert relerr(dx.permute(0, 3, 1, 2).cpu().numpy(), xc.grad.cpu().numpy()) <= 1e-4
    assert relerr(dg.cpu().numpy(), gt.grad.cpu().numpy()) <= 1e-4
    assert relerr(db.cpu().numpy(), bt.grad.cpu().numpy()) <= 1e-4
    # residual add + relu
    r = _rand(dev, N, H, W, C)
    out = torch.empty_like(x)
    ops.bn_act(x, sc, sh, r, None, None, 1, out, M, C)
    ref = torch.relu(yref.detach().permute(0, 2, 3, 1) + r)
    assert relerr(out.cpu().numpy(), ref.cpu().numpy()) <= 1e-5


# ------------------------------------------------------------------ whole network
@pytest.mark.parametrize("tag", list(gi.NET_CASES))
@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "f16x3"])
def test_network_vs_reference_golden(golden, dev, tag, precision):
    """Module surface (get_pose_net / state_dict / train / eval) on the GPU
    against outputs of the UNMODIFIED reference module on the same weights."""
    import lib.models as models
    from oracle import refshim
    c = gi.NET_CASES[tag]
    g = golden("net_" + tag)
    cfg = refshim.make_cfg(num_layers=c["layers"], num_joints=c["J"], volume=c["volume"],
                           depth_res=c["D"], image_size=(c["HW"], c["HW"]))
    model = models.pose3d_resnet.get_pose_net(cfg, False, precision=precision)
    shapes = restate_net.param_shapes(num_layers=c["layers"], num_joints=c["J"], volume=c["volume"],
                                      depth_res=c["D"])
    model.load_state_dict(restate_net.init_state(shapes, c["seed"]))
    model = model.to(dev).train()
    x = torch.from_numpy(gi.images(c["N"], c["HW"], c["seed"])).to(dev)
    out = model(x)
    outs = out if isinstance(out, tuple) else (out,)
    for i, o in enumerate(outs):
        assert tuple(o.shape) == g["out%d" % i].shape
        assert relerr(o.detach().cpu().numpy(), g["out%d" % i]) <= 1e-3
    gs = [torch.from_numpy(gi.grad_like(o.shape, c["seed"] + 1 + i)).to(dev) for i, o in enumerate(outs)]
    sum((o * gg).sum() for o, gg in zip(outs, gs)).backward()
    named = dict(model.named_parameters())
    checked = 0
    for k in g:                    # every gradient the unmodified reference run stored
        if k.startswith("grad/"):
            e = relerr(named[k[5:]].grad.cpu().numpy(), g[k])
            # These toy batches (2-3 images, BatchNorm over as few as 8 values per channel) are
            # ReLU-flip chaotic below the head: ONE flipped unit moves a trunk gradient by 2-9e-3
            # (the float32 oracle restatement itself sits 3.7e-3 from the reference run on r18,
            # the float64 one 8e-6; tests/test_oracle_pinned.py).  Head tensors are held to 1e-3;
            # the 1e-3 bar for every tensor is enforced at the BASELINE sizes (tests/test_gpu_sizes.py).
            tol = 1e-3 if k.startswith("grad/final_layer") or k.startswith("grad/depth_fc") else 2e-2
            assert e <= tol, "%s: %.3e" % (k, e)
            checked += 1
    assert checked >= 1
    sd = model.state_dict()
    assert relerr(sd["bn1.running_mean"].cpu().numpy(), g["bn1.running_mean"]) <= 1e-4
    assert relerr(sd["bn1.running_var"].cpu().numpy(), g["bn1.running_var"]) <= 1e-4
    assert int(sd["bn1.num_batches_tracked"]) == 1
    model.eval()
    with torch.no_grad():
        e = model(x)
    e = e[0] if isinstance(e, tuple) else e
    assert relerr(e.cpu().numpy(), g["eval_out0"]) <= 1e-3


def _engine_relu_masks(plan, S):
    """ReLU masks of one engine forward, in the oracle's call order, as NCHW bool tensors."""
    nchw = lambda t: t.permute(0, 3, 1, 2).cpu()
    masks = []
    x, z0 = S["stem"][0], S["stem"][1]
    b0 = S["bn"]["bn1"]
    masks.append(nchw(z0 * b0.scale + b0.shift > 0))
    for blk, rec in zip(plan.blocks, S["blocks"]):
        for ci in range(len(blk["convs"]) - 1):
            st = S["bn"][blk["bns"][ci][0]]
            masks.append(nchw(rec["z"][ci] * st.scale + st.shift > 0))
        masks.append(nchw(rec["out"] > 0))
    for (conv, (bname, C)), (src, aff, z, h, w) in zip(plan.deconvs, S["deconv"]):
        st = S["bn"][bname]
        masks.append(nchw(z * st.scale + st.shift > 0))
    return masks


@pytest.mark.parametrize("layers,precision", [(18, 0), (18, 3), (50, 3)])
def test_network_gradients_vs_oracle_fp64(dev, layers, precision):
    """EVERY parameter gradient of a full forward + integral-L1 loss + backward
    against the float64 oracle: <= 1e-3 rel per tensor.  ReLU' is discontinuous at
    0 and a network has ~1e6 pre-activations, so some sit within fp32 rounding
    noise of 0 (tools/grad_diag.py: one such element moves a whole layer's
    gradient by 1e-2 between ANY two fp32 implementations, the CPU oracle
    included).  The float64 oracle is therefore evaluated with the activation
    pattern of the run under test (forced_masks); everything else is independent."""
    from epipolarpose_b200 import net, ops
    # R50 at 64x64 would leave 2x2x4 = 16 samples per channel for layer4's batch statistics
    # (conditioning ~1e3: fp32 itself sits at the 1e-3 bar there); 128x128 gives 64.
    J, N = 3, 4
    HW = 128 if layers == 50 else 64
    D = HW // 4
    plan = net.PoseNetPlan(layers, J, True, D, (HW, HW))
    shapes = restate_net.param_shapes(num_layers=layers, num_joints=J, volume=True, depth_res=D)
    sd = restate_net.init_state(shapes, 5)
    x = gi.images(N, HW, 5)
    gt, wt = gi.labels(N, J, 5)
    eng = net.Engine(plan, precision=precision)
    params = {k: v.clone().to(dev) for k, v in sd.items()}
    logits, _, S = eng.forward(torch.from_numpy(x).to(dev), params, training=True)
    masks = _engine_relu_masks(plan, S)
    # loss head through the public criterion on the engine's (channels_last) logits
    import lib.core.integral_loss as il
    lg = logits.permute(0, 3, 1, 2).detach().requires_grad_(True)
    loss = il.L1JointLocationLoss(J)(lg, torch.from_numpy(gt).to(dev), torch.from_numpy(wt).to(dev))
    loss.backward()
    grads = {k: torch.zeros_like(v) for k, v in params.items() if v.is_floating_point() and "running" not in k}
    eng.backward(S, lg.grad.permute(0, 2, 3, 1).contiguous(), None, params, grads)
    dt = torch.float64
    p = {k: (v.to(dt).clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
             else (v.to(dt) if v.is_floating_point() else v)) for k, v in sd.items()}
    o = restate_net.forward(p, torch.from_numpy(x).to(dt), num_layers=layers, training=True,
                            forced_masks=list(masks))
    sm = torch.softmax(o.reshape(N, J, -1), 2).reshape(N, J, D, D, D)
    ar = torch.arange(D, dtype=dt)
    c = torch.stack([(sm.sum((2, 3)) * ar).sum(2) / D - 0.5, (sm.sum((2, 4)) * ar).sum(2) / D - 0.5,
                     (sm.sum((3, 4)) * ar).sum(2) / D - 0.5], 2).reshape(N, J * 3)
    l64 = ((c - torch.from_numpy(gt).to(dt)).abs() * torch.from_numpy(wt).to(dt)).sum() / N
    l64.backward()
    assert abs(loss.item() - l64.item()) <= 1e-4 * abs(l64.item())
    assert relerr(logits.permute(0, 3, 1, 2).cpu().numpy(), o.detach().numpy()) <= 1e-3
    worst = max((relerr(grads[k].cpu().numpy(), p[k].grad.numpy()), k) for k in grads)
    assert worst[0] <= 1e-3, worst


def test_fused_adam_matches_torch(dev):
    import lib.utils.utils as U
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in ((7, 3), (64,), (5, 5, 3))]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    a, b = U.FusedAdam(ps, lr=1e-2), torch.optim.Adam(qs, lr=1e-2)
    for it in range(5):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        a.step(); b.step()
    for p, q in zip(ps, qs):
        assert relerr(p.detach().cpu().numpy(), q.detach().cpu().numpy()) <= 1e-5


def test_graphed_train_step_matches_eager(dev):
    """The CUDA-graph stepper (first call eager, second captures, then replays) against
    plain eager steps: same loss trajectory; BatchNorm counters advance under replay; a
    host-side LR change reaches the captured Adam kernel (lr = 0 freezes the weights)."""
    import lib.models as models
    import lib.core.integral_loss as il
    import lib.core.function as fn
    import lib.utils.utils as U
    from oracle import refshim
    from tests import golden_inputs as gi
    J, D, HW = 4, 16, 64
    logits_np, meta_np = gi.selfsup_case(n_tuples=2, J=J, D=D)
    B = logits_np.shape[0]
    meta = {k: torch.from_numpy(v) for k, v in meta_np.items()}
    cfg = refshim.make_cfg(num_layers=18, num_joints=J, volume=True, depth_res=D, image_size=(HW, HW))
    sd = restate_net.init_state(restate_net.param_shapes(18, J, True, D), 3)
    xs = [torch.from_numpy(gi.images(B, HW, 40 + i)).to(dev) for i in range(5)]
    out = {}
    for mode in ("eager", "graph"):
        model = models.pose3d_resnet.get_pose_net(cfg, False)
        model.load_state_dict(sd)
        model = model.to(dev).train()
        crit = il.SmoothL1JointLocationLoss(J)
        opt = U.FusedAdam(list(model.parameters()), lr=1e-4)
        stepper = fn.GraphedTrainStep(model, crit, opt, online=True)
        losses = []
        for i in range(4):
            if mode == "graph":
                losses.append(float(stepper(xs[i], meta=meta)))
            else:
                import lib.utils.img_utils as iu
                losses.append(float(stepper.eager_step(xs[i], None, None, iu.pack_meta(meta, B, dev))))
        if mode == "graph":
            assert stepper.graph is not None
        before = {k: v.detach().clone() for k, v in model.named_parameters()}
        for gr in opt.param_groups:
            gr["lr"] = 0.0                      # e.g. an lr_scheduler step between epochs
        if mode == "graph":
            stepper(xs[4], meta=meta)
        else:
            opt.sync_hyper()
            stepper.eager_step(xs[4], None, None, iu.pack_meta(meta, B, dev))
        torch.cuda.synchronize()
        for k, v in model.named_parameters():
            assert torch.equal(v.detach(), before[k]), (mode, k)       # lr = 0 reached the kernel
        assert int(model.state_dict()["bn1.num_batches_tracked"]) == 5
        out[mode] = (losses, {k: v.detach().cpu().numpy() for k, v in model.named_parameters()})
    for a, b in zip(out["graph"][0], out["eager"][0]):
        assert abs(a - b) <= 2e-2 * abs(b), (out["graph"][0], out["eager"][0])
    assert abs(out["graph"][0][0] - out["eager"][0][0]) <= 1e-5 * abs(out["eager"][0][0])
    for k, v in out["eager"][1].items():
        assert relerr(out["graph"][1][k], v) <= 5e-2, k


def test_reference_script_flow(dev, tmp_path):
    """The call sequence of the reference's scripts/train.py (:83-182) against the mirror:
    model factory, DataParallel wrap, criterion by name, get_optimizer + MultiStepLR, dataset
    by name, DataLoader, train / validate / eval loops, checkpoint save + reload."""
    import torch.utils.data
    import lib.core.integral_loss as loss            # noqa: F401  (eval by name below)
    import lib.dataset as dataset                    # noqa: F401
    import lib.models as models
    from lib.core.config import config, reset_config
    from lib.core.function import train_integral, validate_integral, eval_integral
    from lib.utils.utils import get_optimizer, save_checkpoint
    reset_config()
    config.MODEL.NUM_JOINTS = 4
    config.MODEL.DEPTH_RES = 16
    config.MODEL.IMAGE_SIZE = [64, 64]
    config.MODEL.EXTRA.NUM_LAYERS = 18
    config.MODEL.INIT_WEIGHTS = False
    config.LOSS.FN = "SmoothL1JointLocationLoss"
    config.DATASET.DATASET = "synthetic_h36m"
    config.DATASET.SYNTHETIC_LEN = 24
    config.TRAIN.BATCH_SIZE = 8
    config.PRINT_FREQ = 1
    model = models.pose3d_resnet.get_pose_net(config, is_train=True)
    model = torch.nn.DataParallel(model, device_ids=[0]).cuda()
    criterion = eval("loss." + config.LOSS.FN)(num_joints=config.MODEL.NUM_JOINTS, norm=config.LOSS.NORM).cuda()
    optimizer = get_optimizer(config, model)
    sched = torch.optim.lr_scheduler.MultiStepLR(optimizer, [1], 0.1)
    ds = eval("dataset." + config.DATASET.DATASET)
    train_ds = ds(cfg=config, root="", image_set="train", is_train=True)
    valid_ds = ds(cfg=config, root="", image_set="valid", is_train=False)
    mk = lambda d, sh: torch.utils.data.DataLoader(d, batch_size=config.TRAIN.BATCH_SIZE, shuffle=sh,
                                                   num_workers=0, pin_memory=True)
    train_loader, valid_loader = mk(train_ds, True), mk(valid_ds, False)
    before = {k: v.detach().clone() for k, v in model.module.state_dict().items()}
    for epoch in range(2):
        avg = train_integral(config, train_loader, model, criterion, optimizer, epoch)
        sched.step()
        assert np.isfinite(avg)
        preds = validate_integral(valid_loader, model)
        assert preds.shape == (len(valid_ds), config.MODEL.NUM_JOINTS, 4) and np.isfinite(preds).all()
        perf = eval_integral(epoch, preds, valid_loader, str(tmp_path), debug=False)
        assert np.isfinite(perf)
        save_checkpoint({"epoch": epoch + 1, "model": "pose3d_resnet", "state_dict": model.state_dict(),
                         "perf": perf, "optimizer": optimizer.state_dict()}, True, str(tmp_path))
    after = model.module.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before if before[k].is_floating_point())
    # DataParallel-prefixed checkpoint reloads through the reference's own prefix-stripping path
    best = torch.load(str(tmp_path / "model_best.pth.tar"), map_location="cpu")
    assert all(k.startswith("module.") for k in best)
    fresh = models.pose3d_resnet.get_pose_net(config, is_train=False)
    torch.save(best, str(tmp_path / "mpii_like.pth.tar"))
    fresh.load_pretrained_pose_model(str(tmp_path / "mpii_like.pth.tar"))
    for k, v in fresh.state_dict().items():
        assert torch.equal(v.cpu(), after[k].cpu()), k
    reset_config()


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_refiner_vs_reference_golden(golden, dev, precision):
    """refiner MLP on the device: forward / backward / running statistics against
    the unmodified refiner/model.py (<= 1e-3 rel per tensor), eval forward, dropout masks replayed
    through the oracle, one clip-grad-norm + fused Adam step as refiner/main.py:49-56."""
    from oracle import restate_refiner as rr
    from epipolarpose_b200.refiner import model as rmodel
    import lib.utils.utils as U
    g = golden("refiner")
    sd = rr.init_state(rr.param_shapes(128, 45, 45), 17)
    m = rmodel.LinearModelPG(linear_size=128, p_dropout=0.0, input_size=45, output_size=45, precision=precision)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    x = torch.from_numpy(gi.grad_like((24, 45), 18)).to(dev).requires_grad_(True)
    tgt = torch.from_numpy(gi.grad_like((24, 45), 19)).to(dev)
    opt = U.FusedAdam(list(m.parameters()), lr=1e-3)
    p1, p2 = m(x)
    loss = torch.nn.functional.mse_loss(p1, tgt) + torch.nn.functional.mse_loss(p2, tgt)
    loss.backward()
    assert relerr(p1.detach().cpu().numpy(), g["p1"]) <= 1e-3 and relerr(p2.detach().cpu().numpy(), g["p2"]) <= 1e-3
    assert relerr(x.grad.cpu().numpy(), g["dx"]) <= 1e-3
    named = dict(m.named_parameters())
    for k in [k[5:] for k in g if k.startswith("grad/")]:
        assert relerr(named[k].grad.cpu().numpy(), g["grad/" + k]) <= 1e-3, k
    assert relerr(m.state_dict()["batch_norm1.running_var"].cpu().numpy(), g["batch_norm1.running_var"]) <= 1e-4
    torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=1.)
    opt.step()
    m.eval()
    with torch.no_grad():
        e1, e2 = m(x.detach())
    assert np.isfinite(e1.cpu().numpy()).all() and np.isfinite(e2.cpu().numpy()).all()
    m2 = rmodel.LinearModelPG(linear_size=128, p_dropout=0.5, input_size=45, output_size=45, precision=precision)
    m2.load_state_dict(sd)
    m2 = m2.to(dev).train()
    torch.manual_seed(7)
    q1, q2 = m2(x.detach())
    torch.manual_seed(7)
    masks = [(torch.rand(24, 128, device=dev) >= 0.5).cpu() for _ in range(10)]
    o1, o2 = rr.forward(sd, x.detach().cpu(), training=True, masks=masks, p_dropout=0.5)
    assert relerr(q1.detach().cpu().numpy(), o1.numpy()) <= 1e-3 and relerr(q2.detach().cpu().numpy(), o2.numpy()) <= 1e-3


def test_refiner_train_loop_and_checkpoint_gpu(dev, tmp_path):
    """refiner/main.py train() / test() / save_ckpt on the device (reference refiner/main.py:31-84):
    an epoch against an independent loop (oracle network + torch optimiser + torch's
    clip_grad_norm_ on the CPU).  The parity run uses momentum SGD -- parameter differences stay
    proportional to gradient differences (Adam divides by sqrt(v): elements with noise-only
    gradients take +-lr steps of arbitrary sign, so parameters are not comparable after Adam
    steps; Adam itself is pinned against torch.optim.Adam in test_fused_optimizers*).  Then the
    reference's configuration (Adam), checkpoint interchange and samples/s."""
    import logging
    import time
    import types
    from oracle import restate_refiner as rr
    from epipolarpose_b200.refiner import main as rmain, model as rmodel, utils as rutils, data as rdata
    import lib.utils.utils as U
    sd = rr.init_state(rr.param_shapes(1024, 45, 45), 17)
    m = rmodel.LinearModelPG(linear_size=1024, p_dropout=0.0, input_size=45, output_size=45).to(dev)
    m.load_state_dict(sd)
    ds = rdata.SyntheticPoses(is_train=True, n=256, seed=3)
    dl = torch.utils.data.DataLoader(ds, batch_size=64, shuffle=False)
    args = types.SimpleNamespace(lr=0.05, lr_decay=2, lr_gamma=0.9)
    opt = U.FusedSGD(list(m.parameters()), lr=args.lr, momentum=0.9)
    crit = torch.nn.MSELoss(reduction='mean')
    step, lr_now = rmain.train(m, dl, opt, 0, args.lr, crit, args, logging.getLogger("t"))
    assert step == 4
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
         for k, v in sd.items()}
    plist = [v for k, v in p.items() if torch.is_tensor(v) and v.requires_grad]
    ro = torch.optim.SGD(plist, lr=args.lr, momentum=0.9)
    g = 0
    for inp, tar in dl:
        g += 1
        if g % args.lr_decay == 0 or g == 1:
            for pg in ro.param_groups:
                pg['lr'] = args.lr * args.lr_gamma ** (g / args.lr_decay)
        o1, o2 = rr.forward(p, inp, training=True)
        ro.zero_grad()
        (torch.nn.functional.mse_loss(o1, tar) + torch.nn.functional.mse_loss(o2, tar)).backward()
        torch.nn.utils.clip_grad_norm_(plist, max_norm=1.)
        ro.step()
    for k, q in m.named_parameters():
        a, b, b0 = q.detach().cpu().numpy(), p[k].detach().numpy(), sd[k].numpy()
        moved = max(float(np.max(np.abs(b - b0))), 1e-12)          # what the epoch changed
        # a ReLU unit whose pre-activation sits within rounding of zero resolves differently in the
        # two float32 evaluations about once per step (10 layers x 65536 pre-activations); ONE such
        # flip moves the 1/64-weighted gradient row of that unit by ~1.5 % of the tensor's maximum
        # (measured per step: 1.6e-5 without a flip, 1-2.5e-2 with one).  So: the bulk of every
        # tensor (median) to 5e-4 of the epoch's movement, the rows of flipped units to 1e-2 / 5e-2.
        d = np.abs(a - b)
        assert np.median(d) <= 5e-4 * moved + 1e-7, k
        assert np.percentile(d, 99) <= 2e-2 * moved + 1e-7, k      # a flip touches a whole 1024-entry row
        assert d.max() <= 0.5 * moved + 1e-6, k                     # ... and is amplified by later layers
    err = rmain.test(m, torch.utils.data.DataLoader(rdata.SyntheticPoses(False, n=128, seed=3), batch_size=64))
    assert np.isfinite(err)
    # the reference's configuration: Adam; checkpoint in the reference's layout
    opt = U.FusedAdam(list(m.parameters()), lr=1e-3)
    args = types.SimpleNamespace(lr=1e-3, lr_decay=100000, lr_gamma=0.96)
    step, lr_now = rmain.train(m, dl, opt, 0, args.lr, crit, args, logging.getLogger("t"))
    rutils.save_ckpt({'epoch': 1, 'lr': lr_now, 'step': step, 'err': err, 'state_dict': m.state_dict(),
                      'optimizer': opt.state_dict()}, ckpt_path=str(tmp_path), is_best=False)
    ck = torch.load(str(tmp_path / 'last.pth.tar'), weights_only=False)
    t_opt = torch.optim.Adam([torch.nn.Parameter(v.detach().cpu().clone()) for v in m.parameters()], lr=1e-3)
    t_opt.load_state_dict(ck['optimizer'])
    assert int(t_opt.state[t_opt.param_groups[0]['params'][0]]['step']) == 4
    # throughput of the loop body at the reference's batch size (64): steps/s -> samples/s
    big = torch.utils.data.DataLoader(rdata.SyntheticPoses(True, n=64 * 50, seed=5), batch_size=64)
    rmain.train(m, big, opt, step, lr_now, crit, args, logging.getLogger("t"))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rmain.train(m, big, opt, step, lr_now, crit, args, logging.getLogger("t"))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("refiner train loop: %.0f samples/s (batch 64, %.2f ms/step, eager)" % (64 * 50 / dt, dt / 50 * 1e3))
