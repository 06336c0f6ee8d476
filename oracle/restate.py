"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy / torch-CPU fp32) of the
EpipolarPose hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path
(epipolarpose_b200/*) never does and fails loudly without its CUDA library.

Every function cites the reference file:line it restates (paths relative to
the reference root).  Pinning: tests/golden/*.npz were produced by running the
UNMODIFIED reference (oracle/refshim.py) in the build container with
tests/golden/make_golden.py; tests/test_oracle_pinned.py checks each function
below against those vectors, so the restatement is pinned to the reference's
own outputs (the reference ships no tests / golden vectors of its own:
SURVEY.md section 4).  Third-party arithmetic (OpenCV 4.x triangulatePoints /
solve(DECOMP_SVD) / getAffineTransform; pinned opencv=4.1.0 in the reference's
environment.yml:96, 4.13.0 executed here) is restated with numpy.linalg in
float64 and pinned through the same golden vectors.
"""
import math

import numpy as np

# --------------------------------------------------------------------------
# a7: soft-argmax  (lib/core/integral_loss.py:49-86)
# --------------------------------------------------------------------------


def softmax_integral(preds, num_joints, hm_width, hm_height, hm_depth):
    """preds [N, J*D, H, W] float32 -> [N, J*3] float32 (x,y,z interleaved).

    integral_loss.py:71-86: softmax over D*H*W per (n,j) (:73-74), marginal
    expectations with channel order [j][z][y][x] (:52), index ranges 0..dim-1
    (:61-63), then coord/dim - 0.5 (:81-83), cat on dim 2 (:84-85).
    float32 arithmetic as in the reference (torch CPU); accumulation order
    differs from ATen, so agreement is to ~1e-6 abs, not bit-exact."""
    p = np.asarray(preds, dtype=np.float32)
    n = p.shape[0]
    v = p.reshape(n, num_joints, -1).astype(np.float64)
    v = v - v.max(axis=2, keepdims=True)
    e = np.exp(v)
    sm = (e / e.sum(axis=2, keepdims=True))
    sm = sm.reshape(n, num_joints, hm_depth, hm_height, hm_width)
    ax = sm.sum(axis=(2, 3)) @ np.arange(hm_width, dtype=np.float64)
    ay = sm.sum(axis=(2, 4)) @ np.arange(hm_height, dtype=np.float64)
    az = sm.sum(axis=(3, 4)) @ np.arange(hm_depth, dtype=np.float64)
    x = ax / float(hm_width) - 0.5
    y = ay / float(hm_height) - 0.5
    z = az / float(hm_depth) - 0.5
    out = np.stack([x, y, z], axis=2).reshape(n, num_joints * 3)
    return out.astype(np.float32)


def softmax_integral_grad(preds, grad_out, num_joints, hm_width, hm_height, hm_depth):
    """d(sum(out*grad_out))/d(preds): analytic backward of softmax_integral
    (what autograd produces through integral_loss.py:71-86).
    dL/dlogit_i = p_i * (g.c_i - sum_k p_k g.c_k) with c_i = (x/W, y/H, z/D)."""
    p = np.asarray(preds, dtype=np.float32)
    n = p.shape[0]
    v = p.reshape(n, num_joints, -1).astype(np.float64)
    v = v - v.max(axis=2, keepdims=True)
    e = np.exp(v)
    sm = (e / e.sum(axis=2, keepdims=True)).reshape(
        n, num_joints, hm_depth, hm_height, hm_width)
    g = np.asarray(grad_out, dtype=np.float64).reshape(n, num_joints, 3)
    ix = np.arange(hm_width, dtype=np.float64) / hm_width
    iy = np.arange(hm_height, dtype=np.float64) / hm_height
    iz = np.arange(hm_depth, dtype=np.float64) / hm_depth
    s = (g[:, :, 0, None, None, None] * ix[None, None, None, None, :]
         + g[:, :, 1, None, None, None] * iy[None, None, None, :, None]
         + g[:, :, 2, None, None, None] * iz[None, None, :, None, None])
    mean = (sm * s).sum(axis=(2, 3, 4), keepdims=True)
    return (sm * (s - mean)).reshape(p.shape).astype(np.float32)


# --------------------------------------------------------------------------
# a9: weighted losses  (lib/core/integral_loss.py:7-47)
# --------------------------------------------------------------------------

def heatmap_joint_loss(hm, target, hm_weight=None, x=None, t=None, w=None, kind="l1",
                       hm_scale=1.0, jt_scale=1.0, size_average=True):
    """Objective of the VOLUME=False head (north_star "MSE heatmap loss + L1 3D loss"; SURVEY
    8(d) C2(ii)): torch.nn.functional.mse_loss(w*hm, w*target) (mean over every element; the
    reference ships no heat-map criterion, only config.py:32-34) plus weighted_loss (:7-47) on
    the joint vector.  float64.  Returns (loss_hm, loss_jt, total, dtotal/dhm, dtotal/dx)."""
    h = np.asarray(hm, dtype=np.float64)
    g = np.asarray(target, dtype=np.float64)
    wr = np.ones(h.shape[:2]) if hm_weight is None else \
        np.asarray(hm_weight, dtype=np.float64).reshape(h.shape[:2])
    wr = wr.reshape(h.shape[:2] + (1,) * (h.ndim - 2))
    d = wr * (h - g)
    loss_hm = float((d * d).mean())
    dhm = hm_scale * 2.0 * wr * d / d.size
    loss_jt, dx = 0.0, None
    if x is not None:
        loss_jt, dxx = weighted_loss(kind, x, t, w, size_average, False)
        dx = jt_scale * dxx
    return loss_hm, float(loss_jt), hm_scale * loss_hm + jt_scale * float(loss_jt), dhm, dx


def weighted_loss(kind, inp, target, weights, size_average=True, norm=False):
    """kind in {'mse','l1','smoothl1'}; integral_loss.py:7-18 / 20-31 / 33-47.
    Divisor is len(input) = batch size (:16,29,45).  Returns (loss, dL/dinput)
    in float64 for checking (reference computes in float32)."""
    x = np.asarray(inp, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64)
    sx = st = 1.0
    if norm:  # :9-11 divide each by its global L1 norm
        sx = np.abs(x).sum()
        st = np.abs(t).sum()
    xn, tn = x / sx, t / st
    d = xn - tn
    if kind == "mse":
        out, dd = d * d, 2 * d
    elif kind == "l1":
        out, dd = np.abs(d), np.sign(d)
    elif kind == "smoothl1":
        a = np.abs(d)
        out = np.where(a < 1.0, 0.5 * d * d, a - 0.5)
        dd = np.where(a < 1.0, d, np.sign(d))
    else:
        raise ValueError(kind)
    div = float(len(x)) if size_average else 1.0
    loss = (out * w).sum() / div
    gxn = dd * w / div
    if norm:
        # d(x/||x||_1)/dx = I/s - x sign(x)^T / s^2
        gx = gxn / sx - np.sign(x) * (gxn * x).sum() / (sx * sx)
    else:
        gx = gxn
    return loss, gx


# --------------------------------------------------------------------------
# a8: hard argmax  (lib/core/inference.py:12-40)
# --------------------------------------------------------------------------

def get_max_preds(batch_heatmaps):
    """inference.py:12-40: flat argmax (first index on ties) -> (x=idx%W,
    y=floor(idx/W)) float32, masked to 0 where max<=0 (:35-39)."""
    hm = np.asarray(batch_heatmaps)
    n, j, h, w = hm.shape
    flat = hm.reshape(n, j, -1)
    idx = np.argmax(flat, axis=2)
    maxvals = np.max(flat, axis=2).reshape(n, j, 1)
    preds = np.zeros((n, j, 2), dtype=np.float32)
    preds[:, :, 0] = (idx % w).astype(np.float32)
    preds[:, :, 1] = np.floor(idx.astype(np.float32) / w)
    preds *= (maxvals > 0.0).astype(np.float32)
    return preds, maxvals, idx


def final_preds_refine(batch_heatmaps, coords):
    """inference.py:49-61 (+-0.25 px toward the higher neighbour)."""
    hm = np.asarray(batch_heatmaps)
    n, j, h, w = hm.shape
    c = np.array(coords, dtype=np.float32, copy=True)
    for a in range(n):
        for b in range(j):
            px = int(math.floor(c[a, b, 0] + 0.5))
            py = int(math.floor(c[a, b, 1] + 0.5))
            if 1 < px < w - 1 and 1 < py < h - 1:
                diff = np.array([hm[a, b, py, px + 1] - hm[a, b, py, px - 1],
                                 hm[a, b, py + 1, px] - hm[a, b, py - 1, px]])
                c[a, b] += np.sign(diff) * .25
    return c


def _lu6_affine(src, dst):
    """cv2.getAffineTransform(src, dst): the 6x6 system A.x = b solved as cv::solve(DECOMP_LU)
    does (partial pivoting by largest magnitude, in float64) -- the same elimination order as
    OpenCV so that the float64 coefficients agree to the last bit (lib/utils/transforms.py:75-77)."""
    A = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        A[2 * i, 0:3] = [float(src[i][0]), float(src[i][1]), 1.0]
        A[2 * i + 1, 3:6] = [float(src[i][0]), float(src[i][1]), 1.0]
        b[2 * i], b[2 * i + 1] = float(dst[i][0]), float(dst[i][1])
    for i in range(6):
        k = i
        for j in range(i + 1, 6):
            if abs(A[j, i]) > abs(A[k, i]):
                k = j
        if k != i:
            A[[i, k], i:] = A[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / A[i, i]
        for j in range(i + 1, 6):
            alpha = A[j, i] * d
            for kk in range(i + 1, 6):
                A[j, kk] += alpha * A[i, kk]
            b[j] += alpha * b[i]
    for i in range(5, -1, -1):
        s_ = b[i]
        for kk in range(i + 1, 6):
            s_ -= A[i, kk] * b[kk]
        b[i] = s_ / A[i, i]
    return b.reshape(2, 3)


def final_preds(batch_heatmaps, center, scale, post_process=True):
    """inference.py:43-68: get_max_preds, +-0.25 px refinement, transform_preds
    (transforms.py:39-44 with get_affine_transform(center, scale, 0, [W, H], inv=1), :47-79).
    Returns (preds [N,J,2] float32, maxvals [N,J,1])."""
    hm = np.asarray(batch_heatmaps)
    n, j, h, w = hm.shape
    coords, maxvals, _ = get_max_preds(hm)
    if post_process:
        coords = final_preds_refine(hm, coords)
    preds = coords.copy()
    for i in range(n):
        sc = np.asarray(scale[i], dtype=np.float64).reshape(-1)
        if sc.size == 1:
            sc = np.array([sc[0], sc[0]])
        scale_tmp = sc * 200.0
        src_w = scale_tmp[0]
        src_dir = [0 * 1.0 - (src_w * -0.5) * 0.0, 0 * 0.0 + (src_w * -0.5) * 1.0]    # get_dir, rot = 0
        dst_dir = np.array([0, w * -0.5], np.float32)
        src = np.zeros((3, 2), dtype=np.float32)
        dst = np.zeros((3, 2), dtype=np.float32)
        c = np.asarray(center[i], dtype=np.float64)
        src[0, :] = c
        src[1, :] = c + src_dir
        dst[0, :] = [w * 0.5, h * 0.5]
        dst[1, :] = np.array([w * 0.5, h * 0.5]) + dst_dir
        d = src[0] - src[1]
        src[2, :] = src[1] + np.array([-d[1], d[0]], dtype=np.float32)
        d = dst[0] - dst[1]
        dst[2, :] = dst[1] + np.array([-d[1], d[0]], dtype=np.float32)
        t = _lu6_affine(dst, src)                       # inv = 1: heat-map -> image
        for p_ in range(j):
            x, y = float(coords[i, p_, 0]), float(coords[i, p_, 1])
            preds[i, p_, 0] = t[0, 0] * x + t[0, 1] * y + t[0, 2] * 1.0
            preds[i, p_, 1] = t[1, 0] * x + t[1, 1] * y + t[1, 2] * 1.0
    return preds, maxvals


# --------------------------------------------------------------------------
# a10: decode  (lib/core/integral_loss.py:187-207)
# --------------------------------------------------------------------------

def joint_location_result(patch_width, patch_height, coords_norm):
    """coords_norm [N, J*3] float32 (output of softmax_integral) ->
    [N,J,4] float64: (x+.5)*pw, (y+.5)*ph, z*pw, score 1 (:196-205)."""
    c = np.asarray(coords_norm).astype(float)
    c = c.reshape(c.shape[0], c.shape[1] // 3, 3).copy()
    c[:, :, 0] = (c[:, :, 0] + 0.5) * patch_width
    c[:, :, 1] = (c[:, :, 1] + 0.5) * patch_height
    c[:, :, 2] = c[:, :, 2] * patch_width
    return np.concatenate([c, np.ones(c.shape[:2] + (1,))], axis=2)


# --------------------------------------------------------------------------
# a11: patch <-> image affine  (lib/utils/img_utils.py:63-111,141-155)
# --------------------------------------------------------------------------

def _affine_from_3pts(src, dst):
    """Restates cv2.getAffineTransform (OpenCV imgproc/imgwarp.cpp): the 2x3 float64 M with
    M*[sx,sy,1]^T = [dx,dy]^T for three float32 point pairs, solved like cv::solve(DECOMP_LU)
    (hal::LU64f: 6x6 system, partial pivoting, alpha = A[j][i] * (-1/A[i][i]), back
    substitution) -- bit-exact against the installed OpenCV on 2000 random triplets."""
    s = np.asarray(src, dtype=np.float32).astype(np.float64)
    d = np.asarray(dst, dtype=np.float32).astype(np.float64)
    A = [[0.0] * 6 for _ in range(6)]
    b = [0.0] * 6
    for i in range(3):
        A[2 * i][0] = A[2 * i + 1][3] = float(s[i][0])
        A[2 * i][1] = A[2 * i + 1][4] = float(s[i][1])
        A[2 * i][2] = A[2 * i + 1][5] = 1.0
        b[2 * i], b[2 * i + 1] = float(d[i][0]), float(d[i][1])
    eps = np.finfo(np.float64).eps * 100
    for i in range(6):
        k = i
        for j in range(i + 1, 6):
            if abs(A[j][i]) > abs(A[k][i]):
                k = j
        if abs(A[k][i]) < eps:
            return np.zeros((2, 3))
        if k != i:
            A[i], A[k] = A[k], A[i]
            b[i], b[k] = b[k], b[i]
        dd = -1.0 / A[i][i]
        for j in range(i + 1, 6):
            alpha = A[j][i] * dd
            for kk in range(i + 1, 6):
                A[j][kk] += alpha * A[i][kk]
            b[j] += alpha * b[i]
    for i in range(5, -1, -1):
        sacc = b[i]
        for kk in range(i + 1, 6):
            sacc -= A[i][kk] * b[kk]
        b[i] = sacc / A[i][i]
    return np.array(b, dtype=np.float64).reshape(2, 3)


def gen_trans_from_patch(c_x, c_y, src_width, src_height, dst_width, dst_height,
                         scale, rot, inv=False):
    """img_utils.py:72-105 incl. its float32 roundings: rotate_2d returns f32
    (:69), src/dst point arrays are f32 (:91-99)."""
    src_w = src_width * scale
    src_h = src_height * scale
    src_center = np.array([c_x, c_y], dtype=np.float64)
    rot_rad = np.pi * rot / 180

    def rotate_2d(pt, r):  # :63-69 (inputs are f32 arrays, math in f32*f64 -> f64, cast f32)
        x, y = pt[0], pt[1]
        sn, cs = np.sin(r), np.cos(r)
        return np.array([x * cs - y * sn, x * sn + y * cs], dtype=np.float32)

    src_downdir = rotate_2d(np.array([0, src_h * 0.5], dtype=np.float32), rot_rad)
    src_rightdir = rotate_2d(np.array([src_w * 0.5, 0], dtype=np.float32), rot_rad)
    dst_center = np.array([dst_width * 0.5, dst_height * 0.5], dtype=np.float32)
    dst_downdir = np.array([0, dst_height * 0.5], dtype=np.float32)
    dst_rightdir = np.array([dst_width * 0.5, 0], dtype=np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = src_center
    src[1, :] = src_center + src_downdir
    src[2, :] = src_center + src_rightdir
    dst = np.zeros((3, 2), dtype=np.float32)
    dst[0, :] = dst_center
    dst[1, :] = dst_center + dst_downdir
    dst[2, :] = dst_center + dst_rightdir
    if inv:
        return _affine_from_3pts(dst, src)
    return _affine_from_3pts(src, dst)


def trans_coords_from_patch_to_org_3d(coords_in_patch, c_x, c_y, bb_w, bb_h,
                                      patch_w, patch_h, rect_3d_w, rect_3d_h,
                                      scale=1.0, rot=0):
    """img_utils.py:141-155."""
    out = np.array(coords_in_patch, dtype=np.float64, copy=True)
    t = gen_trans_from_patch(c_x, c_y, bb_w, bb_h, patch_w, patch_h, scale, rot, inv=True)
    xy1 = np.concatenate([out[:, 0:2], np.ones((out.shape[0], 1))], axis=1)
    out[:, 0:2] = xy1 @ t.T
    out[:, 2] = np.asarray(coords_in_patch)[:, 2] / patch_w * rect_3d_w
    return out


# --------------------------------------------------------------------------
# a13: triangulators  (lib/utils/triangulation.py)
# --------------------------------------------------------------------------

def linear_eigen_triangulation(u1, P1, u2, P2, max_coordinate_value=1.e16):
    """triangulation.py:8-27 = cv2.triangulatePoints (OpenCV
    calib3d/triangulate.cpp icvTriangulatePoints): per point the 4x4 matrix
    A with rows  x*P[2]-P[0], y*P[2]-P[1]  for each view; homogeneous solution
    = right-singular vector of the smallest singular value; then /w (:24)."""
    u1 = np.asarray(u1, dtype=np.float64)
    u2 = np.asarray(u2, dtype=np.float64)
    P = [np.asarray(P1, dtype=np.float64)[0:3, 0:4], np.asarray(P2, dtype=np.float64)[0:3, 0:4]]
    n = len(u1)
    X = np.zeros((n, 3))
    for i in range(n):
        A = np.zeros((4, 4))
        for v, u in enumerate((u1[i], u2[i])):
            A[2 * v + 0] = u[0] * P[v][2] - P[v][0]
            A[2 * v + 1] = u[1] * P[v][2] - P[v][1]
        _, _, vt = np.linalg.svd(A)
        h = vt[-1]
        X[i] = h[0:3] / h[3]
    with np.errstate(invalid="ignore"):
        status = np.max(np.abs(X), axis=1) <= max_coordinate_value
    return X, status


def _build_Ab(u1, P1, u2, P2):
    """triangulation.py:139-150 (= :80-92): A rows C*P[:3,:3], b = -C*P[:3,3]
    with C = [[-1,0,u],[0,-1,v]]."""
    A = np.zeros((4, 3))
    b = np.zeros(4)
    for v, (u, P) in enumerate(((u1, P1), (u2, P2))):
        C = np.array([[-1.0, 0.0, u[0]], [0.0, -1.0, u[1]]])
        A[2 * v:2 * v + 2] = C @ P[0:3, 0:3]
        b[2 * v:2 * v + 2] = -(C @ P[0:3, 3])
    return A, b


def linear_LS_triangulation(u1, P1, u2, P2):
    """triangulation.py:34-97; cv2.solve(DECOMP_SVD) == least squares."""
    u1 = np.asarray(u1, dtype=np.float64)
    u2 = np.asarray(u2, dtype=np.float64)
    P1 = np.asarray(P1, dtype=np.float64)
    P2 = np.asarray(P2, dtype=np.float64)
    X = np.zeros((len(u1), 3))
    for i in range(len(u1)):
        A, b = _build_Ab(u1[i], P1, u2[i], P2)
        X[i] = np.linalg.lstsq(A, b, rcond=None)[0]
    return X, np.ones(len(u1), dtype=bool)


def iterative_LS_triangulation(u1, P1, u2, P2, tolerance=3.e-5):
    """triangulation.py:104-181, including the CUMULATIVE re-weighting of A,b
    (:165-169 multiply the already re-weighted rows again) and the status
    arithmetic (:175-178; `i < 10` is always true)."""
    u1 = np.asarray(u1, dtype=np.float64)
    u2 = np.asarray(u2, dtype=np.float64)
    P1 = np.asarray(P1, dtype=np.float64)
    P2 = np.asarray(P2, dtype=np.float64)
    n = len(u1)
    X = np.zeros((n, 3))
    status = np.zeros(n, dtype=int)
    for xi in range(n):
        A, b = _build_Ab(u1[xi], P1, u2[xi], P2)
        d1 = d2 = 1.0
        x = np.zeros(3)
        d1n = d2n = 1.0
        for _ in range(10):
            x = np.linalg.lstsq(A, b, rcond=None)[0]
            xh = np.array([x[0], x[1], x[2], 1.0])
            d1n = P1[2, :].dot(xh)
            d2n = P2[2, :].dot(xh)
            if abs(d1n - d1) <= tolerance and abs(d2n - d2) <= tolerance:
                break
            A[0:2] *= 1 / d1n
            A[2:4] *= 1 / d2n
            b[0:2] *= 1 / d1n
            b[2:4] *= 1 / d2n
            d1, d2 = d1n, d2n
        X[xi] = x
        st = int(d1n > 0 and d2n > 0)
        if d1n <= 0:
            st -= 1
        if d2n <= 0:
            st -= 2
        status[xi] = st
    return X, status


# --------------------------------------------------------------------------
# a6/a12/a14/a15: cameras, pairing, projection to labels, self_supervision
# --------------------------------------------------------------------------

def correct_matches(F, u1, u2):
    """cv2.correctMatches restated (OpenCV calib3d triangulate.cpp, Hartley-Sturm optimal
    correction, Hartley & Zisserman Alg. 12.1; the reference calls it at
    lib/utils/triangulation.py:212): per match, translate both points to the origin, rotate
    the epipoles onto the x axes, minimise s(t) = t^2/(1+f1^2 t^2) + (ct+d)^2/((at+b)^2 +
    f2^2 (ct+d)^2) over the real parts of the roots of the degree-6 stationarity polynomial and
    the asymptote t = inf, and map the closest points on the two epipolar lines back.
    u1, u2 [N,2] float64 -> corrected (u1, u2)."""
    F = np.asarray(F, dtype=np.float64)
    u1 = np.asarray(u1, dtype=np.float64)
    u2 = np.asarray(u2, dtype=np.float64)
    o1, o2 = np.empty_like(u1), np.empty_like(u2)
    for p in range(len(u1)):
        x1, y1, x2, y2 = u1[p, 0], u1[p, 1], u2[p, 0], u2[p, 1]
        T1i = np.array([[1, 0, x1], [0, 1, y1], [0, 0, 1.0]])
        T2i = np.array([[1, 0, x2], [0, 1, y2], [0, 0, 1.0]])
        TFT = T2i.T @ F @ T1i
        U, _, Vt = np.linalg.svd(TFT)
        e1 = Vt[2] / np.sqrt(Vt[2, 0] ** 2 + Vt[2, 1] ** 2)          # right epipole, F e1 = 0
        e2 = U[:, 2] / np.sqrt(U[0, 2] ** 2 + U[1, 2] ** 2)          # left epipole, e2^T F = 0
        R1 = np.array([[e1[0], e1[1], 0], [-e1[1], e1[0], 0], [0, 0, 1.0]])
        R2 = np.array([[e2[0], e2[1], 0], [-e2[1], e2[0], 0], [0, 0, 1.0]])
        RF = R2 @ TFT @ R1.T
        f1, f2 = e1[2], e2[2]
        a, b, c, d = RF[1, 1], RF[1, 2], RF[2, 1], RF[2, 2]
        # g(t) = t((at+b)^2 + f2^2(ct+d)^2)^2 - (ad-bc)(1+f1^2 t^2)^2 (at+b)(ct+d)
        q = np.polyadd(np.polymul([a, b], [a, b]), f2 * f2 * np.polymul([c, d], [c, d]))
        g = np.polysub(np.polymul([1.0, 0.0], np.polymul(q, q)),
                       (a * d - b * c) * np.polymul(np.polymul([f1 * f1, 0, 1.0], [f1 * f1, 0, 1.0]),
                                                    np.polymul([a, b], [c, d])))
        if not np.isfinite(g).all():               # F = 0 (identical cameras): no epipoles
            o1[p] = np.nan
            o2[p] = np.nan
            continue
        cand = [r.real for r in np.roots(g)]

        def cost(t):
            return t * t / (1 + f1 * f1 * t * t) + (c * t + d) ** 2 / ((a * t + b) ** 2 + f2 * f2 * (c * t + d) ** 2)
        s_min, t_min = np.inf, np.inf
        for t in cand:
            sv = cost(t)
            if sv < s_min:
                s_min, t_min = sv, t
        s_inf = 1.0 / (f1 * f1) + c * c / (a * a + f2 * f2 * c * c)
        if s_inf < s_min:
            # OpenCV evaluates the asymptote but keeps the finite minimiser's lines only when it
            # wins; at t = inf the lines are l1 = (f1, 0, -1), l2 = (-f2 c, a, c)
            l1 = np.array([f1, 0.0, -1.0])
            l2 = np.array([-f2 * c, a, c])
        else:
            l1 = np.array([t_min * f1, 1.0, -t_min])
            l2 = np.array([-f2 * (c * t_min + d), a * t_min + b, c * t_min + d])
        xh1 = np.array([-l1[0] * l1[2], -l1[1] * l1[2], l1[0] ** 2 + l1[1] ** 2])
        xh2 = np.array([-l2[0] * l2[2], -l2[1] * l2[2], l2[0] ** 2 + l2[1] ** 2])
        n1 = T1i @ R1.T @ xh1
        n2 = T2i @ R2.T @ xh2
        o1[p] = n1[:2] / n1[2]
        o2[p] = n2[:2] / n2[2]
    return o1, o2


def fundamental_from_projections(P1, P2):
    """lib/utils/triangulation.py:198-204: canonical P = P2_full * inv(P1_full), F = [t]x R."""
    P1f, P2f = np.eye(4), np.eye(4)
    P1f[0:3, :] = np.asarray(P1, dtype=np.float64)[0:3, :]
    P2f[0:3, :] = np.asarray(P2, dtype=np.float64)[0:3, :]
    Pi = np.linalg.inv(P1f)
    Pc = P2f.dot(Pi)
    t = Pc[0:3, 3].copy()
    # a translation that is pure cancellation noise (identical camera centres) is exactly zero:
    # the reference's F is then the zero matrix and its correction all-NaN (:213-217)
    mag = np.abs(P2f[0:3, :]).dot(np.abs(Pi[:, 3]))
    t[np.abs(t) <= 64 * np.finfo(np.float64).eps * mag] = 0.0
    return np.cross(t, Pc[0:3, 0:3], axisb=0).T


def fundamental_8point(u1, u2):
    """cv2.findFundamentalMat(u1, u2, cv2.FM_8POINT)[0] restated (OpenCV calib3d fundam.cpp
    run8Point; called by the reference at lib/utils/triangulation.py:216): inputs rounded to
    float32, isotropic normalisation, eigenvector of the smallest eigenvalue of the 9x9 normal
    matrix, rank-2 projection, de-normalisation, F[2,2] = 1.  None for degenerate point sets."""
    a = np.asarray(u1, dtype=np.float64).astype(np.float32).astype(np.float64)
    b = np.asarray(u2, dtype=np.float64).astype(np.float32).astype(np.float64)
    c1, c2 = a.mean(0), b.mean(0)
    s1 = np.sqrt(((a - c1) ** 2).sum(1)).mean()
    s2 = np.sqrt(((b - c2) ** 2).sum(1)).mean()
    if s1 < np.finfo(np.float32).eps or s2 < np.finfo(np.float32).eps:
        return None
    s1, s2 = np.sqrt(2.0) / s1, np.sqrt(2.0) / s2
    p, q = (a - c1) * s1, (b - c2) * s2
    r = np.stack([q[:, 0] * p[:, 0], q[:, 0] * p[:, 1], q[:, 0], q[:, 1] * p[:, 0], q[:, 1] * p[:, 1],
                  q[:, 1], p[:, 0], p[:, 1], np.ones(len(p))], axis=1)
    w, v = np.linalg.eigh(r.T @ r)
    if (np.abs(w) >= np.finfo(np.float64).eps).sum() < 8:
        return None
    F0 = v[:, 0].reshape(3, 3)
    U, sv, Vt = np.linalg.svd(F0)
    sv[2] = 0.0
    F0 = U @ np.diag(sv) @ Vt
    T1 = np.array([[s1, 0, -s1 * c1[0]], [0, s1, -s1 * c1[1]], [0, 0, 1]])
    T2 = np.array([[s2, 0, -s2 * c2[0]], [0, s2, -s2 * c2[1]], [0, 0, 1]])
    F = T2.T @ F0 @ T1
    if abs(F[2, 2]) > np.finfo(np.float32).eps:
        F = F / F[2, 2]
    return F


def polynomial_triangulation_8point(u1, P1, u2, P2):
    """The fallback branch of triangulation.py:215-217 on its own: F from the matches."""
    F = fundamental_8point(u1, u2)
    n1, n2 = correct_matches(F, u1, u2)
    return linear_eigen_triangulation(n1, P1, n2, P2)


def polynomial_triangulation(u1, P1, u2, P2):
    """lib/utils/triangulation.py:184-220, including the fallback of :213-217: when the optimal
    correction is NaN for every match (F = 0: identical / degenerate cameras), F is re-estimated
    from the matches with the 8-point algorithm and the correction repeated."""
    u1 = np.asarray(u1, dtype=np.float64)[:, :2]
    u2 = np.asarray(u2, dtype=np.float64)[:, :2]
    F = fundamental_from_projections(P1, P2)
    with np.errstate(all="ignore"):
        n1, n2 = correct_matches(F, u1, u2)
    if np.isnan(n1).all() or np.isnan(n2).all():
        F8 = fundamental_8point(u1, u2)
        if F8 is not None:
            n1, n2 = correct_matches(F8, u1, u2)
    return linear_eigen_triangulation(n1, P1, n2, P2)


def linear_eigen_triangulation_nview(us, Ps):
    """V-view homogeneous DLT: us [V,J,2], Ps [V,3,4] -> (x [J,3], status);
    rows u*P[2]-P[0], v*P[2]-P[1] per view as cv2.triangulatePoints stacks them for two views
    (triangulation.py:22), solved with numpy's SVD."""
    us = np.asarray(us, dtype=np.float64)
    Ps = np.asarray(Ps, dtype=np.float64)
    V, J = us.shape[0], us.shape[1]
    x = np.zeros((J, 3))
    for j in range(J):
        A = np.zeros((2 * V, 4))
        for v in range(V):
            A[2 * v] = us[v, j, 0] * Ps[v, 2] - Ps[v, 0]
            A[2 * v + 1] = us[v, j, 1] * Ps[v, 2] - Ps[v, 1]
        h = np.linalg.svd(A)[2][-1]
        x[j] = h[:3] / h[3]
    return x, np.max(np.abs(x), axis=1) <= 1e16


def projection_matrix(R, T, f, c):
    """lib/utils/cameras.py:120-131,149-150: K.[R | R.(-T)] float64 3x4."""
    R = np.asarray(R, dtype=np.float64)
    T = np.asarray(T, dtype=np.float64).reshape(3, 1)
    K = np.array([[f[0], 0., c[0]], [0., f[1], c[1]], [0., 0., 1.]], dtype=np.float64)
    return K @ np.concatenate([R, R @ (-T)], axis=1)


def triangulate_batch(kps, Pmats, method="iterative"):
    """lib/utils/img_utils.py:193-209: sample i pairs with i + B/2; result
    duplicated for both halves (:207-208)."""
    fn = iterative_LS_triangulation if method == "iterative" else linear_eigen_triangulation
    half = kps.shape[0] // 2
    out = []
    for i in range(half):
        x, _ = fn(kps[i, :, 0:2], Pmats[i], kps[half + i, :, 0:2], Pmats[half + i])
        out.append(x)
    out = np.asarray(out)
    return np.vstack([out, out])


def labels_from_global_coords(X, meta):
    """lib/utils/img_utils.py:212-243 + lib/utils/prep_h36m.py:170-204 +
    lib/core/integral_loss.py:170-177.  meta: dict of arrays
    scale, rot, center_x, center_y, width, height, T[B,3(,1)], R[B,3,3], f[B,2], c[B,2]."""
    B, J = X.shape[0], X.shape[1]
    label = np.zeros((B, J * 3), dtype=np.float32)
    weight = np.ones((B, J * 3), dtype=np.float32)
    for i in range(B):
        R = np.asarray(meta["R"][i], dtype=np.float64)
        T = np.asarray(meta["T"][i], dtype=np.float64).reshape(3)
        f = np.asarray(meta["f"][i], dtype=np.float64).reshape(2)
        c = np.asarray(meta["c"][i], dtype=np.float64).reshape(2)
        cam = (X[i] - T) @ R.T                       # prep_h36m.py:186
        jt = np.zeros((J, 3))
        jt[:, 0] = cam[:, 0] / cam[:, 2] * f[0] + c[0]   # CamProj :170-175
        jt[:, 1] = cam[:, 1] / cam[:, 2] * f[1] + c[1]
        jt[:, 2] = cam[:, 2] - cam[0, 2]             # :199 (root joint 0)
        scale = float(meta["scale"][i])
        t = gen_trans_from_patch(float(meta["center_x"][i]), float(meta["center_y"][i]),
                                 float(meta["width"][i]), float(meta["height"][i]),
                                 256, 256, scale, float(meta["rot"][i]), inv=False)
        xy1 = np.concatenate([jt[:, 0:2], np.ones((J, 1))], axis=1)
        jt[:, 0:2] = xy1 @ t.T                       # img_utils.py:234-235
        jt[:, 2] = jt[:, 2] / (2000. * scale) * 256.  # :236
        jt[:, 0] = jt[:, 0] / 256. - 0.5             # integral_loss.py:171-173
        jt[:, 1] = jt[:, 1] / 256. - 0.5
        jt[:, 2] = jt[:, 2] / 256.
        label[i] = jt.reshape(-1).astype(np.float32)
    return label, weight


def self_supervision(coords_norm, meta, method="iterative"):
    """lib/utils/img_utils.py:166-190 downstream of the soft-argmax:
    coords_norm [B, J*3] f32 -> (label, weight) [B, J*3] f32."""
    res = joint_location_result(256, 256, coords_norm)
    B = res.shape[0]
    img = np.stack([
        trans_coords_from_patch_to_org_3d(
            res[i], float(meta["center_x"][i]), float(meta["center_y"][i]),
            float(meta["width"][i]), float(meta["height"][i]), 256, 256, 2000, 2000,
            scale=float(meta["scale"][i]), rot=float(meta["rot"][i]))
        for i in range(B)])
    X = triangulate_batch(img, np.asarray(meta["projection_matrix"]), method)
    return labels_from_global_coords(X, meta) + (X, img)


# --------------------------------------------------------------------------
# synthetic generators -- shared by tests and bench
# --------------------------------------------------------------------------

def synthetic_cameras(rng, n_tuples, n_views=4):
    """4 cameras per tuple on a ring r=4.5m+-0.5 at azimuths {45,135,225,315}+-10deg,
    height 1.5m+-0.2, looking at the origin; f=(1145,1144), c=(512,515).
    Returns R[n,v,3,3], T[n,v,3] (camera centre, world mm), f[n,v,2], c[n,v,2],
    P[n,v,3,4] in the reference convention X_cam = R.(X - T) (cameras.py:149-150)."""
    R = np.zeros((n_tuples, n_views, 3, 3))
    T = np.zeros((n_tuples, n_views, 3))
    f = np.tile(np.array([1145.0, 1144.0]), (n_tuples, n_views, 1))
    c = np.tile(np.array([512.0, 515.0]), (n_tuples, n_views, 1))
    P = np.zeros((n_tuples, n_views, 3, 4))
    for t in range(n_tuples):
        for v in range(n_views):
            az = np.deg2rad(45.0 + 90.0 * v + rng.uniform(-10, 10))
            r = 4500.0 + rng.uniform(-500, 500)
            h = 1500.0 + rng.uniform(-200, 200)
            C = np.array([r * np.cos(az), r * np.sin(az), h])
            zc = -C / np.linalg.norm(C)                 # optical axis -> origin
            up = np.array([0.0, 0.0, 1.0])
            xc = np.cross(zc, up)
            xc /= np.linalg.norm(xc)
            yc = np.cross(zc, xc)
            R[t, v] = np.stack([xc, yc, zc], axis=0)
            T[t, v] = C
            P[t, v] = projection_matrix(R[t, v], C, f[t, v], c[t, v])
    return R, T, f, c, P


def project(P, X):
    """X[...,3] world -> pixel (u,v) with P 3x4."""
    Xh = np.concatenate([X, np.ones(X.shape[:-1] + (1,))], axis=-1)
    uvw = Xh @ P.T
    return uvw[..., 0:2] / uvw[..., 2:3]


# ---------------------------------------------------------------- H36M evaluation protocol
H36M_NAMES = ['Hip', 'RHip', 'RKnee', 'RFoot', 'LHip', 'LKnee', 'LFoot', 'Spine', 'Thorax',
              'Neck/Nose', 'Head', 'LShoulder', 'LElbow', 'LWrist', 'RShoulder', 'RElbow', 'RWrist']
MPII_NAMES = ['RFoot', 'RKnee', 'RHip', 'LHip', 'LKnee', 'LFoot', 'Hip', 'Thorax', 'Neck/Nose',
              'Head', 'RWrist', 'RElbow', 'RShoulder', 'LShoulder', 'LElbow', 'LWrist']
# lib/dataset/h36m.py:17 (names: lib/dataset/JointIntegralDataset.py)
H36M_TO_MPII_PERM = np.array([H36M_NAMES.index(h) for h in MPII_NAMES if h != '' and h in H36M_NAMES])
J14_MPII = [0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 12, 13, 14, 15]          # h36m.py:186
J14_H36M = [0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 15]


def cam_back_proj(cam_x, cam_y, depth, fx, fy, u, v):
    """lib/utils/prep_h36m.py:85-89."""
    return (cam_x - u) / fx * depth, (cam_y - v) / fy * depth, depth


def compute_similarity_transform(X, Y, compute_optimal_scale=False):
    """lib/utils/prep_h36m.py:108-168 (MATLAB procrustes): X targets [N,M], Y inputs.
    Returns (d, Z, T, b, c)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    muX, muY = X.mean(0), Y.mean(0)
    X0, Y0 = X - muX, Y - muY
    ssX, ssY = (X0 ** 2.).sum(), (Y0 ** 2.).sum()
    normX, normY = np.sqrt(ssX), np.sqrt(ssY)
    X0, Y0 = X0 / normX, Y0 / normY
    A = np.dot(X0.T, Y0)
    U, s, Vt = np.linalg.svd(A, full_matrices=False)
    V = Vt.T
    T = np.dot(V, U.T)
    detT = np.linalg.det(T)
    V[:, -1] *= np.sign(detT)
    s[-1] *= np.sign(detT)
    T = np.dot(V, U.T)
    traceTA = s.sum()
    if compute_optimal_scale:
        b = traceTA * normX / normY
        d = 1 - traceTA ** 2
        Z = normX * traceTA * np.dot(Y0, T) + muX
    else:
        b = 1
        d = 1 + ssY / ssX - 2 * traceTA * normY / normX
        Z = normY * np.dot(Y0, T) + muX
    c = muX - b * np.dot(muY, T)
    return d, Z, T, b, c


def h36m_evaluate(preds, gt_joints, pelvis_z, fl, c_p, mpii_order=False, pck_thr=150.0):
    """lib/dataset/h36m.py:168-378 on arrays instead of db records.
    preds [S,J,>=3] image-space predictions (x, y px; root-relative depth mm), gt_joints [S,J,3]
    (`joints_3d`, H36M order), pelvis_z [S] (gt['pelvis'][2]), fl / c_p [S,2].
    Returns dict(metrics [S,9], per_joint [S,J], pck [S,J], poses [S,J,9], name_value, mean)."""
    preds = np.asarray(preds, dtype=np.float64)[:, :, 0:3]
    S, J = preds.shape[0], preds.shape[1]
    root = 6 if mpii_order else 0
    j14 = J14_MPII if mpii_order else J14_H36M
    metrics = np.zeros((S, 9))
    per_joint = np.zeros((S, J))
    pck = np.zeros((S, J), dtype=np.int32)
    poses = np.zeros((S, J, 9))
    for n in range(S):
        gt2 = np.asarray(gt_joints[n], dtype=np.float64).copy()
        pre2 = preds[n].copy()
        if mpii_order:
            gt2 = gt2[H36M_TO_MPII_PERM, :]
        pre2[:, 2] = pre2[:, 2] + pelvis_z[n]
        gt2[:, 2] = gt2[:, 2] + pelvis_z[n]
        pre3, gt3 = np.zeros((J, 3)), np.zeros((J, 3))
        for j in range(J):
            pre3[j] = cam_back_proj(pre2[j, 0], pre2[j, 1], pre2[j, 2], fl[n][0], fl[n][1], c_p[n][0], c_p[n][1])
            gt3[j] = cam_back_proj(gt2[j, 0], gt2[j, 1], gt2[j, 2], fl[n][0], fl[n][1], c_p[n][0], c_p[n][1])
        _, Z, T, b, c = compute_similarity_transform(gt3, pre3, compute_optimal_scale=True)
        align = (b * pre3.dot(T)) + c
        norm = b * pre3
        pre3 = pre3 - pre3[root]
        gt3 = gt3 - gt3[root]
        align = align - align[root]
        norm = norm - norm[root]
        diff, diff_a, diff_n = gt3 - pre3, gt3 - align, gt3 - norm
        e = np.linalg.norm(diff, axis=1)
        ea = np.linalg.norm(diff_a, axis=1)
        en = np.linalg.norm(diff_n, axis=1)
        metrics[n] = [e.mean(), ea.mean(), en.mean(), e[j14].mean(), ea[j14].mean(), en[j14].mean(),
                      np.abs(diff[:, 0]).mean(), np.abs(diff[:, 1]).mean(), np.abs(diff[:, 2]).mean()]
        per_joint[n] = e
        pck[n] = (e < pck_thr).astype(np.int32)
        poses[n] = np.concatenate([pre3, align, gt3], axis=1)
    names = ['hm36_17j      :', 'hm36_17j_align:', 'hm36_17j_norm:', 'hm36_17j_14   :', 'hm36_17j_14_al:',
             'hm36_17j_14_nm:', 'hm36_17j_x    :', 'hm36_17j_y    :', 'hm36_17j_z    :']
    means = metrics.mean(axis=0) if S else np.zeros(9)
    return dict(metrics=metrics, per_joint=per_joint, pck=pck, poses=poses,
                name_value=list(zip(names, means.tolist())), mean=float(means[0]))


# ---------------------------------------------------------------- input pipeline (8(f) row 1)
def warp_affine_linear_u8(img, M, dst_w, dst_h):
    """cv2.warpAffine(img u8 [H,W,C], M 2x3 f64, (dst_w, dst_h), flags=INTER_LINEAR), constant
    border 0, restated from OpenCV imgproc/imgwarp.cpp (the reference calls it at
    lib/utils/img_utils.py:125): M is inverted in float64; source coordinates are fixed point
    with AB_BITS = 10 and rounded per term (cvRound of M00*x*1024 per column, of
    (M01*y + M02)*1024 per row) plus round_delta = 16, then reduced to INTER_BITS = 5 fractional
    bits; the four neighbours (0 outside the image) are blended with integer weights
    (32-ax)(32-ay)*32 ... of sum 2^15 and rounded (+2^14) >> 15.  Bit-exact against the
    installed OpenCV 4.13 on the golden cases."""
    img = np.asarray(img)
    M = np.asarray(M, dtype=np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    iM = np.zeros((2, 3))
    iM[0, 0], iM[0, 1] = M[1, 1] * D, M[0, 1] * (-D)
    iM[1, 0], iM[1, 1] = M[1, 0] * (-D), M[0, 0] * D
    iM[0, 2] = -iM[0, 0] * M[0, 2] - iM[0, 1] * M[1, 2]
    iM[1, 2] = -iM[1, 0] * M[0, 2] - iM[1, 1] * M[1, 2]
    H, W = img.shape[:2]
    xs = np.arange(dst_w)
    adelta = np.rint(iM[0, 0] * xs * 1024).astype(np.int64)
    bdelta = np.rint(iM[1, 0] * xs * 1024).astype(np.int64)
    out = np.zeros((dst_h, dst_w) + img.shape[2:], dtype=np.uint8)

    def px(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
        v[~ok] = 0
        return v
    for y in range(dst_h):
        X0 = int(np.rint((iM[0, 1] * y + iM[0, 2]) * 1024)) + 16
        Y0 = int(np.rint((iM[1, 1] * y + iM[1, 2]) * 1024)) + 16
        X, Y = (X0 + adelta) >> 5, (Y0 + bdelta) >> 5
        sx, sy, ax, ay = X >> 5, Y >> 5, X & 31, Y & 31
        w = [((32 - ax) * (32 - ay) * 32), (ax * (32 - ay) * 32), ((32 - ax) * ay * 32), (ax * ay * 32)]
        sh = (-1,) + (1,) * (img.ndim - 2)
        acc = px(sy, sx) * w[0].reshape(sh) + px(sy, sx + 1) * w[1].reshape(sh) + \
            px(sy + 1, sx) * w[2].reshape(sh) + px(sy + 1, sx + 1) * w[3].reshape(sh)
        out[y] = np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)
    return out


def paste_over(im_src, im_dst, center):
    """lib/utils/augmentation.py:81-114: alpha-blend the RGBA im_src onto the uint8 im_dst in
    place, centred at np.round(center); float32 arithmetic, truncated back to uint8."""
    wh_src = np.asarray([im_src.shape[1], im_src.shape[0]])
    wh_dst = np.asarray([im_dst.shape[1], im_dst.shape[0]])
    center = np.round(center).astype(np.int32)
    raw_start = center - wh_src // 2
    raw_end = raw_start + wh_src
    start = np.clip(raw_start, 0, wh_dst)
    end = np.clip(raw_end, 0, wh_dst)
    region_dst = im_dst[start[1]:end[1], start[0]:end[0]]
    s0 = start - raw_start
    s1 = wh_src + (end - raw_end)
    region_src = im_src[s0[1]:s1[1], s0[0]:s1[0]]
    alpha = region_src[..., 3:].astype(np.float32) / np.float32(255)
    blend = alpha * region_src[..., 0:3].astype(np.float32) + (np.float32(1) - alpha) * region_dst.astype(np.float32)
    im_dst[start[1]:end[1], start[0]:end[0]] = blend.astype(np.uint8)


def patch_sample(cvimg, center_x, center_y, width, height, joints, joints_vis, patch_width,
                 patch_height, rect_3d_width, mean, std, scale=1.0, rot=0.0, do_flip=False,
                 color_scale=(1.0, 1.0, 1.0), flip_pairs=(), depth_in_image=False, occluders=None):
    """lib/utils/img_utils.py:246-298 (get_single_patch_sample) after the image is decoded and
    the augmentation parameters are drawn (occluders: the already resized RGBA images and their
    centres, pasted by paste_over in order): crop by warpAffine
    (:114-127), BGR->RGB (:268), per-channel colour scale + clip + (x-mean)/std (:277-281,
    float32 * python float stays float32; minus / divided by np.float64 scalars is float64,
    stored back as float32 -- numpy >= 2 promotion, the version the golden run used), joints
    through the same affine + depth scaling (:283-293) and generate_joint_location_label.
    Returns (img_patch f32 [3,ph,pw], label f64 [J*3], label_weight [J*3], trans 2x3)."""
    img = np.asarray(cvimg)
    img_h, img_w = img.shape[:2]
    c_x = center_x
    if do_flip:
        img = img[:, ::-1, :]
        c_x = img_w - c_x - 1
    trans = gen_trans_from_patch(c_x, center_y, width, height, patch_width, patch_height, scale, rot, inv=False)
    patch = warp_affine_linear_u8(img, trans, int(patch_width), int(patch_height))
    image = patch[:, :, ::-1]
    if occluders:                      # :269-270, augmentation.py:61-78 after the draws / resizes
        image = image.copy()
        for rgba, center in occluders:
            paste_over(rgba, image, np.asarray(center, dtype=np.float64))
    t = np.transpose(image, (2, 0, 1)).astype(np.float32)
    for c in range(t.shape[0]):
        t[c] = np.clip(t[c] * np.float32(color_scale[c]), 0, 255)
        if mean is not None and std is not None:
            t[c] = ((t[c].astype(np.float64) - np.float64(mean[c])) / np.float64(std[c])).astype(np.float32)
    joints = np.array(joints, dtype=np.float64, copy=True)
    joints_vis = np.array(joints_vis, dtype=np.float64, copy=True)
    if do_flip:
        joints[:, 0] = img_w - joints[:, 0] - 1
        for a, b in flip_pairs:
            joints[[a, b]] = joints[[b, a]]
            joints_vis[[a, b]] = joints_vis[[b, a]]
    for j in range(len(joints)):
        joints[j, 0:2] = trans @ np.array([joints[j, 0], joints[j, 1], 1.0])
        den = (width * scale) if depth_in_image else (rect_3d_width * scale)
        joints[j, 2] = joints[j, 2] / den * patch_width
    joints[:, 0] = joints[:, 0] / patch_width - 0.5
    joints[:, 1] = joints[:, 1] / patch_height - 0.5
    joints[:, 2] = joints[:, 2] / patch_width
    return t, joints.reshape(-1), joints_vis.reshape(-1), trans
