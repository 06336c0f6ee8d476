"""Two-view triangulation -- host-side mirror of the reference
lib/utils/triangulation.py function surface:
    f(u1[J,2], P1[>=3,4], u2[J,2], P2[>=3,4]) -> (x[J,3] of output_dtype, status[J])
for linear_eigen_triangulation (:8-27, homogeneous DLT = cv2.triangulatePoints),
linear_LS_triangulation (:34-97) and iterative_LS_triangulation (:104-181), plus
set_triangl_output_dtype (:226-232).  The arithmetic (float64 one-sided Jacobi
SVD per joint) runs in the sm_100a kernel epb_triangulate; numpy in / numpy out
like the reference.  `triangulate_pairs` is the batched tensor API the training
loop uses (no host round trip).  polynomial_triangulation (:184-220: fundamental matrix
from the projection matrices, cv2.correctMatches = Hartley-Sturm optimal correction, then
the homogeneous DLT) runs as method "polynomial" of the same kernel, including the reference's
8-point fallback (:215-217): when the correction is NaN for every joint of a pair, F is
re-estimated from the matches (cv2.findFundamentalMat FM_8POINT) on the device and the
correction repeated; "polynomial_8point" runs that branch unconditionally."""
import numpy as np
import torch

from epipolarpose_b200 import ops as _ops

_backend = [_ops]
METHODS = {"linear_eigen": 0, "linear_LS": 1, "iterative_LS": 2, "polynomial": 3,
           "polynomial_8point": 4,
           "eigen": 0, "ls": 1, "iterative": 2}

output_dtype = float


def set_triangl_output_dtype(output_dtype_):
    global output_dtype
    output_dtype = output_dtype_


def triangulate_pairs(u1, u2, P1, P2, method="iterative", tolerance=3.e-5, stride_u=None):
    """u1,u2 [NP,J,S>=2] float64 (first two columns used), P1,P2 [NP,3,4] float64,
    all on the device -> (X [NP,J,3] float64, status [NP,J] int32)."""
    ops = _backend[0]
    NP, J = u1.shape[0], u1.shape[1]
    S = u1.shape[2] if stride_u is None else stride_u
    X = torch.empty((NP, J, 3), device=u1.device, dtype=torch.float64)
    status = torch.empty((NP, J), device=u1.device, dtype=torch.int32)
    if NP * J:
        ops.triangulate(u1.contiguous(), u2.contiguous(), S, P1.contiguous(), P2.contiguous(),
                        NP, J, METHODS[method], tolerance, X, status)
    return X, status


def triangulate_views(u, P):
    """V-view homogeneous DLT (2 <= V <= 4; not in the reference):
    u [NT,V,J,S>=2] float64, P [NT,V,3,4] float64 on the device -> (X [NT,J,3], status [NT,J])."""
    ops = _backend[0]
    NT, V, J = u.shape[0], u.shape[1], u.shape[2]
    if not 2 <= V <= 4:
        raise ValueError("triangulate_views handles 2..4 views per tuple, got %d" % V)
    X = torch.empty((NT, J, 3), device=u.device, dtype=torch.float64)
    status = torch.empty((NT, J), device=u.device, dtype=torch.int32)
    if NT * J:
        ops.triangulate_nview(u.contiguous(), u.shape[3], P.reshape(NT, V, 12).contiguous(), NT, V, J, X, status)
    return X, status


def _device():
    return torch.device("cuda") if _backend[0] is _ops else torch.device("cpu")


def _run(u1, P1, u2, P2, method, tolerance=3.e-5):
    u1 = np.ascontiguousarray(u1, dtype=np.float64)
    u2 = np.ascontiguousarray(u2, dtype=np.float64)
    assert u1.ndim == 2 and u1.shape == u2.shape and u1.shape[1] >= 2
    dev = _device()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    X, st = triangulate_pairs(t(u1)[None], t(u2)[None], t(np.asarray(P1)[0:3, 0:4])[None],
                              t(np.asarray(P2)[0:3, 0:4])[None], method, tolerance)
    return X[0].cpu().numpy(), st[0].cpu().numpy()


def linear_eigen_triangulation(u1, P1, u2, P2, max_coordinate_value=1.e16):
    x, st = _run(u1, P1, u2, P2, "linear_eigen")
    if max_coordinate_value != 1.e16:
        with np.errstate(invalid="ignore"):
            st = np.max(np.abs(x), axis=1) <= max_coordinate_value
    return x.astype(output_dtype), st.astype(bool)


def linear_LS_triangulation(u1, P1, u2, P2):
    x, _ = _run(u1, P1, u2, P2, "linear_LS")
    return x.astype(output_dtype), np.ones(len(u1), dtype=bool)


def iterative_LS_triangulation(u1, P1, u2, P2, tolerance=3.e-5):
    x, st = _run(u1, P1, u2, P2, "iterative_LS", tolerance)
    return x.astype(output_dtype), st.astype(int)


def polynomial_triangulation(u1, P1, u2, P2):
    x, st = _run(u1, P1, u2, P2, "polynomial")
    return x.astype(output_dtype), st.astype(bool)
