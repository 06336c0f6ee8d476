"""H36M evaluation protocol on the device -- host-side mirror of the arithmetic in the
reference lib/dataset/h36m.py::H36M_Integral.evaluate (:168-378): back-projection
(lib/utils/prep_h36m.py:85-89), similarity alignment with optimal scale
(compute_similarity_transform, :108-168), root alignment, per-joint errors and the nine
protocol means -- one sm_100a kernel launch (epb_h36m_eval, float64, one thread per sample)
instead of the per-sample / per-joint Python loops.

`evaluate_h36m` takes the arrays the reference reads from its db records; datasets call it
from their `.evaluate(preds, save_path, debug)` and return `(name_value, perf)` unchanged."""
import numpy as np
import torch

from epipolarpose_b200 import ops as _ops

_backend = [_ops]     # test hook: tests may swap in the CPU emulation of the C ABI

H36M_NAMES = ['Hip', 'RHip', 'RKnee', 'RFoot', 'LHip', 'LKnee', 'LFoot', 'Spine', 'Thorax',
              'Neck/Nose', 'Head', 'LShoulder', 'LElbow', 'LWrist', 'RShoulder', 'RElbow', 'RWrist']
MPII_NAMES = ['RFoot', 'RKnee', 'RHip', 'LHip', 'LKnee', 'LFoot', 'Hip', 'Thorax', 'Neck/Nose',
              'Head', 'RWrist', 'RElbow', 'RShoulder', 'LShoulder', 'LElbow', 'LWrist']
H36M_TO_MPII_PERM = np.array([H36M_NAMES.index(h) for h in MPII_NAMES if h != '' and h in H36M_NAMES])
_J14 = {True: [0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 12, 13, 14, 15],        # h36m.py:186
        False: [0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 15]}
METRIC_NAMES = ['hm36_17j      :', 'hm36_17j_align:', 'hm36_17j_norm:', 'hm36_17j_14   :',
                'hm36_17j_14_al:', 'hm36_17j_14_nm:', 'hm36_17j_x    :', 'hm36_17j_y    :',
                'hm36_17j_z    :']
PCK_THRESHOLD = 150.0                                                      # h36m.py:283


def _device():
    return torch.device("cuda") if _backend[0] is _ops else torch.device("cpu")


def evaluate_h36m(preds, gt_joints_3d, pelvis, fl, c_p, mpii_order=False, return_poses=False):
    """preds [S,J,>=3] image-space predictions (x, y px, root-relative depth mm; extra columns
    ignored, h36m.py:169); gt_joints_3d [S,17,3] (`joints_3d`, H36M order -- permuted here when
    mpii_order, h36m.py:223-224); pelvis [S,3] or [S] (depth), fl / c_p [S,>=2].
    Returns (name_value, perf, details) -- name_value / perf exactly as the reference returns,
    details = dict(metrics [S,9], per_joint [S,J], pck [S,J], poses [S,J,9] or None)."""
    ops = _backend[0]
    dev = _device()
    preds = np.asarray(preds, dtype=np.float64)
    if preds.ndim != 3 or preds.shape[2] < 3:
        raise ValueError("preds must be [S, J, >=3]")
    S, J = preds.shape[0], preds.shape[1]
    gt = np.asarray(gt_joints_3d, dtype=np.float64)
    if mpii_order:
        gt = gt[:, H36M_TO_MPII_PERM, :]
    if gt.shape[:2] != (S, J):
        raise ValueError("ground truth %s does not match predictions %s" % (gt.shape, preds.shape))
    pz = np.asarray(pelvis, dtype=np.float64)
    pz = pz[:, 2] if pz.ndim == 2 else pz
    cam = np.concatenate([np.asarray(fl, dtype=np.float64)[:, 0:2],
                          np.asarray(c_p, dtype=np.float64)[:, 0:2], pz.reshape(S, 1)], axis=1)
    root = 6 if mpii_order else 0
    mask = 0
    for j in _J14[bool(mpii_order)]:
        if j < J:
            mask |= 1 << j
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    metrics = torch.empty((S, 9), device=dev, dtype=torch.float64)
    per_joint = torch.empty((S, J), device=dev, dtype=torch.float64)
    pck = torch.empty((S, J), device=dev, dtype=torch.int32)
    poses = torch.empty((S, J, 9), device=dev, dtype=torch.float64) if return_poses else None
    if S:
        ops.h36m_eval(t(preds[:, :, 0:3]), t(gt), t(cam), S, J, root, mask, PCK_THRESHOLD, metrics,
                      per_joint, pck, poses)
    means = metrics.mean(dim=0).cpu().numpy() if S else np.zeros(9)
    name_value = list(zip(METRIC_NAMES, [float(v) for v in means]))
    details = dict(metrics=metrics.cpu().numpy(), per_joint=per_joint.cpu().numpy(),
                   pck=pck.cpu().numpy(), poses=None if poses is None else poses.cpu().numpy())
    return name_value, float(means[0]), details
