"""Self-supervision glue and patch<->image geometry -- host-side mirror of the
hot-path part of the reference lib/utils/img_utils.py:
  trans_coords_from_patch_to_org_3d (:150-155), self_supervision (:166-190),
  triangulate (:193-209), get_batch_labels_from_global_coords (:212-243).
Everything downstream of the network output runs on the device in float64
kernels (epb_patch_to_image -> epb_triangulate -> epb_project_labels); the
reference's per-sample / per-joint Python loops disappear.  `meta` follows the
dataset contract of reference lib/dataset/h36m.py:73-86 (collated: tensors or
lists of length B).  numpy in / numpy out like the reference;
`self_supervision_device` returns CUDA tensors for the training loop."""
import numpy as np
import torch

from epipolarpose_b200 import ops as _ops
from . import triangulation as _tri
from ..core.integral_loss import get_joint_location_coords

_backend = [_ops]


def _dev():
    return torch.device("cuda") if _backend[0] is _ops else torch.device("cpu")


def _t64(v, dev):
    if isinstance(v, torch.Tensor):
        return v.to(device=dev, dtype=torch.float64)
    return torch.as_tensor(np.asarray(v, dtype=np.float64), device=dev)


def _boxes(meta, B, dev):
    """[B,6] float64: c_x, c_y, width, height, scale, rot."""
    cols = [_t64(meta[k], dev).reshape(-1)[:B]
            for k in ('center_x', 'center_y', 'width', 'height', 'scale', 'rot')]
    return torch.stack(cols, dim=1).contiguous()


def _cams(meta, B, dev):
    """[B,16] float64: R(9) T(3) f(2) c(2)."""
    return torch.cat([_t64(meta['R'], dev).reshape(B, 9), _t64(meta['T'], dev).reshape(B, 3),
                      _t64(meta['f'], dev).reshape(B, 2), _t64(meta['c'], dev).reshape(B, 2)],
                     dim=1).contiguous()


def pack_meta(meta, B, dev):
    """Collated `meta` -> device float64 tensors {box [B,6], cam [B,16], P [B,3,4]} (the
    form the kernels consume; static-shaped, so a CUDA graph can re-read them)."""
    if isinstance(meta, dict) and "_packed" in meta:
        return meta["_packed"]
    out = {"box": _boxes(meta, B, dev)}
    if 'R' in meta:
        out["cam"] = _cams(meta, B, dev)
    if 'projection_matrix' in meta:
        out["P"] = _t64(meta['projection_matrix'], dev).reshape(B, -1, 4)[:, 0:3, :].contiguous()
    return out


def patch_to_image_device(coords_norm, meta, patch_w=256, patch_h=256, rect_3d_w=2000):
    """coords_norm [B, J*3] float32 (soft-argmax output) -> kps [B,J,4] float64."""
    ops = _backend[0]
    B = coords_norm.shape[0]
    J = coords_norm.shape[1] // 3
    dev = coords_norm.device
    kps = torch.empty((B, J, 4), device=dev, dtype=torch.float64)
    ops.patch_to_image(coords_norm.contiguous(), pack_meta(meta, B, dev)["box"], B, J, patch_w,
                       patch_h, rect_3d_w, kps)
    return kps


def triangulate_device(kps, meta, method="iterative"):
    """reference :193-209: sample i pairs with i + B/2; both halves receive the
    same world-frame result."""
    B, J = kps.shape[0], kps.shape[1]
    half = B // 2
    P = pack_meta(meta, B, kps.device)["P"]
    X, _ = _tri.triangulate_pairs(kps[:half], kps[half:2 * half], P[:half], P[half:2 * half],
                                  method=method, stride_u=kps.shape[2])
    return torch.cat([X, X], dim=0)


def labels_from_global_coords_device(X, meta, patch_w=256., patch_h=256., rect_3d_w=2000.):
    ops = _backend[0]
    B, J = X.shape[0], X.shape[1]
    dev = X.device
    label = torch.empty((B, J * 3), device=dev, dtype=torch.float32)
    weight = torch.empty((B, J * 3), device=dev, dtype=torch.float32)
    pm = pack_meta(meta, B, dev)
    ops.project_labels(X.contiguous(), pm["cam"], pm["box"], B, J, patch_w, patch_h, rect_3d_w,
                       label, weight)
    return label, weight


def self_supervision_device(preds, meta, method="iterative"):
    """preds: network output [B, J*D, H, W] (CUDA) -> (label, weight) CUDA f32
    [B, J*3]; labels carry no gradient (reference integral_loss.py:88-91)."""
    coords = get_joint_location_coords(preds)
    kps = patch_to_image_device(coords, meta)
    X = triangulate_device(kps, meta, method)
    return labels_from_global_coords_device(X, meta)


def self_supervision(preds, meta):
    """reference :166-190 -> numpy float32 (label, weight)."""
    label, weight = self_supervision_device(preds, meta)
    return label.cpu().numpy(), weight.cpu().numpy()


def triangulate(kps, meta):
    """reference :193-209, numpy [B,J,>=2] -> numpy [B,J,3] float64."""
    k = torch.as_tensor(np.ascontiguousarray(kps, dtype=np.float64), device=_dev())
    return triangulate_device(k, meta).cpu().numpy()


def get_batch_labels_from_global_coords(coords_3d_in_global_frame, meta):
    """reference :212-243 -> numpy float32 (label, weight)."""
    X = torch.as_tensor(np.ascontiguousarray(coords_3d_in_global_frame, dtype=np.float64),
                        device=_dev())
    label, weight = labels_from_global_coords_device(X, meta)
    return label.cpu().numpy(), weight.cpu().numpy()


def trans_coords_from_patch_to_org_3d(coords_in_patch, c_x, c_y, bb_width, bb_height,
                                      patch_width, patch_height, rect_3d_width, rect_3d_height,
                                      scale=1.0, rot=0):
    """reference :150-155 for ONE sample: [J,>=3] patch px -> image px (+ z in mm)."""
    return trans_coords_from_patch_to_org_3d_batch(
        np.asarray(coords_in_patch)[None], [c_x], [c_y], [bb_width], [bb_height], patch_width,
        patch_height, rect_3d_width, [scale], [rot])[0]


def trans_coords_from_patch_to_org_3d_batch(coords, c_x, c_y, bb_w, bb_h, patch_w, patch_h,
                                            rect_3d_w, scale=None, rot=None):
    """Batched form used by eval_integral: coords [B,J,>=3] (x,y,z in patch px)."""
    ops = _backend[0]
    coords = np.asarray(coords, dtype=np.float64)
    B, J = coords.shape[0], coords.shape[1]
    dev = _dev()
    # back to the normalised soft-argmax units the kernel consumes
    norm = np.empty((B, J, 3), dtype=np.float32)
    norm[:, :, 0] = coords[:, :, 0] / patch_w - 0.5
    norm[:, :, 1] = coords[:, :, 1] / patch_h - 0.5
    norm[:, :, 2] = coords[:, :, 2] / patch_w
    meta = {'center_x': c_x, 'center_y': c_y, 'width': bb_w, 'height': bb_h,
            'scale': scale if scale is not None else np.ones(B),
            'rot': rot if rot is not None else np.zeros(B)}
    kps = torch.empty((B, J, 4), device=dev, dtype=torch.float64)
    ops.patch_to_image(torch.from_numpy(norm.reshape(B, J * 3)).to(dev), _boxes(meta, B, dev), B, J,
                       float(patch_w), float(patch_h), float(rect_3d_w), kps)
    out = coords.copy()
    res = kps.cpu().numpy()
    out[:, :, 0:3] = res[:, :, 0:3]
    return out
