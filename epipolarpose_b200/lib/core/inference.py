"""Hard-argmax keypoint decode -- host-side mirror of the reference
lib/core/inference.py:12-68 (`get_max_preds`, `get_final_preds`), with the
per-(n,j) argmax running in the warp-shuffle kernel epb_argmax2d (first-index
tie-break == numpy.argmax; indices are bit-exact).  numpy in / numpy out like
the reference; `get_max_preds_device` is the tensor-in / tensor-out variant."""
import math

import numpy as np
import torch

from epipolarpose_b200 import ops as _ops
from ..utils.transforms import transform_preds

_backend = [_ops]


def get_max_preds_device(heatmaps):
    """heatmaps [N,J,H,W] float32 CUDA contiguous -> (preds [N,J,2] f32,
    maxvals [N,J,1] f32, idx [N,J] int32) on the device."""
    ops = _backend[0]
    assert heatmaps.dim() == 4, 'batch_images should be 4-ndim'
    hm = heatmaps.contiguous()
    N, J, H, W = hm.shape
    idx = torch.empty((N, J), device=hm.device, dtype=torch.int32)
    maxvals = torch.empty((N, J, 1), device=hm.device, dtype=torch.float32)
    preds = torch.empty((N, J, 2), device=hm.device, dtype=torch.float32)
    if N * J:
        ops.argmax2d(hm, N * J, H, W, idx, maxvals, preds)
    return preds, maxvals, idx


def get_max_preds(batch_heatmaps):
    """reference :12-40 (numpy [N,J,H,W] -> preds [N,J,2] f32, maxvals [N,J,1])."""
    assert isinstance(batch_heatmaps, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
    assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    dev = torch.device("cuda") if _backend[0] is _ops else torch.device("cpu")
    hm = torch.from_numpy(np.ascontiguousarray(batch_heatmaps, dtype=np.float32)).to(dev)
    preds, maxvals, _ = get_max_preds_device(hm)
    return preds.cpu().numpy(), maxvals.cpu().numpy().astype(batch_heatmaps.dtype)


def get_final_preds(config, batch_heatmaps, center, scale):
    """reference :43-68."""
    coords, maxvals = get_max_preds(batch_heatmaps)
    h, w = batch_heatmaps.shape[2], batch_heatmaps.shape[3]
    if config.TEST.POST_PROCESS:        # +-0.25 px toward the higher neighbour (:49-61)
        for n in range(coords.shape[0]):
            for p in range(coords.shape[1]):
                hm = batch_heatmaps[n][p]
                px = int(math.floor(coords[n][p][0] + 0.5))
                py = int(math.floor(coords[n][p][1] + 0.5))
                if 1 < px < w - 1 and 1 < py < h - 1:
                    diff = np.array([hm[py][px + 1] - hm[py][px - 1],
                                     hm[py + 1][px] - hm[py - 1][px]])
                    coords[n][p] += np.sign(diff) * .25
    preds = coords.copy()
    for i in range(coords.shape[0]):
        preds[i] = transform_preds(coords[i], center[i], scale[i], [w, h])
    return preds, maxvals
