"""Hard-argmax keypoint decode -- host-side mirror of the reference
lib/core/inference.py:12-68 (`get_max_preds`, `get_final_preds`), with the
per-(n,j) argmax running in the warp-shuffle kernel epb_argmax2d (first-index
tie-break == numpy.argmax; indices are bit-exact).  numpy in / numpy out like
the reference; `get_max_preds_device` is the tensor-in / tensor-out variant."""
import numpy as np
import torch

from epipolarpose_b200 import ops as _ops

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


def get_final_preds_device(heatmaps, center, scale, post_process=True):
    """heatmaps [N,J,H,W] float32 (device), center / scale [N,2] -> (preds [N,J,2] f32 image
    coordinates, maxvals [N,J,1] f32) on the device: argmax, +-0.25 px refinement and the
    heat-map -> image affine in ONE launch (epb_final_preds)."""
    ops = _backend[0]
    hm = heatmaps.contiguous()
    N, J, H, W = hm.shape
    dev = hm.device
    c = torch.as_tensor(np.asarray(center, dtype=np.float64).reshape(N, 2)).to(dev)
    sc = torch.as_tensor(np.asarray(scale, dtype=np.float64).reshape(N, 2)).to(dev)
    preds = torch.empty((N, J, 2), device=dev, dtype=torch.float32)
    maxvals = torch.empty((N, J, 1), device=dev, dtype=torch.float32)
    if N * J:
        ops.final_preds(hm, N, J, H, W, c, sc, post_process, preds, maxvals)
    return preds, maxvals


def get_final_preds(config, batch_heatmaps, center, scale):
    """reference :43-68 (numpy in / numpy out)."""
    assert isinstance(batch_heatmaps, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
    dev = torch.device("cuda") if _backend[0] is _ops else torch.device("cpu")
    hm = torch.from_numpy(np.ascontiguousarray(batch_heatmaps, dtype=np.float32)).to(dev)
    scale = np.stack([np.asarray(s_, dtype=np.float64).reshape(-1)[:2] if np.ndim(s_) else
                      np.array([s_, s_], dtype=np.float64) for s_ in scale])
    preds, maxvals = get_final_preds_device(hm, np.asarray(center), scale,
                                            config.TEST.POST_PROCESS)
    return preds.cpu().numpy(), maxvals.cpu().numpy().astype(batch_heatmaps.dtype)
