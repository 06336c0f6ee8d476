"""Train / validate / evaluate loops -- host-side mirror of the reference
lib/core/function.py signatures: train_integral(config, train_loader, model,
criterion, optimizer, epoch) (:14-63), validate_integral(val_loader, model)
(:66-110), eval_integral(epoch, preds, val_loader, path, debug) (:113-135).

Differences from the reference body (same observable behaviour):
  * the loss value is kept on the device and read back only when a log line
    is printed (the reference's per-step loss.item() sync, :48, is gone);
  * when config.TRAIN.ONLINE_TRIANGULATION is set the labels are produced per
    batch by the epipolar self-supervision path (lib/utils/img_utils.py
    self_supervision: soft-argmax -> patch->image -> two-view triangulation ->
    re-projection) entirely on the device -- the glue the released reference
    leaves unwired;
  * the gradient all-reduce for multi-GPU data parallelism happens inside the
    model's backward (one NCCL call on the flat gradient buffer).
"""
import logging
import time

import numpy as np
import torch

from ..utils.img_utils import (self_supervision_device,
                               trans_coords_from_patch_to_org_3d_batch)
from .integral_loss import get_result_func
from ..utils.utils import AverageMeter

logger = logging.getLogger(__name__)


class _fused_head:
    """Inside the training step nobody but the criterion consumes the logits: let the criterion hand
    their gradient to the network's backward as split planes (_sinks.py).  Off elsewhere, so that
    user code that inspects the logit gradient (retain_grad, hooks, autograd.grad) sees fp32 values."""

    def __init__(self, model):
        net = getattr(model, "module", model)
        self.net = net if hasattr(net, "fused_head_gradient") else None

    def __enter__(self):
        if self.net is not None:
            self.prev, self.net.fused_head_gradient = self.net.fused_head_gradient, True

    def __exit__(self, *exc):
        if self.net is not None:
            self.net.fused_head_gradient = self.prev


def _online_tri(config):
    train = getattr(config, 'TRAIN', None)
    return bool(train is not None and getattr(train, 'ONLINE_TRIANGULATION', False))


def online_epipolar_loss(criterion, preds, meta, method="iterative"):
    """criterion(preds, labels(preds), 1) with the labels produced by the epipolar
    self-supervision path from the SAME soft-argmax coordinates the loss uses
    (labels carry no gradient, reference integral_loss.py:88-91)."""
    from .integral_loss import softmax_integral_tensor, _WeightedLossFn
    from ..utils.img_utils import (patch_to_image_device, triangulate_device,
                                   labels_from_global_coords_device)
    J = criterion.num_joints
    W, H = preds.shape[-1], preds.shape[-2]
    D = preds.shape[-3] // J
    coords = softmax_integral_tensor(preds, J, True, W, H, D)
    with torch.no_grad():
        kps = patch_to_image_device(coords.detach(), meta)
        X = triangulate_device(kps, meta, method)
        label, weight = labels_from_global_coords_device(X, meta)
    return _WeightedLossFn.apply(coords, label, weight, criterion._kind, criterion.size_average,
                                 criterion.norm)


class GraphedTrainStep:
    """One training step (forward, loss, backward incl. the gradient all-reduce, optimiser)
    captured ONCE in a CUDA graph and replayed: the step is ~3000 kernel launches whose
    Python/driver issue time is otherwise comparable to the GPU time.  Inputs are copied
    into static device buffers; hyper-parameters and the step counter live in device
    memory (FusedAdam / FusedSGD), so LR schedules keep working.  The first call runs
    eagerly (sizes workspaces, one-time attributes), the second captures and replays."""

    def __init__(self, model, criterion, optimizer, online=False, method="iterative"):
        # scripts/train.py:94 wraps the model in nn.DataParallel(device_ids=[k]); with one
        # device id that wrapper only forwards the call (and its scatter is not capturable)
        if isinstance(model, torch.nn.DataParallel) and len(model.device_ids) == 1:
            model = model.module
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.online, self.method = online, method
        self.graph = None
        self.key = None
        self.warm_key = None          # shape of the last eager (warm-up) step
        self.failed = False
        self.calls = 0
        self.copy_stream, self.pending, self.stage_next = None, None, 0
        # eager warm-up and capture share ONE side stream so that the parameters'
        # AccumulateGrad nodes are never bound to the legacy default stream (which
        # may not join a capture)
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None

    def eager_step(self, x, label, weight, geom):
        """One un-captured step on the stepper's side stream."""
        if self.stream is None:
            return self._eager(x, label, weight, geom)
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            out = self._eager(x, label, weight, geom)
        cur.wait_stream(self.stream)
        return out

    def _eager(self, x, label, weight, geom):
        self.optimizer.zero_grad()
        with _fused_head(self.model):
            preds = self.model(x)
        if self.online:
            loss = online_epipolar_loss(self.criterion, preds, {"_packed": geom}, self.method)
        else:
            loss = self.criterion(preds, label, weight)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    # ---- input staging: the H2D copy of batch i+1 overlaps the replay of step i
    def prefetch(self, batch_data):
        """Start copying a (pinned) host batch into the idle device staging buffer on a copy
        stream; the next __call__ with the SAME tensor object hands it over to the graph's
        static input with a device-side copy (100 MB at HBM speed instead of PCIe speed in
        front of the replay)."""
        if self.graph is None or not torch.cuda.is_available() or batch_data.is_cuda:
            return
        if tuple(batch_data.shape) != tuple(self.sx.shape):
            return
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream()
            self.stage = [torch.empty_like(self.sx) for _ in range(2)]
            self.stage_evt = [torch.cuda.Event() for _ in range(2)]
            self.stage_free = [None, None]
        k = self.stage_next
        with torch.cuda.stream(self.copy_stream):
            if self.stage_free[k] is not None:
                self.copy_stream.wait_event(self.stage_free[k])    # its previous hand-over is done
            self.stage[k].copy_(batch_data, non_blocking=True)
            self.stage_evt[k].record(self.copy_stream)
        self.pending = (batch_data, k)
        self.stage_next = 1 - k

    def _load_input(self, batch_data):
        if self.pending is not None and self.pending[0] is batch_data:
            k = self.pending[1]
            self.pending = None
            cur = torch.cuda.current_stream()
            cur.wait_event(self.stage_evt[k])
            self.sx.copy_(self.stage[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cur)
            self.stage_free[k] = ev
        else:
            self.sx.copy_(batch_data, non_blocking=True)

    def __call__(self, batch_data, label=None, weight=None, meta=None):
        dev = next(self.model.parameters()).device
        B = batch_data.shape[0]
        geom = None
        if self.online:
            from ..utils.img_utils import pack_meta
            geom = pack_meta(meta, B, dev)
        key = (tuple(batch_data.shape), self.online)
        self.calls += 1
        if self.graph is not None and key == self.key:
            self._load_input(batch_data)
            if self.online:
                for k in self.sgeom:
                    self.sgeom[k].copy_(geom[k], non_blocking=True)
            else:
                self.slabel.copy_(label, non_blocking=True)
                self.sweight.copy_(weight, non_blocking=True)
            if hasattr(self.optimizer, "sync_hyper"):
                self.optimizer.sync_hyper()
            self.graph.replay()
            return self.sloss
        x = batch_data.to(dev, non_blocking=True)
        if not self.online:
            label, weight = label.to(dev, non_blocking=True), weight.to(dev, non_blocking=True)
        # eager when: first call (sizes every scratch buffer), a shape other than the warmed-up
        # one (ragged last batch), no CUDA, or a capture that failed before
        if self.graph is not None or self.failed or not torch.cuda.is_available() \
                or self.warm_key != key:
            self.warm_key = key
            return self.eager_step(x, label, weight, geom)
        # second call with the warmed-up shape: capture, then replay (capture does not execute)
        sx = x.clone()
        sgeom = {k: v.clone() for k, v in geom.items()} if self.online else None
        slabel = label.clone() if not self.online else None
        sweight = weight.clone() if not self.online else None
        if hasattr(self.optimizer, "sync_hyper"):
            self.optimizer.sync_hyper()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        try:
            with torch.cuda.graph(graph, stream=self.stream):
                sloss = self._eager(sx, slabel, sweight, sgeom)
        except Exception as e:                     # non-capturable optimiser / wrapper / op
            logger.warning("CUDA-graph capture of the training step failed (%s); running eagerly", e)
            self.failed = True
            torch.cuda.synchronize()
            self.optimizer.zero_grad(set_to_none=True)
            return self.eager_step(x, label, weight, geom)
        # published only after a successful capture
        self.sx, self.sgeom, self.slabel, self.sweight, self.sloss = sx, sgeom, slabel, sweight, sloss
        self.graph, self.key = graph, key
        self.graph.replay()
        return self.sloss


def _graph_capable(model, criterion, optimizer):
    """The captured step needs the fused optimisers (hyper-parameters in device memory) and a
    single-device model: torch.optim.Adam / SGD steps and nn.DataParallel scatter (more than
    one device id: the reference's multi-GPU path, scripts/train.py:94) are not capturable."""
    from ..utils.utils import _FlatOptimizer
    if not hasattr(criterion, '_kind') or not isinstance(optimizer, _FlatOptimizer):
        return False
    if isinstance(model, torch.nn.DataParallel) and len(model.device_ids) != 1:
        return False
    return True


def train_integral(config, train_loader, model, criterion, optimizer, epoch):
    batch_time = AverageMeter()
    data_time = AverageMeter()
    losses = AverageMeter()
    model.train()
    online = _online_tri(config)
    method = getattr(config.TRAIN, 'TRIANGULATION_METHOD', 'iterative') if online else None
    pending = []           # (device loss, batch size) not yet folded into `losses`
    use_graph = bool(getattr(config.TRAIN, 'CUDA_GRAPH', True)) and \
        _graph_capable(model, criterion, optimizer)
    stepper = getattr(model, '_epb_graphed_step', None)
    if use_graph and (stepper is None or stepper.optimizer is not optimizer
                      or stepper.criterion is not criterion or stepper.online != online):
        stepper = GraphedTrainStep(model, criterion, optimizer, online, method)
        model._epb_graphed_step = stepper
    end = time.time()
    it = iter(train_loader)
    nxt = next(it, None)
    i = -1
    while nxt is not None:
        data, i = nxt, i + 1
        nxt = next(it, None)
        data_time.update(time.time() - end)
        batch_data, batch_label, batch_label_weight, meta = data
        batch_size = batch_data.size(0)
        if use_graph:
            if stepper.pending is None and stepper.graph is not None:
                stepper.prefetch(batch_data)              # first replayed batch of this call
            loss = stepper(batch_data, batch_label, batch_label_weight, meta)
            if stepper.graph is not None:
                loss = loss.clone()      # the static loss buffer is overwritten by the next replay
                if nxt is not None:
                    stepper.prefetch(nxt[0])     # H2D of the next batch overlaps this step
            pending.append((loss, batch_size))
            del loss
        else:
            optimizer.zero_grad()
            batch_data = batch_data.cuda(non_blocking=True)
            with _fused_head(model):
                preds = model(batch_data)
            if online:
                # one soft-argmax pass serves both the epipolar labels and the loss
                loss = online_epipolar_loss(criterion, preds, meta, method)
                batch_label = batch_label_weight = None
            else:
                batch_label = batch_label.cuda(non_blocking=True)
                batch_label_weight = batch_label_weight.cuda(non_blocking=True)
                loss = criterion(preds, batch_label, batch_label_weight)
            del batch_data, batch_label, batch_label_weight, preds
            loss.backward()
            optimizer.step()
            pending.append((loss.detach(), batch_size))
            del loss
        if i % config.PRINT_FREQ == 0:
            for lv, bs in pending:
                losses.update(lv.item(), bs)      # the only device sync of the loop
            pending = []
            batch_time.update(time.time() - end)
            msg = 'Epoch: [{0}][{1}/{2}]\t' \
                  'Time {batch_time.val:.3f}s ({batch_time.avg:.3f}s)\t' \
                  'Speed {speed:.1f} samples/s\t' \
                  'Data {data_time.val:.3f}s ({data_time.avg:.3f}s)\t' \
                  'Loss {loss.val:.5f} ({loss.avg:.5f})'.format(
                      epoch, i, len(train_loader), batch_time=batch_time,
                      speed=batch_size / max(batch_time.val, 1e-9),
                      data_time=data_time, loss=losses)
            logger.info(msg)
        else:
            batch_time.update(time.time() - end)
        end = time.time()
    for lv, bs in pending:
        losses.update(lv.item(), bs)
    return losses.avg


def validate_integral(val_loader, model):
    print("Validation stage")
    result_func = get_result_func()
    model.eval()
    chunks = []
    with torch.no_grad():
        for i, data in enumerate(val_loader):
            batch_data = data[0].cuda(non_blocking=True)
            preds = model(batch_data)
            chunks.append(result_func(256, 256, preds))     # hard-coded 256 as reference :87
            del preds, batch_data
    if not chunks:
        return np.zeros((0, 0, 4))
    out = np.concatenate(chunks, axis=0)                   # ragged last batch handled
    return out[0:len(val_loader.dataset)]


def eval_integral(epoch, preds_in_patch_with_score, val_loader, final_output_path, debug=False):
    print("Evaluation stage")
    imdb_list = val_loader.dataset.db
    imdb = val_loader.dataset
    n = len(val_loader.dataset)
    get = lambda k: np.array([imdb_list[s][k] for s in range(n)], dtype=np.float64)
    preds_in_img_with_score = trans_coords_from_patch_to_org_3d_batch(
        np.asarray(preds_in_patch_with_score)[:n], get('center_x'), get('center_y'), get('width'),
        get('height'), 256, 256, 2000)
    name_value, perf = imdb.evaluate(preds_in_img_with_score.copy(), final_output_path, debug=debug)
    for name, value in name_value:
        logger.info('Epoch[%d] Validation-%s %f', epoch, name, value)
    return perf
