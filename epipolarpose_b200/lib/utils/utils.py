"""Logger / optimiser / checkpoint helpers -- host-side mirror of the reference
lib/utils/utils.py:13-69,199-214 (`create_logger`, `get_optimizer`,
`save_checkpoint`, `AverageMeter`).  `get_optimizer` returns fused optimisers
whose update is ONE sm_100a kernel over a flat parameter buffer
(epb_adam_step / epb_sgd_step) with torch.optim semantics: Adam(lr) with betas
(0.9, 0.999), eps 1e-8 and -- like the reference :56-60 -- NO weight decay;
SGD(lr, momentum, weight_decay, nesterov).  They subclass
torch.optim.Optimizer so MultiStepLR (scripts/train.py:107-109) and
state_dict()/load_state_dict() keep working."""
import logging
import os
import time

import torch

from epipolarpose_b200 import ops as _ops

_backend = [_ops]


def create_logger(cfg, cfg_name, phase='train'):
    """reference :13-42: <OUTPUT_DIR>/<dataset>/<model>/<EXP_NAME>/ + log file."""
    from ..core.config import get_model_name
    root = cfg.OUTPUT_DIR
    os.makedirs(root, exist_ok=True)
    dataset = cfg.DATASET.DATASET + '_' + cfg.DATASET.HYBRID_JOINTS_TYPE \
        if cfg.DATASET.HYBRID_JOINTS_TYPE else cfg.DATASET.DATASET
    model, _ = get_model_name(cfg)
    out_dir = os.path.join(root, dataset, model, cfg.EXP_NAME)
    os.makedirs(out_dir, exist_ok=True)
    stamp = time.strftime('%Y-%m-%d-%H-%M')
    log_file = os.path.join(out_dir, '{}_{}_{}.log'.format(
        os.path.basename(cfg_name).split('.')[0], stamp, phase))
    logging.basicConfig(filename=str(log_file), format='%(asctime)-15s %(message)s')
    logger = logging.getLogger()
    logger.setLevel(logging.INFO)
    logging.getLogger('').addHandler(logging.StreamHandler())
    return logger, str(out_dir)


class _FlatOptimizer(torch.optim.Optimizer):
    """Flattens every group's parameters into one buffer (4-float aligned
    slots, same layout as the gradient buffer PoseResNet's backward emits) so
    the update is a single kernel and the gradient all-reduce a single call."""

    def __init__(self, params, defaults):
        super().__init__(params, defaults)
        self._flat = []
        for group in self.param_groups:
            ps = [p for p in group['params']]
            sizes = [p.numel() for p in ps]
            offs, tot = [], 0
            for s in sizes:
                offs.append(tot)
                tot += (s + 3) // 4 * 4
            dev = ps[0].device if ps else torch.device('cpu')
            flat = torch.zeros(tot, device=dev, dtype=torch.float32)
            for p, o, s in zip(ps, offs, sizes):
                flat[o:o + s].copy_(p.data.reshape(-1))
                p.data = flat[o:o + s].view(p.shape)
            self._flat.append({'buf': flat, 'offs': offs, 'sizes': sizes, 'n': tot})

    def _grads_are_flat(self, group, info):
        ps = group['params']
        if not ps or any(p.grad is None for p in ps):
            return None
        g0 = ps[0].grad
        base = g0.data_ptr()
        for p, o in zip(ps, info['offs']):
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != base + 4 * o:
                return None
        try:
            st = g0.untyped_storage()
            start = (base - st.data_ptr()) // 4
            if start + info['n'] > st.nbytes() // 4:
                return None
            return torch.empty(0, device=g0.device, dtype=torch.float32).set_(st, start, (info['n'],))
        except Exception:
            return None

    def flat_params(self):
        return [f['buf'] for f in self._flat]


class FusedAdam(_FlatOptimizer):
    """torch.optim.Adam semantics; hyper-parameters and the step count live in device
    memory (epb_adam_step_dev) so that a captured CUDA graph of the training step keeps
    following the LR schedule."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _hyper(self, group):
        b1, b2 = group['betas']
        return [group['lr'], b1, b2, group['eps'], group['weight_decay'], 1.0]

    def sync_hyper(self):
        """Push host-side hyper-parameters (e.g. after lr_scheduler.step()) to the device."""
        for gi, (group, info) in enumerate(zip(self.param_groups, self._flat)):
            st = self.state.setdefault('flat%d' % gi, {})
            h = self._hyper(group)
            if st.get('hyper_host') != h:
                dev = info['buf'].device
                if 'hyper' not in st:
                    st['hyper'] = torch.zeros(len(h), device=dev, dtype=torch.float32)
                st['hyper'].copy_(torch.tensor(h, dtype=torch.float32), non_blocking=True)
                st['hyper_host'] = h

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        ops = _backend[0]
        capturing = info_dev_is_capturing()
        if not capturing:
            self.sync_hyper()
        for gi, (group, info) in enumerate(zip(self.param_groups, self._flat)):
            st = self.state.setdefault('flat%d' % gi, {})
            if 'exp_avg' not in st:
                st['step'] = 0
                st['step_dev'] = torch.zeros(1, device=info['buf'].device, dtype=torch.int32)
                st['exp_avg'] = torch.zeros_like(info['buf'])
                st['exp_avg_sq'] = torch.zeros_like(info['buf'])
            st['step'] += 1
            st['step_dev'] += 1
            b1, b2 = group['betas']
            gflat = self._grads_are_flat(group, info)
            if gflat is not None and info['n'] % 4 == 0:
                ops.adam_step_dev(info['buf'], gflat, st['exp_avg'], st['exp_avg_sq'], info['n'],
                                  st['hyper'], st['step_dev'])
                continue
            for p, o, s in zip(group['params'], info['offs'], info['sizes']):
                if p.grad is None:
                    continue
                ops.adam_step(info['buf'][o:o + s], p.grad.contiguous().reshape(-1),
                              st['exp_avg'][o:o + s], st['exp_avg_sq'][o:o + s], s, group['lr'],
                              b1, b2, group['eps'], group['weight_decay'], st['step'])
        return loss


class FusedSGD(_FlatOptimizer):
    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0, nesterov=False):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay,
                                      nesterov=nesterov))

    def _hyper(self, group):
        return [group['lr'], group['momentum'], group['weight_decay'],
                1.0 if group['nesterov'] else 0.0, 1.0]

    sync_hyper = FusedAdam.sync_hyper

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        ops = _backend[0]
        if not info_dev_is_capturing():
            self.sync_hyper()
        for gi, (group, info) in enumerate(zip(self.param_groups, self._flat)):
            st = self.state.setdefault('flat%d' % gi, {})
            if 'buf' not in st:
                st['step'] = 0
                st['step_dev'] = torch.zeros(1, device=info['buf'].device, dtype=torch.int32)
                st['buf'] = torch.zeros_like(info['buf'])
            st['step'] += 1
            st['step_dev'] += 1
            gflat = self._grads_are_flat(group, info)
            if gflat is not None:
                ops.sgd_step_dev(info['buf'], gflat, st['buf'], info['n'], st['hyper'], st['step_dev'])
                continue
            args = (group['lr'], group['momentum'], group['weight_decay'], group['nesterov'],
                    st['step'] == 1)
            for p, o, s in zip(group['params'], info['offs'], info['sizes']):
                if p.grad is None:
                    continue
                ops.sgd_step(info['buf'][o:o + s], p.grad.contiguous().reshape(-1),
                             st['buf'][o:o + s], s, *args)
        return loss


def info_dev_is_capturing():
    try:
        return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
    except Exception:
        return False


def get_optimizer(cfg, model):
    """reference :45-61."""
    optimizer = None
    params = [p for p in model.parameters()]
    if cfg.TRAIN.OPTIMIZER == 'sgd':
        optimizer = FusedSGD(params, lr=cfg.TRAIN.LR, momentum=cfg.TRAIN.MOMENTUM,
                             weight_decay=cfg.TRAIN.WD, nesterov=cfg.TRAIN.NESTEROV)
    elif cfg.TRAIN.OPTIMIZER == 'adam':
        optimizer = FusedAdam(params, lr=cfg.TRAIN.LR)
    return optimizer


def save_checkpoint(states, is_best, output_dir, filename='checkpoint.pth.tar'):
    """reference :64-69."""
    torch.save(states, os.path.join(output_dir, filename))
    if is_best and 'state_dict' in states:
        torch.save(states['state_dict'], os.path.join(output_dir, 'model_best.pth.tar'))


class AverageMeter(object):
    """reference :199-214."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count if self.count != 0 else 0
