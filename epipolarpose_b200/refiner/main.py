"""Refiner training / evaluation loop -- host-side mirror of the reference refiner/main.py:
`parse_args` (:17-29), `train(model, train_dl, optimizer, glob_step, lr_now, criterion, args,
logger)` (:31-62), `test(model, test_dl)` (:64-84) and the `__main__` driver (:86-175), same
checkpoint dictionary (epoch / lr / step / err / state_dict / optimizer; optimizer state in the
torch.optim layout, so checkpoints interchange with the reference).

The model is refiner.model.LinearModelPG on the libepb.so kernels, the optimiser the fused Adam
(one kernel over the flat parameter buffer), gradient clipping refiner.utils.clip_grad_norm_ on
the device.  The loss value stays on the device and is folded into the meter at the end of the
epoch (the reference's per-step loss.item() sync, :54, is gone)."""
import argparse
import logging
import os
import time

import numpy as np
import torch
import torch.nn as nn

from .model import get_model, weight_init
from .utils import AverageMeter, clip_grad_norm_, lr_decay, save_ckpt


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--exp', type=str, default='test', help='ID of experiment')
    parser.add_argument('--load', type=str, default=None, help='path to load a pretrained checkpoint')
    parser.add_argument('--mode', type=str, default='train', help='mode: [train, test]')
    parser.add_argument('--num_epochs', type=int, default=200, help='num epochs')
    parser.add_argument('--lr', type=float, default=1e-3, help='learning rate')
    parser.add_argument('--lr_decay', type=int, default=100000, help='# steps of lr decay')
    parser.add_argument('--lr_gamma', type=float, default=0.96)
    return parser.parse_args(argv)


def train(model, train_dl, optimizer, glob_step, lr_now, criterion, args, logger):
    losses = AverageMeter()
    model.train()
    pending = []
    for i, (inp, tar) in enumerate(train_dl):
        glob_step += 1
        if glob_step % args.lr_decay == 0 or glob_step == 1:
            lr_now = lr_decay(optimizer, glob_step, args.lr, args.lr_decay, args.lr_gamma)
        inputs = inp.cuda(non_blocking=True) if torch.cuda.is_available() else inp
        targets = tar.cuda(non_blocking=True) if torch.cuda.is_available() else tar
        outputs = model(inputs)
        optimizer.zero_grad()
        loss = criterion(outputs[0], targets) + criterion(outputs[1], targets)
        pending.append((loss.detach(), targets.size(0)))
        loss.backward()
        clip_grad_norm_(model.parameters(), max_norm=1.)
        optimizer.step()
    for lv, n in pending:                    # one read-back per epoch
        losses.update(lv.item(), n)
    logger.info('Avg Loss: %.5f' % losses.avg)
    return glob_step, lr_now


def test(model, test_dl):
    model.eval()
    preds = []
    with torch.no_grad():
        for i, (inp, tar) in enumerate(test_dl):
            inputs = inp.cuda(non_blocking=True) if torch.cuda.is_available() else inp
            preds.append(model(inputs)[-1])
    preds = torch.cat(preds, 0).cpu().numpy() if preds else np.zeros((0, 45), np.float32)
    return test_dl.dataset.evaluate(preds)


def main(argv=None, train_dl=None, test_dl=None, log_root='refiner/experiments'):
    args = parse_args(argv)
    err_best = 1000
    log_dir = os.path.join(log_root, args.exp)
    os.makedirs(log_dir, exist_ok=True)
    logger = logging.getLogger("refiner")
    logger.setLevel(logging.INFO)
    model = get_model(weights=None)
    if torch.cuda.is_available():
        model = model.cuda()
    model.apply(weight_init)
    criterion = nn.MSELoss(reduction='mean')
    from lib.utils.utils import FusedAdam
    optimizer = FusedAdam(list(model.parameters()), lr=args.lr)
    glob_step, lr_now, start_epoch = 0, args.lr, 0
    if args.load:
        logger.info(">>> loading ckpt from '{}'".format(args.load))
        ckpt = torch.load(args.load, map_location='cpu', weights_only=False)
        start_epoch, err_best = ckpt['epoch'], ckpt['err']
        glob_step, lr_now = ckpt['step'], ckpt['lr']
        model.load_state_dict(ckpt['state_dict'])
        optimizer.load_state_dict(ckpt['optimizer'])
        logger.info(">>> ckpt loaded (epoch: {} | err: {})".format(start_epoch, err_best))
    if train_dl is None or test_dl is None:
        from .data import Human36M
        train_dl = torch.utils.data.DataLoader(Human36M(is_train=True), batch_size=64, shuffle=True)
        test_dl = torch.utils.data.DataLoader(Human36M(is_train=False), batch_size=64, shuffle=False)
    if args.mode == 'train':
        logger.info("Starting training for {} epoch(s)".format(args.num_epochs))
        for epoch in range(args.num_epochs):
            logger.info('%s | %s | lr: %.6f' % (epoch, args.num_epochs, lr_now))
            glob_step, lr_now = train(model, train_dl, optimizer, glob_step, lr_now, criterion, args, logger)
            error = test(model, test_dl)
            is_best = error < err_best
            err_best = min(error, err_best)
            save_ckpt({'epoch': epoch + 1, 'lr': lr_now, 'step': glob_step, 'err': error,
                       'state_dict': model.state_dict(), 'optimizer': optimizer.state_dict()},
                      ckpt_path=log_dir, is_best=is_best)
            if is_best:
                logger.info('Found new best, error: %s' % error)
        return err_best
    if args.mode == 'test':
        return test(model, test_dl)
    print('mode input error!')


if __name__ == '__main__':
    main()
