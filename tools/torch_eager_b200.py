"""Kernel-for-kernel context for the bench number: the SAME network and
step -- PoseResNet-50 (VOLUME, J16 D64) forward, soft-argmax integral loss, backward, Adam, 128
images of 256x256 -- written with stock torch.nn modules and run by torch eager on one B200
(cuDNN / cuBLAS / ATen kernels), in fp32 (TF32 off: the reference's numerics class on CPU) and
with TF32 allowed (the reference's default GPU path).  None of this repo's kernels run here.

    python tools/torch_eager_b200.py [images=128] [steps=10]
"""
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


class Bottleneck(nn.Module):
    def __init__(self, inp, planes, stride=1, down=None):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(inp, planes, 1, bias=False), nn.BatchNorm2d(planes, momentum=0.1)
        self.conv2, self.bn2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False), nn.BatchNorm2d(planes, momentum=0.1)
        self.conv3, self.bn3 = nn.Conv2d(planes, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4, momentum=0.1)
        self.down = down

    def forward(self, x):
        r = x if self.down is None else self.down(x)
        o = F.relu(self.bn1(self.conv1(x)))
        o = F.relu(self.bn2(self.conv2(o)))
        return F.relu(self.bn3(self.conv3(o)) + r)


class PoseNet(nn.Module):
    def __init__(self, J=16, D=64, layers=(3, 4, 6, 3)):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64, momentum=0.1),
                                  nn.ReLU(), nn.MaxPool2d(3, 2, 1))
        inp, blocks = 64, []
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            for b in range(n):
                s = 2 if (b == 0 and i > 0) else 1
                down = None
                if b == 0:
                    down = nn.Sequential(nn.Conv2d(inp, planes * 4, 1, s, bias=False),
                                         nn.BatchNorm2d(planes * 4, momentum=0.1))
                blocks.append(Bottleneck(inp, planes, s, down))
                inp = planes * 4
        self.trunk = nn.Sequential(*blocks)
        head = []
        for _ in range(3):
            head += [nn.ConvTranspose2d(inp, 256, 4, 2, 1, bias=False), nn.BatchNorm2d(256, momentum=0.1), nn.ReLU()]
            inp = 256
        self.head = nn.Sequential(*head)
        self.final = nn.Conv2d(256, J * D, 1)

    def forward(self, x):
        return self.final(self.head(self.trunk(self.stem(x))))


def integral_l1(preds, J, D, label):
    n = preds.shape[0]
    sm = torch.softmax(preds.reshape(n, J, -1), 2).reshape(n, J, D, D, D)
    ar = torch.arange(D, device=preds.device, dtype=torch.float32)
    c = torch.stack([(sm.sum((2, 3)) * ar).sum(2) / D - 0.5, (sm.sum((2, 4)) * ar).sum(2) / D - 0.5,
                     (sm.sum((3, 4)) * ar).sum(2) / D - 0.5], 2).reshape(n, J * 3)
    return (c - label).abs().sum() / n


def run(n_img, steps, tf32, channels_last):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = PoseNet().to(dev).train()
    x = torch.randn(n_img, 3, 256, 256, device=dev)
    if channels_last:
        m = m.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    lab = torch.rand(n_img, 48, device=dev) - 0.5
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = integral_l1(m(x), 16, 64, lab)
        loss.backward()
        opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("| torch %s eager, %s, %s | %.2f | %.1f |" % (torch.__version__, "TF32 allowed" if tf32 else "fp32 (TF32 off)",
                                                      "channels_last" if channels_last else "NCHW", ms,
                                                      n_img / 4 / (ms / 1e3)))


if __name__ == "__main__":
    n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    print("| stock torch.nn PoseResNet-50 step on one B200 (%d images) | ms/step | view-tuples/s |\n|---|---:|---:|" % n_img)
    for tf32 in (False, True):
        for cl in (False, True):
            try:
                run(n_img, steps, tf32, cl)
            except Exception as e:       # e.g. out of memory in one layout
                print("| tf32=%s channels_last=%s | failed: %s | |" % (tf32, cl, str(e)[:80]))
            torch.cuda.empty_cache()
