"""TEST INFRASTRUCTURE ONLY -- torch-CPU fp32 restatement of PoseResNet
(reference lib/models/pose3d_resnet.py) as a pure function of a state_dict.

It is the floating-point oracle for the CNN part of the hot path (the task's
"plain PyTorch fp32 reference"): parity tests compare the CUDA path's heatmaps
and gradients against it with tolerance 1e-3 (max|d| / max|ref| per tensor).
Pinned against the unmodified reference module through
tests/golden/net_*.npz (tests/test_oracle_pinned.py).  Never imported by the
product path.
"""
import collections

import torch
import torch.nn.functional as F

BN_MOMENTUM = 0.1   # pose3d_resnet.py:8
BN_EPS = 1e-5       # torch.nn.BatchNorm2d default used by the reference

# pose3d_resnet.py:288-292
RESNET_SPEC = {18: ("basic", [2, 2, 2, 2]), 34: ("basic", [3, 4, 6, 3]),
               50: ("bottleneck", [3, 4, 6, 3]), 101: ("bottleneck", [3, 4, 23, 3]),
               152: ("bottleneck", [3, 8, 36, 3])}


def deconv_cfg(k):
    """pose3d_resnet.py:145-156 -> (kernel, padding, output_padding)."""
    return {4: (4, 1, 0), 3: (3, 1, 1), 2: (2, 0, 0)}[k]


def param_shapes(num_layers=50, num_joints=17, volume=True, depth_res=64,
                 deconv_filters=(256, 256, 256), deconv_kernels=(4, 4, 4),
                 deconv_with_bias=False, final_kernel=1):
    """Ordered {state_dict key: shape} identical to the reference module's
    state_dict() (pose3d_resnet.py:93-126)."""
    kind, layers = RESNET_SPEC[num_layers]
    exp = 4 if kind == "bottleneck" else 1
    d = collections.OrderedDict()

    def bn(prefix, c):
        d[prefix + ".weight"] = (c,)
        d[prefix + ".bias"] = (c,)
        d[prefix + ".running_mean"] = (c,)
        d[prefix + ".running_var"] = (c,)
        d[prefix + ".num_batches_tracked"] = ()

    d["conv1.weight"] = (64, 3, 7, 7)
    bn("bn1", 64)
    inpl = 64
    for li, (planes, nb) in enumerate(zip((64, 128, 256, 512), layers)):
        stride = 1 if li == 0 else 2
        for b in range(nb):
            p = "layer%d.%d" % (li + 1, b)
            s = stride if b == 0 else 1
            if kind == "bottleneck":
                d[p + ".conv1.weight"] = (planes, inpl, 1, 1)
                bn(p + ".bn1", planes)
                d[p + ".conv2.weight"] = (planes, planes, 3, 3)
                bn(p + ".bn2", planes)
                d[p + ".conv3.weight"] = (planes * 4, planes, 1, 1)
                bn(p + ".bn3", planes * 4)
            else:
                d[p + ".conv1.weight"] = (planes, inpl, 3, 3)
                bn(p + ".bn1", planes)
                d[p + ".conv2.weight"] = (planes, planes, 3, 3)
                bn(p + ".bn2", planes)
            if b == 0 and (s != 1 or inpl != planes * exp):
                d[p + ".downsample.0.weight"] = (planes * exp, inpl, 1, 1)
                bn(p + ".downsample.1", planes * exp)
            inpl = planes * exp
    for i, (nf, k) in enumerate(zip(deconv_filters, deconv_kernels)):
        d["deconv_layers.%d.weight" % (3 * i)] = (inpl, nf, k, k)
        if deconv_with_bias:
            d["deconv_layers.%d.bias" % (3 * i)] = (nf,)
        bn("deconv_layers.%d" % (3 * i + 1), nf)
        inpl = nf
    out_ch = num_joints * depth_res if volume else num_joints
    d["final_layer.weight"] = (out_ch, inpl, final_kernel, final_kernel)
    d["final_layer.bias"] = (out_ch,)
    if not volume:
        d["depth_fc.weight"] = (num_joints * depth_res, 2048)
        d["depth_fc.bias"] = (num_joints * depth_res,)
    return d


def _bn(sd, prefix, x, training, new_stats):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if new_stats is not None:
            n = x.numel() / x.shape[1]
            with torch.no_grad():
                new_stats[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
                new_stats[prefix + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * n / max(n - 1, 1)
    else:
        mean, var = rm, rv
    xh = (x - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + BN_EPS)
    return xh * w[None, :, None, None] + b[None, :, None, None]


def forward(sd, x, num_layers=50, volume=True, image_size=(256, 256),
            deconv_kernels=(4, 4, 4), final_kernel=1, training=True,
            new_stats=None, taps=None, margins=None, forced_masks=None):
    """pose3d_resnet.py:185-212.  `taps`: optional dict filled with named
    intermediate activations (for per-layer parity checks)."""
    kind, layers = RESNET_SPEC[num_layers]

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    def relu(t):
        # `margins` collects min|pre-activation|: ReLU' is discontinuous at 0, so a
        # gradient parity test is only meaningful when no pre-activation sits within
        # rounding noise of 0 (tests pick seeds with a healthy margin).
        if margins is not None:
            margins.append(float(t.detach().abs().min()))
        if forced_masks is not None:
            # ReLU with an externally supplied mask (call order): lets a gradient parity
            # test use the SAME activation pattern as the implementation under test, so
            # pre-activations within rounding noise of 0 cannot flip the comparison.
            return t * forced_masks.pop(0).to(t.dtype)
        return F.relu(t)

    x = F.conv2d(x, sd["conv1.weight"], None, 2, 3)
    tap("conv1", x)
    x = relu(_bn(sd, "bn1", x, training, new_stats))
    x = F.max_pool2d(x, 3, 2, 1)
    tap("maxpool", x)
    inpl = 64
    for li, (planes, nb) in enumerate(zip((64, 128, 256, 512), layers)):
        stride = 1 if li == 0 else 2
        for b in range(nb):
            p = "layer%d.%d" % (li + 1, b)
            s = stride if b == 0 else 1
            res = x
            if kind == "bottleneck":      # :68-88
                o = F.conv2d(x, sd[p + ".conv1.weight"])
                o = relu(_bn(sd, p + ".bn1", o, training, new_stats))
                o = F.conv2d(o, sd[p + ".conv2.weight"], None, s, 1)
                o = relu(_bn(sd, p + ".bn2", o, training, new_stats))
                o = F.conv2d(o, sd[p + ".conv3.weight"])
                o = _bn(sd, p + ".bn3", o, training, new_stats)
            else:                          # :31-47
                o = F.conv2d(x, sd[p + ".conv1.weight"], None, s, 1)
                o = relu(_bn(sd, p + ".bn1", o, training, new_stats))
                o = F.conv2d(o, sd[p + ".conv2.weight"], None, 1, 1)
                o = _bn(sd, p + ".bn2", o, training, new_stats)
            if (p + ".downsample.0.weight") in sd:
                res = F.conv2d(x, sd[p + ".downsample.0.weight"], None, s)
                res = _bn(sd, p + ".downsample.1", res, training, new_stats)
            x = relu(o + res)
            tap(p, x)
    y = x
    for i, k in enumerate(deconv_kernels):
        kk, pad, opad = deconv_cfg(k)
        x = F.conv_transpose2d(x, sd["deconv_layers.%d.weight" % (3 * i)],
                               sd.get("deconv_layers.%d.bias" % (3 * i)), 2, pad, opad)
        x = relu(_bn(sd, "deconv_layers.%d" % (3 * i + 1), x, training, new_stats))
        tap("deconv%d" % i, x)
    x = F.conv2d(x, sd["final_layer.weight"], sd["final_layer.bias"], 1,
                 1 if final_kernel == 3 else 0)
    if volume:
        return x
    y = F.avg_pool2d(y, int(image_size[0] / 32), 1)
    y = y.reshape(y.shape[0], -1)
    y = F.linear(y, sd["depth_fc.weight"], sd["depth_fc.bias"])
    return x, y


def init_state(shapes, seed=0, scale_final=None):
    """Seeded synthetic weights (no checkpoints offline): kaiming-like normal
    for conv/deconv weights, BN gamma ~ U(0.5,1.5), beta ~ N(0,0.1),
    running_mean 0 / running_var 1.  Deterministic given (shapes, seed) using
    numpy so the GPU box regenerates identical weights without torch RNG drift."""
    import numpy as np
    rng = np.random.default_rng(seed)
    sd = collections.OrderedDict()
    for k, shp in shapes.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = torch.zeros(shp)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(shp)
        elif len(shp) == 1 and ("bn" in k or "downsample.1" in k or
                                 (k.startswith("deconv_layers") and int(k.split(".")[1]) % 3 == 1)):
            if k.endswith("weight"):
                sd[k] = torch.from_numpy(rng.uniform(0.5, 1.5, shp).astype(np.float32))
            else:
                sd[k] = torch.from_numpy((0.1 * rng.standard_normal(shp)).astype(np.float32))
        elif len(shp) == 1:   # conv / fc biases
            sd[k] = torch.from_numpy((0.05 * rng.standard_normal(shp)).astype(np.float32))
        else:
            fan_in = int(np.prod(shp[1:])) if not k.startswith("deconv_layers") else int(shp[0] * shp[2] * shp[3] / 4)
            std = (2.0 / fan_in) ** 0.5
            if k.startswith("final_layer") and scale_final is not None:
                std = scale_final
            sd[k] = torch.from_numpy((std * rng.standard_normal(shp)).astype(np.float32))
    return sd
